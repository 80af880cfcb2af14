#!/bin/bash
# Session-3 tuning cycle: bench line of every variant library given, live schedule of the trace variant, fused parity tests on one variant.
#   tools/variant_cycle.sh TAG TESTVARIANT name1 name2 ...
TAG=$1; shift
TV=$1; shift
mkdir -p gpurun_out
tools/run_variants.sh "$@" > gpurun_out/${TAG}_variants.txt 2>&1
cat gpurun_out/${TAG}_variants.txt
if [ -f variants/libfmhr_trace.so ]; then
  FMHR_B200_LIB=/root/repo/variants/libfmhr_trace.so python tools/trace_timeline.py --out gpurun_out/${TAG}_trace.json > gpurun_out/${TAG}_trace.txt 2>&1
  cat gpurun_out/${TAG}_trace.txt | tail -20
fi
if [ "$TV" != none ]; then
  FMHR_B200_LIB=/root/repo/variants/libfmhr_$TV.so timeout 600 python -m pytest tests/test_gpu_ham.py tests/test_gpu_parity_full.py -x -q -m gpu > gpurun_out/${TAG}_tests_$TV.log 2>&1
  tail -3 gpurun_out/${TAG}_tests_$TV.log
fi
