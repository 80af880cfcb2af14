#!/bin/bash
(FMHR_IDX_COPY_ALWAYS=1 tools/run_variants.sh tail; tools/run_variants.sh tail rsv tail rsv) > gpurun_out/s3h_variants.txt 2>&1; cat gpurun_out/s3h_variants.txt
FMHR_B200_LIB=/root/repo/variants/libfmhr_trace.so python tools/trace_timeline.py > gpurun_out/s3h_trace.txt 2>&1; tail -15 gpurun_out/s3h_trace.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s3h_tests.log 2>&1; tail -3 gpurun_out/s3h_tests.log
