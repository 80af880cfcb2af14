#!/bin/bash
# Final single-GPU evidence run: tests, default bench line, live schedule, cost model over the view count, the other BASELINE shapes.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
FMHR_B200_LIB=variants/libfmhr_trace.so python tools/trace_timeline.py --out gpurun_out/final_trace_1gpu.json > gpurun_out/final_trace_1gpu.txt 2>/dev/null; tail -15 gpurun_out/final_trace_1gpu.txt
for v in 6 12 24 36 48; do python bench.py --views $v --no-e2e --no-cpu-baseline --steps 500 --warmup 50 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('views $v ms', round(d['ms_per_step'],4))"; done | tee gpurun_out/final_views_sweep.txt
for w in demo_full capture_16x1024x1024 two_hands_48x512x334 stress_128x2048x2048; do st=300; [ $w = stress_128x2048x2048 ] && st=60
  python bench.py --workload $w --no-e2e --no-cpu-baseline --steps $st > gpurun_out/final_cfg_$w.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/final_cfg_$w.json')); print('$w', round(d['value'],1), 'it/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))"; done
python bench.py --workload capture_16x1024x1024 --ncc --steps 200 > gpurun_out/final_cfg_capture_ncc.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/final_cfg_capture_ncc.json')); print('capture+ncc', round(d['value'],1), 'it/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))"
