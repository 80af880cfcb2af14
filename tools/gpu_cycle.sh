#!/bin/bash
# One GPU measurement cycle: parity tests, bench line, optional extras.  Usage: tools/gpu_cycle.sh TAG [extra command]
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/test_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/test_$TAG.log
tail -3 gpurun_out/test_$TAG.log
python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4))
    print({k: round(v * 1000, 1) for k, v in d["roofline"]["stage_ms"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
shift
if [ -n "$1" ]; then bash -c "$*"; fi
