#!/bin/bash
# Multi-GPU evidence run (N ranks on one box): parity check, exchange timeline, weak / strong scaling lines.
# Usage: tools/multi_gpu_suite.sh N TAG [items...]   (under gpurun --gpus >= N)
# items: check trace weak_peer weak_nccl strong_cfg2 strong_cfg5 strong_cfg1full strong_cfg3 frames_cfg4   (default: all but frames_cfg4)
N=${1:-8}; TAG=${2:-m$N}; shift; shift
ITEMS=${*:-check trace weak_peer weak_nccl strong_cfg2 strong_cfg5 strong_cfg1full strong_cfg3}
if [ "$N" = 1 ]; then TR="timeout 600 python"; else TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"; fi
port() { if [ "$N" = 1 ]; then echo ""; else echo $((29620 + RANDOM % 300)); fi; }
mkdir -p gpurun_out
run() { # name, args...
  local name=$1; shift
  $TR $(port) bench.py --gpus $N --steps 300 --warmup 20 "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_$name.json"))
    print("$name", "N=$N value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), d["scaling"], "views/gpu", d["config"]["views_per_gpu"], "global", d["config"]["global_views"], "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("$name parse failed", e)
PY
}
for it in $ITEMS; do
  case $it in
    check) $TR $(port) tests/multi_gpu_check.py > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"; tail -1 gpurun_out/${TAG}_check.log;;
    trace) FMHR_B200_LIB=/root/repo/variants/libfmhr_trace.so $TR $(port) tools/trace_timeline.py > gpurun_out/${TAG}_trace.txt 2>&1; echo "trace rc=$?"
           grep -E "rank|pair_bwd|peer_|rendezvous|update_adam|normal_grad" gpurun_out/${TAG}_trace.txt | head -24;;
    weak_peer) run weak_peer;;
    weak_nccl) run weak_nccl --exchange nccl --no-e2e;;
    strong_cfg2) run strong_cfg2 --scaling strong --no-e2e --no-cpu-baseline;;
    strong_cfg5) run strong_cfg5 --scaling strong --workload stress_128x2048x2048 --no-e2e --no-cpu-baseline --steps 100;;
    strong_cfg1full) run strong_cfg1full --scaling strong --workload demo_full --no-e2e --no-cpu-baseline;;
    frames_cfg4) run frames_cfg4 --scaling frames --workload two_hands_48x512x334 --no-cpu-baseline;;
    strong_cfg3) run strong_cfg3 --scaling strong --workload capture_16x1024x1024 --no-e2e --no-cpu-baseline;;
  esac
done
