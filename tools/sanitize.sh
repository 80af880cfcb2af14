#!/bin/bash
# compute-sanitizer passes over the small-scene GPU tests: memcheck, racecheck (shared-memory hazards
# of the coverage / scan / pixel kernels' per-warp queues), synccheck.  Usage (on a GPU box): tools/sanitize.sh [outdir]
# Logs: <outdir>/sanitize_<tool>.log; the summary lines ("ERROR SUMMARY") are what profiles/r2/sanitizer.md quotes.
OUT=${1:-gpurun_out}
mkdir -p $OUT
SEL_HAM='fused_phase_b_step or fused_phase_a_step or graph_replay or peer_exchange_single or empty_and_clipped or initialisation'
run() {  # tool, extra flags, pytest selection...
  tool=$1; shift; flags=$1; shift
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool $flags --error-exitcode 86 --launch-timeout 0 \
      python -m pytest -m gpu -q -x "$@" > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' $OUT/sanitize_$tool.log) summaries: $(grep 'ERROR SUMMARY' $OUT/sanitize_$tool.log | sort | uniq -c | tr '\n' ';')"
  tail -2 $OUT/sanitize_$tool.log
}
run memcheck "--leak-check no" tests/test_gpu_ops.py tests/test_gpu_ham.py -k "rasterize or interpolate or antialias or $SEL_HAM"
run racecheck "--racecheck-report all" tests/test_gpu_ops.py tests/test_gpu_ham.py -k "rasterize_bit_exact or antialias or fused_phase_b_step or fused_phase_a_step"
run synccheck "" tests/test_gpu_ops.py tests/test_gpu_ham.py -k "rasterize_bit_exact or fused_phase_b_step or peer_exchange_single"
