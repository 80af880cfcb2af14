#!/bin/bash
# ncu --set full capture of one steady-state launch of each hot kernel.  Usage: tools/ncu_full.sh TAG [regex]
TAG=${1:-x}
RE=${2:-"coverage_meshlet|ham_shade|ham_aa_loss|ham_pixel_bwd"}
N=$(echo "$RE" | tr '|' '\n' | wc -l)
ncu --set full --clock-control none --import-source on -k regex:"$RE" --launch-skip $((2 * N)) -c $N \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 3 --warmup 3 --no-graphs --no-e2e --no-cpu-baseline > gpurun_out/ncufull_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
