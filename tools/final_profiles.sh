#!/bin/bash
# Evidence run for profiles/: bench line, ncu launch list (recipe flags), ncu --set full of every kernel of the iteration.
TAG=${1:-final}
python bench.py --steps 1000 --warmup 200 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
# launch list exactly as B200_PROFILING.md prescribes (cold caches, serialised) - only the iteration's kernels
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ham_" -c 80 --csv --log-file gpurun_out/launchlist_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-graphs --no-e2e --no-cpu-baseline > gpurun_out/ncu_ll_$TAG.log 2>&1
echo "ncu launch list rc=$?"
tools/ncu_launches.sh ${TAG}w
RE="ham_vertex_prep|ham_normals|ham_regulariser|ham_reg_grad|ham_trirec|coverage_meshlet|ham_scan|ham_shade|ham_aa_loss|ham_pair_bwd|ham_pixel_bwd|ham_finalize|ham_normal_grad|ham_update_pass2"
tools/ncu_full.sh $TAG "$RE"
