#!/bin/bash
# ncu launch list of the fused-iteration kernels only (time, warp instructions, LSU wavefronts).  Usage: tools/ncu_launches.sh TAG
TAG=${1:-x}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --cache-control none -k regex:"ham_|snapped" -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-graphs --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
