#!/bin/bash
# Session-3 final single-GPU evidence run (ordered by importance; every step under its own timeout).
T=s3f; mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${T}_tests.log
timeout 300 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_default.json')); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['within_tolerance'])"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ham_" -c 80 --csv --log-file gpurun_out/${T}_launchlist.csv \
    python bench.py --steps 3 --warmup 3 --no-graphs --no-e2e --no-cpu-baseline > gpurun_out/${T}_ncu_ll.log 2>&1; echo "ncu launch list rc=$?"
timeout 200 tools/ncu_launches.sh ${T}w
FMHR_B200_LIB=/root/repo/variants/libfmhr_trace.so timeout 120 python tools/trace_timeline.py --out gpurun_out/${T}_trace.json > gpurun_out/${T}_trace.txt 2>/dev/null; tail -14 gpurun_out/${T}_trace.txt
RE="ham_vertex_prep|ham_normals|ham_regulariser|ham_reg_grad|ham_trirec|coverage_meshlet|ham_scan|ham_shade|ham_aa_loss|ham_pair_bwd|ham_finalize|ham_normal_grad|ham_update_pass2"
timeout 400 tools/ncu_full.sh $T "$RE"
for w in two_hands_48x512x334 demo_full capture_16x1024x1024; do
  timeout 120 python bench.py --workload $w --no-e2e --no-cpu-baseline --steps 300 > gpurun_out/${T}_cfg_$w.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/${T}_cfg_$w.json')); print('$w', round(d['value'],1), 'it/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))"; done
