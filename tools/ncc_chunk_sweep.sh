for c in 1 2 3 5 8 15; do FMHR_NCC_CHUNK=$c python bench.py --workload capture_16x1024x1024 --ncc --no-e2e --no-cpu-baseline --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"; done
