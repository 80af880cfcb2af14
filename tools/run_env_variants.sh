#!/bin/bash
# Bench the product library under different environment settings: tools/run_env_variants.sh "FMHR_VIEW_GROUPS=2" "FMHR_NO_SPEC=1" ...
# ("base" = no extra setting).  One line per setting: iters/s + the CUDA-event stage times.
for v in "$@"; do
  if [ "$v" = base ]; then pre=""; else pre="$v"; fi
  env $pre python bench.py --steps 600 --warmup 100 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', round(d['value'],1), round(d['ms_per_step']*1000,1), 'us')"
done
