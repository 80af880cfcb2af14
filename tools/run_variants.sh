#!/bin/bash
# Bench each tuning variant: tools/run_variants.sh name1 name2 ...
for v in "$@"; do
  if [ "$v" = base ]; then unset FMHR_B200_LIB; else export FMHR_B200_LIB=/root/repo/variants/libfmhr_$v.so; fi
  python bench.py --steps 600 --warmup 100 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', round(d['value'],1), {k: round(x*1000,1) for k,x in d['roofline']['stage_ms'].items()})"
done
