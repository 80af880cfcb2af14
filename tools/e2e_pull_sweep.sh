#!/bin/bash
# end-to-end rate against the number of SMs the pull kernel occupies (FMHR_PULL_BLOCKS blocks of 256 threads)
for b in "$@"; do FMHR_PULL_BLOCKS=$b python bench.py --no-cpu-baseline --steps 100 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('pull blocks $b: value', round(d['value'],1), 'e2e', round(e['value'],1), 'sync', round(e['sync_every_step']['value'],1))"; done
