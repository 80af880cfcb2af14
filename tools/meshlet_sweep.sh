for t in 1024 512 256; do FMHR_MESHLET_TRIS=$t python bench.py --steps 600 --warmup 100 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$t', round(d['value'],1), {k: round(x*1000,1) for k,x in d['roofline']['stage_ms'].items()})"; done
