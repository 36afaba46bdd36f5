for st in 1 2 4 8; do echo "== lane stride $st"; PCG_ICP_LANE_STRIDE=$st python bench.py --only icp --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({m:(round(v['ms_per_alignment'],3), {k:round(x['avg_us'],1) for k,x in v['kernels'].items() if 'terms' in k}) for m,v in d['modes'].items()})"; done
