show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
if 'modes' in d:
    print({m:(round(v['ms_per_alignment'],3), {k:round(x['avg_us'],1) for k,x in v['kernels'].items() if 'icp' in k}) for m,v in d['modes'].items()})
else:
    print('NN %.1f Mq/s %.2f ms' % (d['value']/1e6, d['ms_per_step']), {k:(v['launches'],round(v['avg_us'],1)) for k,v in d['kernels'].items()}, d['index_build_ms'])"; }
PCG_LIB=build_variants/libpcg_nnstats.so python tools/nn_stats.py
python -m pytest tests -m gpu -x -q -k "not 50m" 2>&1 | tail -2
python bench.py --only nn --steps 5 --warmup 3 2>/dev/null | show; python bench.py --only icp --steps 5 --warmup 3 2>/dev/null | show
