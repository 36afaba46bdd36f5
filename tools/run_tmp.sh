python -m pytest tests/test_gpu_next.py tests/test_gpu_large.py -x -q -k "multi_kernel or 50m" > gpurun_out/pytest_mk.log 2>&1; tail -5 gpurun_out/pytest_mk.log
python bench.py --config5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_b.json 2> gpurun_out/bench_cfg5_b.err; tail -2 gpurun_out/bench_cfg5_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cfg5_b.json'))
c=d['extra']['config5_50m']
vg=c['voxelgrid']; print('VG 50M', round(vg['value_mpts'],1), 'Mpts/s', round(vg['ms_per_step'],3), 'ms', {k:(v['launches'],round(v['avg_us'],1)) for k,v in vg['kernels'].items()})
print('index build ms', c['index_build_ms'], 'nearest', c['nearest']['value']/1e6, 'Mq/s', c['nearest']['ms_per_step'], 'range', c['range'])
PY
