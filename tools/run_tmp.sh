python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_next.py tests/test_gpu_multi.py -x -q -k "icp or sum or replay or strict or hessian or gauss or shard or pipeline" > gpurun_out/pytest_icp.log 2>&1; tail -4 gpurun_out/pytest_icp.log
python bench.py --only icp --steps 5 --warmup 3 > gpurun_out/bench_icp_x.json 2> gpurun_out/bench_icp_x.err; tail -1 gpurun_out/bench_icp_x.err
python -c "
import json; d=json.load(open('gpurun_out/bench_icp_x.json'))
for m,v in d['modes'].items(): print(m, round(v['value'],1), round(v['ms_per_alignment'],3), {k:round(x['avg_us'],1) for k,x in v['kernels'].items()})"
