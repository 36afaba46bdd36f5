for v in q8 qmorton q8morton; do
  export PCG_LIB=$PWD/build_variants/libpcg_$v.so
  python bench.py --only nn --steps 5 --warmup 3 > gpurun_out/bench_nn_$v.json 2> gpurun_out/bench_nn_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_nn_$v.json')); print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],3), {k:(x['launches'],round(x['avg_us'],1)) for k,x in d['kernels'].items()})"
  python bench.py --only icp --steps 5 --warmup 3 > gpurun_out/bench_icp_$v.json 2> gpurun_out/bench_icp_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_icp_$v.json')); print('$v', {m:(round(v['value'],1), round([x['avg_us'] for k,x in v['kernels'].items() if 'terms' in k][0],1)) for m,v in d['modes'].items()})"
done
