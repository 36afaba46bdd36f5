python -m pytest tests/test_gpu_parity.py -x -q -k "pairs or farm" 2>&1 | tail -2
for t in 8 8 16 32; do PCG_FARM_STREAMS=$t python bench.py --only farm --steps 5 --warmup 3 > gpurun_out/bench_farm_s$t.json 2> gpurun_out/bench_farm_s$t.err; tail -1 gpurun_out/bench_farm_s$t.err; python -c "
import json; d=json.load(open('gpurun_out/bench_farm_s$t.json')); print($t, {m:(round(v['value'],1), round(v['ms_per_pair'],3), v['failed']) for m,v in d['modes'].items()})"; done
