python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py -x -q -k "slots or nearest or Nearest or range or icp" > gpurun_out/pytest_kd6.log 2>&1; tail -3 gpurun_out/pytest_kd6.log
python tools/build_prof.py 15625 2>/dev/null | head -3
for i in 1 2; do python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/bench_vg_n$i.json 2> gpurun_out/bench_vg_n$i.err; python -c "
import json; d=json.load(open('gpurun_out/bench_vg_n$i.json')); print(round(d['value'],1), d['ms_per_step'], json.dumps(d['e2e'])[:330])"; done
