python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; tail -2 gpurun_out/bench_r1_e.err; python tools/show_bench.py gpurun_out/bench_r1_e.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_e.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_ref_e.json')); print('reference arm', d['value'], d['unit'], d['cpu_baseline']['cores'], 'threads', d['ms_per_step'],'ms/step')"
nproc; free -g | head -2
