python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; tail -3 gpurun_out/bench_r1_c.err; python tools/show_bench.py gpurun_out/bench_r1_c.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c.json 2>/dev/null; cat gpurun_out/bench_ref_c.json | cut -c1-400
