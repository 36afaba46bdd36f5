python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; tail -5 gpurun_out/pytest_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; tail -2 gpurun_out/bench_r1_g.err; python tools/show_bench.py gpurun_out/bench_r1_g.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err; tail -1 gpurun_out/bench_ref_g.err; head -c 600 gpurun_out/bench_ref_g.json
