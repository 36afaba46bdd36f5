python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/b_fused2.json 2>gpurun_out/b_fused2.err; tail -2 gpurun_out/b_fused2.err; python tools/show_bench.py gpurun_out/b_fused2.json
