python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; tail -2 gpurun_out/bench_r1_f.err; python tools/show_bench.py gpurun_out/bench_r1_f.json | grep -E "VoxelGrid|^NN|^ICP|icp_|replay|terms"
