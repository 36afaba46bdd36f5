python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all3.log 2>&1; tail -3 gpurun_out/pytest_all3.log
python bench.py > gpurun_out/bench_r1_i.json 2> gpurun_out/bench_r1_i.err; tail -1 gpurun_out/bench_r1_i.err; python tools/show_bench.py gpurun_out/bench_r1_i.json 2>/dev/null | grep -E "^VoxelGrid|^NN|^ICP|icp_|roofline"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_i.json 2> gpurun_out/bench_ref_i.err; head -c 300 gpurun_out/bench_ref_i.json
