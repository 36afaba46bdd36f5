python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_large.py -x -q > gpurun_out/pytest_kd4.log 2>&1; tail -4 gpurun_out/pytest_kd4.log
PCG_LIB=$PWD/build_variants/libpcg_vgtiming.so python tools/vg_stamps.py | grep -E "last block|staged|reduced|syncC" | tail -5
python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/bench_vg_j.json 2> gpurun_out/bench_vg_j.err; tail -1 gpurun_out/bench_vg_j.err
python tools/show_bench.py gpurun_out/bench_vg_j.json | head -6
python tools/build_prof.py 15625 2>/dev/null | head -6
python tools/build_prof.py 1875 2>/dev/null | head -5
