python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_large.py -x -q -k "voxel or Voxel or vg" > gpurun_out/pytest_vg.log 2>&1; tail -4 gpurun_out/pytest_vg.log
PCG_LIB=$PWD/build_variants/libpcg_vgtiming.so python tools/vg_stamps.py | grep -E "staged|syncC|reduced|slowest|median"
python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/bench_vg_l.json 2> gpurun_out/bench_vg_l.err; tail -1 gpurun_out/bench_vg_l.err
python tools/show_bench.py gpurun_out/bench_vg_l.json | head -6
