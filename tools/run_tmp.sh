python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q > gpurun_out/pytest_kd5.log 2>&1; tail -4 gpurun_out/pytest_kd5.log
python tools/build_prof.py 15625 2>/dev/null | head -7
python tools/build_prof.py 1875 2>/dev/null | head -6
