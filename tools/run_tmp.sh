python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q > gpurun_out/pytest_kd2.log 2>&1; tail -5 gpurun_out/pytest_kd2.log
python tools/build_prof.py 15625
python tools/build_prof.py 1875
PCG_INDEX_ORDER=hilbert python tools/build_prof.py 15625 | head -3
