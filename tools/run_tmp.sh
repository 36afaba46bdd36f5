python -m pytest tests/test_gpu_edges.py -x -q -k "threads" > gpurun_out/pytest_thr.log 2>&1; tail -25 gpurun_out/pytest_thr.log
