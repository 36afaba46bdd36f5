python -m pytest tests/test_gpu_edges.py -x -q -k "wider" > gpurun_out/pytest_wide.log 2>&1; tail -25 gpurun_out/pytest_wide.log
