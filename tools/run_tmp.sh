python -m pytest tests/test_gpu_next.py -x -q -k "sharded" > gpurun_out/pytest_sh.log 2>&1; tail -12 gpurun_out/pytest_sh.log
