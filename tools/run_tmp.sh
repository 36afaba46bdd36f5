python -m pytest tests/test_gpu_next.py tests/test_gpu_multi.py -x -q -k "sharded" > gpurun_out/pytest_sh.log 2>&1; tail -6 gpurun_out/pytest_sh.log
