python -m pytest tests -x -q -m gpu > gpurun_out/pytest_final.log 2>&1; tail -4 gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -2 gpurun_out/bench_final_n1.err; head -c 1500 gpurun_out/bench_final_n1.json
