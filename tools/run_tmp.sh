python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q > gpurun_out/pytest_kd.log 2>&1; tail -5 gpurun_out/pytest_kd.log
for ord in kd hilbert; do
  PCG_INDEX_ORDER=$ord python bench.py --only nn --steps 5 --warmup 3 > gpurun_out/bench_nn_$ord.json 2> gpurun_out/bench_nn_$ord.err; tail -2 gpurun_out/bench_nn_$ord.err
  PCG_INDEX_ORDER=$ord python bench.py --only icp --steps 5 --warmup 3 > gpurun_out/bench_icp_$ord.json 2> gpurun_out/bench_icp_$ord.err; tail -2 gpurun_out/bench_icp_$ord.err
done
python - <<'PY'
import json
for o in ("kd","hilbert"):
    d=json.load(open(f"gpurun_out/bench_nn_{o}.json"))
    print(o, "NN", round(d["value"]/1e6,1), "Mq/s", round(d["ms_per_step"],3), "ms; build", d.get("index_build_ms"), {k:round(v["avg_us"],1) for k,v in d["kernels"].items()}, d.get("parity_sample"))
    d=json.load(open(f"gpurun_out/bench_icp_{o}.json"))
    for m,v in d["modes"].items(): print(o, "ICP", m, round(v["value"],1), round(v["ms_per_alignment"],3), {k:round(x["avg_us"],1) for k,x in v["kernels"].items()})
    print(o, d.get("parity"))
PY
