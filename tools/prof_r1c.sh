# Round-1 final profiling pass (run under gpurun, one GPU). Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --no-extra --steps 3 --warmup 3 > /dev/null 2>&1   # warm page cache + scan cache
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1c.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused -s 4 -c 1 -o gpurun_out/prof_vg_fused_r1c -f python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icp_ -s 6 -c 5 -o gpurun_out/prof_icp_r1c -f python bench.py --only icp --steps 3 --warmup 3 > gpurun_out/ncu4c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kd_ -s 0 -c 12 -o gpurun_out/prof_kd_r1c -f python bench.py --only nn --steps 3 --warmup 3 > gpurun_out/ncu5c.log 2>&1
ls -la gpurun_out | tail -6
