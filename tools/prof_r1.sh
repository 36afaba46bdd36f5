# Round-1 profiling pass (run under gpurun, one GPU). Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --no-extra --steps 3 --warmup 3 > /dev/null 2>&1   # warm page cache + scan cache
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused -s 4 -c 1 -o gpurun_out/prof_vg_fused_r1 -f python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearest_simple -s 1 -c 1 -o gpurun_out/prof_nearest_r1 -f python bench.py --only nn --steps 3 --warmup 3 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 4 -c 2 -o gpurun_out/prof_onesweep_nn_r1 -f python bench.py --only nn --steps 3 --warmup 3 > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icp_ -s 6 -c 3 -o gpurun_out/prof_icp_r1 -f python bench.py --only icp --steps 3 --warmup 3 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out | tail -12
