# Round-2 profiling pass (run under gpurun, one GPU). Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --no-extra --steps 3 --warmup 3 > /dev/null 2>&1   # warm page cache + scan cache
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused -s 4 -c 1 -o gpurun_out/prof_vg_fused_r2 -f python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu1_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearest_kernel -s 1 -c 1 -o gpurun_out/prof_nearest_r2 -f python bench.py --only nn --steps 3 --warmup 3 > gpurun_out/ncu2_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icp_ -s 6 -c 5 -o gpurun_out/prof_icp_r2 -f python bench.py --only icp --steps 3 --warmup 3 > gpurun_out/ncu3_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"minmax_bulk|onesweep|voxel_key|voxel_reduce" -s 8 -c 8 -o gpurun_out/prof_vg50m_r2 -f python tools/vg_one.py 1 10 5 > gpurun_out/ncu4_r2.log 2>&1
ls -la gpurun_out | tail -8
