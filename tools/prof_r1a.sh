set -x
mkdir -p gpurun_out
python bench.py --no-extra --steps 3 --warmup 3 > /dev/null 2>&1   # warm the page cache + scan cache
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_vg_r1a.csv python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu_vg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel_reduce -s 3 -c 1 -o gpurun_out/prof_voxel_reduce_r1a -f python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu_vr.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:onesweep -s 12 -c 2 -o gpurun_out/prof_onesweep_r1a -f python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu_os.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:minmax_kernel -s 3 -c 1 -o gpurun_out/prof_minmax_r1a -f python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu_mm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearest_kernel -s 1 -c 1 -o gpurun_out/prof_nearest_r1a -f python bench.py --only nn --steps 3 --warmup 3 > gpurun_out/ncu_nn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icp_ -s 4 -c 2 -o gpurun_out/prof_icp_r1a -f python bench.py --only icp --steps 3 --warmup 3 > gpurun_out/ncu_icp.log 2>&1
ls -la gpurun_out
