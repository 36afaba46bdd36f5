PCG_NN_LEAF_VOTES=24 ncu --set full --clock-control none --import-source on -k regex:nearest_kernel -s 1 -c 1 -o gpurun_out/prof_nearest_persist -f python bench.py --only nn --steps 3 --warmup 3 > gpurun_out/ncu_nn2.log 2>&1
PCG_NN_KERNEL=simple ncu --set full --clock-control none --import-source on -k regex:nearest_simple -s 1 -c 1 -o gpurun_out/prof_nearest_simple_perm -f python bench.py --only nn --steps 3 --warmup 3 > gpurun_out/ncu_nn3.log 2>&1
ls -la gpurun_out/*.ncu-rep
