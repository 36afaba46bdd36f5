python -m pytest tests -m gpu -x -q -k voxelgrid 2>&1 | tail -2
PCG_LIB=build_variants/libpcg_vgtiming.so PCG_VG_PRINT=1 python bench.py --no-extra --steps 3 --warmup 3 2>&1 | grep "vg stamps" | tail -2
