PCG_LIB=$PWD/build_variants/libpcg_vgtiming.so python tools/vg_stamps.py | grep "last block"
