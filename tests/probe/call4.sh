cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tests/variant_probe.py 2048 zxc_b200/lib/libzxc.so.4 build/var_y1/libzxc.so.4 build/var_y2/libzxc.so.4 build/var_y3/libzxc.so.4 build/var_y4/libzxc.so.4 build/var_y5/libzxc.so.4 build/var_y6/libzxc.so.4 build/var_y7/libzxc.so.4 zxc_b200/lib/libzxc.so.4 2>&1 | tee gpurun_out/r02d_variants.txt
python tests/dict_variant_probe.py 262144 zxc_b200/lib/libzxc.so.4 build/var_y2/libzxc.so.4 2>&1 | tail -2 | tee gpurun_out/r02d_dict.txt
ZXC_B200_UNITS=0 python tests/dict_variant_probe.py 262144 zxc_b200/lib/libzxc.so.4 build/var_y2/libzxc.so.4 build/var_y4/libzxc.so.4 2>&1 | tail -3 | tee -a gpurun_out/r02d_dict.txt
