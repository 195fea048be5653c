cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 280 python tests/variant_probe.py 2048 build/var_base/libzxc.so.4 zxc_b200/lib/libzxc.so.4 build/var_bf/libzxc.so.4 build/var_sti/libzxc.so.4 build/var_stl/libzxc.so.4 build/var_base/libzxc.so.4 zxc_b200/lib/libzxc.so.4 build/var_bf/libzxc.so.4 2>&1 | tee gpurun_out/r02i_variants.txt
timeout 120 python tests/dict_variant_probe.py 262144 build/var_base/libzxc.so.4 zxc_b200/lib/libzxc.so.4 build/var_bf/libzxc.so.4 2>&1 | tail -3 | tee gpurun_out/r02i_dict.txt
