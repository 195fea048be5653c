cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tests/variant_probe.py 2048 build/var_s0/libzxc.so.4 zxc_b200/lib/libzxc.so.4 build/var_s8/libzxc.so.4 build/var_f2/libzxc.so.4 build/var_s0/libzxc.so.4 zxc_b200/lib/libzxc.so.4 build/var_f2/libzxc.so.4 2>&1 | tee gpurun_out/r02f_variants.txt
