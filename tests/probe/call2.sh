cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tests/variant_probe.py 2048 build/var_v0/libzxc.so.4 build/var_v1/libzxc.so.4 build/var_v2/libzxc.so.4 build/var_v3/libzxc.so.4 build/var_v4/libzxc.so.4 build/var_v5/libzxc.so.4 build/var_v6/libzxc.so.4 build/var_v7/libzxc.so.4 build/var_v8/libzxc.so.4 build/var_v9/libzxc.so.4 2>&1 | tee gpurun_out/r02b_variants.txt
