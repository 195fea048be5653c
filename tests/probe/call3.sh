cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tests/variant_probe.py 2048 zxc_b200/lib/libzxc.so.4 build/var_w1/libzxc.so.4 build/var_w2/libzxc.so.4 build/var_w3/libzxc.so.4 build/var_w4/libzxc.so.4 build/var_w5/libzxc.so.4 build/var_w6/libzxc.so.4 build/var_w7/libzxc.so.4 build/var_w8/libzxc.so.4 2>&1 | tee gpurun_out/r02c_variants.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:zxc_decode_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02c_decode python tests/variant_probe.py 1024 zxc_b200/lib/libzxc.so.4 > gpurun_out/r02c_ncu.log 2>&1; tail -2 gpurun_out/r02c_ncu.log
