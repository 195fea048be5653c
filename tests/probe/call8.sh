cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python tests/variant_probe.py 1024 build/var_st0/libzxc.so.4 build/var_sti/libzxc.so.4 2>&1 | tee gpurun_out/r02h_variants.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:zxc_decode_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02h_stage python tests/variant_probe.py 1024 build/var_sti/libzxc.so.4 > gpurun_out/r02h_ncu.log 2>&1; tail -2 gpurun_out/r02h_ncu.log
