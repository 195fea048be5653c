cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python tests/variant_probe.py 2048 build/var_st0/libzxc.so.4 zxc_b200/lib/libzxc.so.4 build/var_lit/libzxc.so.4 build/var_bf/libzxc.so.4 build/var_st0/libzxc.so.4 zxc_b200/lib/libzxc.so.4 build/var_lit/libzxc.so.4 build/var_bf/libzxc.so.4 2>&1 | tee gpurun_out/r02g_variants.txt
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02g_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_gputests.log
tail -6 gpurun_out/r02g_gputests.log
