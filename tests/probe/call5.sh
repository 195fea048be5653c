cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tests/variant_probe.py 2048 zxc_b200/lib/libzxc.so.4 build/var_z1/libzxc.so.4 build/var_z2/libzxc.so.4 build/var_z3/libzxc.so.4 2>&1 | tee gpurun_out/r02e_variants.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02e_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_gputests.log
tail -5 gpurun_out/r02e_gputests.log
python bench.py > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench.err; tail -3 gpurun_out/r02e_bench.err; cut -c1-400 gpurun_out/r02e_bench_n1.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:zxc_decode_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02e_decode python tests/variant_probe.py 1024 zxc_b200/lib/libzxc.so.4 > gpurun_out/r02e_ncu.log 2>&1; tail -2 gpurun_out/r02e_ncu.log
