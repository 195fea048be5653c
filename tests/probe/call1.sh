set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02b_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_gputests.log
tail -5 gpurun_out/r02b_gputests.log
( time python bench.py ) > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench.err; tail -3 gpurun_out/r02b_bench.err
python tests/probe/pcie_probe.py > gpurun_out/r02b_pcie_probe.txt 2>&1; cat gpurun_out/r02b_pcie_probe.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:zxc_decode_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02b_decode python tests/variant_probe.py 1024 zxc_b200/lib/libzxc.so.4 > gpurun_out/r02b_ncu.log 2>&1; tail -3 gpurun_out/r02b_ncu.log
ls -la gpurun_out | tail
