cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02j_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_gputests.log
tail -4 gpurun_out/r02j_gputests.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:zxc_decode_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02j_decode python tests/variant_probe.py 1024 zxc_b200/lib/libzxc.so.4 > gpurun_out/r02j_ncu.log 2>&1; tail -1 gpurun_out/r02j_ncu.log
python profiles/update_traffic.py gpurun_out/r02j_decode.ncu-rep 1024 r02j_decode_ncu_summary.txt > gpurun_out/r02j_traffic.json 2> gpurun_out/r02j_traffic.err
timeout 400 python bench.py > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench.err; tail -2 gpurun_out/r02j_bench.err; cut -c1-300 gpurun_out/r02j_bench_n1.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02j_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r02j_bench_under_ncu.log | cut -c1-120
