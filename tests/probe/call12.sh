cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 70 python bench.py --decode-only > gpurun_out/r02l_bench_decode_only.json 2> gpurun_out/r02l_bench.err; tail -2 gpurun_out/r02l_bench.err; cut -c1-400 gpurun_out/r02l_bench_decode_only.json
