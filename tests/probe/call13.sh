cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( echo "# compute-sanitizer --tool memcheck python tests/sanitize_invalid.py (ZXC_SANITIZE_QUICK=1) on the final build of round 2: reject vectors, valid vectors, damaged level 1/3/6/7 frames"; ZXC_SANITIZE_QUICK=1 timeout 50 compute-sanitizer --tool memcheck python tests/sanitize_invalid.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize_invalid|ERROR SUMMARY|Invalid|at 0x|by thread" | head -40 ) > gpurun_out/r02m_sanitizer_memcheck.txt
cat gpurun_out/r02m_sanitizer_memcheck.txt | tail -5
