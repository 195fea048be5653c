cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python tests/variant_probe.py 2048 zxc_b200/lib/libzxc.so.4 build/var_c1/libzxc.so.4 build/var_c2/libzxc.so.4 build/var_c3/libzxc.so.4 build/var_c4/libzxc.so.4 build/var_call/libzxc.so.4 build/var_c134/libzxc.so.4 build/var_o8n6/libzxc.so.4 zxc_b200/lib/libzxc.so.4 2>&1 | tee gpurun_out/r02k_variants.txt
best=$(python - <<'PY'
import re
names=["base","c1","c2","c3","c4","call","c134","o8n6","base"]
rows=[l for l in open("gpurun_out/r02k_variants.txt") if "GB/s" in l]
vals=[(float(re.search(r"([\d.]+) GB/s",l).group(1)), "identical=True" in l) for l in rows]
base=max(v for (v,ok),n in zip(vals,names) if n=="base")
cand=[(v,n) for (v,ok),n in zip(vals,names) if n!="base" and ok]
v,n=max(cand)
print(n if v>base*1.01 else "none")
PY
)
echo "best=$best" | tee gpurun_out/r02k_best.txt
if [ "$best" != "none" ]; then
  ( time ZXC_B200_LIB=$GRAFT_REPO_ROOT/build/var_$best/libzxc.so.4 timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02k_gputests.log 2>&1; echo "pytest rc=$? lib=var_$best" >> gpurun_out/r02k_gputests.log
  tail -4 gpurun_out/r02k_gputests.log
fi
