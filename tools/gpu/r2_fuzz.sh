set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 420 python tools/gpu_fuzz.py 7000 120 ) > gpurun_out/r2_fuzz.log 2>&1
tail -3 gpurun_out/r2_fuzz.log; grep -c "^ok" gpurun_out/r2_fuzz.log; grep "^FAIL" gpurun_out/r2_fuzz.log | head
