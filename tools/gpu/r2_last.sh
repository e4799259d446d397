set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 330 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_last.log 2>&1
grep -E "passed|failed" gpurun_out/r2_pytest_last.log
