set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2_pytest2.log 2>&1
tail -30 gpurun_out/r2_pytest2.log
