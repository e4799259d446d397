set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or two_stream or chunked or dense_component or workload_shape or randomised" ) > gpurun_out/r2_iter_pytest.log 2>&1
tail -5 gpurun_out/r2_iter_pytest.log
for v in 0 16 24; do
PSC_CRIT_SMS=$v PSC_DEBUG_HOST=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-files > gpurun_out/r2_crit$v.json 2> gpurun_out/r2_crit$v.err
head -c 300 gpurun_out/r2_crit$v.json; echo
done
