set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or two_stream or chunked or dense_component or workload_shape or randomised" ) > gpurun_out/r2_iter_pytest.log 2>&1
tail -5 gpurun_out/r2_iter_pytest.log
PSC_DEBUG_HOST=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-files > gpurun_out/r2_iter_host.json 2> gpurun_out/r2_iter_host.err
PSC_DEBUG_UNITS=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-files > gpurun_out/r2_iter_units.json 2> gpurun_out/r2_iter_units.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/r2_iter_bench.json 2> gpurun_out/r2_iter_bench.err
cat gpurun_out/r2_iter_bench.json | head -c 400
