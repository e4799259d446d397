set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest1.log
PSC_DEBUG_DP=1 PSC_DEBUG_HOST=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_dbg.json 2> gpurun_out/r2_bench_dbg.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err
PSC_DEBUG_HOST=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_host.json 2> gpurun_out/r2_bench_host.err
