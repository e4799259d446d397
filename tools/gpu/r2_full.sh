set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L; nproc; free -g | head -2; df -h /tmp | tail -1
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2_pytest_full.log 2>&1
grep -E "passed|failed" gpurun_out/r2_pytest_full.log
( time timeout 1500 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
( time timeout 1500 python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -c 2500 gpurun_out/r2_bench_full.json; tail -c 1200 gpurun_out/r2_bench_ref.json; tail -5 gpurun_out/r2_bench_full.err gpurun_out/r2_bench_ref.err
