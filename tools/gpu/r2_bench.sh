set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L; nproc; free -g | head -2; df -h /tmp | tail -1
( time timeout 1500 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
( time timeout 1500 python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -c 1500 gpurun_out/r2_bench_full.json; tail -c 1200 gpurun_out/r2_bench_ref.json
