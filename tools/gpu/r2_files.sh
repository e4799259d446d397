set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2_bench_files.json 2> gpurun_out/r2_bench_files.err
python3 -c "
import json; d=json.load(open('gpurun_out/r2_bench_files.json')); print(round(d['value']), round(d['ms_per_step']), round(d['e2e']['value'])); print(json.dumps(d['e2e_files'])[:1500]); print({k:round(x) for k,x in d['kernel_ms_per_step'].items()})"
