set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L; nproc
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2_pytest_final.log 2>&1
grep -E "passed|failed" gpurun_out/r2_pytest_final.log
( time timeout 1500 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python3 -c "
import json; d=json.load(open('gpurun_out/r2_bench_final.json')); print(round(d['value']), round(d['ms_per_step'],1), round(d['e2e']['value'])); print(json.dumps(d['e2e_files'])[:900]); print(d['roofline']); print(d['roofline_streaming']); print(d['cpu_baseline'])"
python -c "import __graft_entry__ as g; g.smoke()"
