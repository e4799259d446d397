set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python3 -c "
import json; d=json.load(open('gpurun_out/r2_bench_final.json')); print(round(d['value']), round(d['ms_per_step'],1), round(d['e2e']['value'])); print(json.dumps(d['e2e_files'])[:700]); print(d['cpu_baseline'])"
timeout 300 python -m pytest tests -m gpu -x -q -k "drop_in or fused_vote or psiclass_driver or two_stream" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
