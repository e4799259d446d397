set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
( time timeout 900 python -m pytest tests -m gpu -x -q -k "drop_in or fused_vote or psiclass_driver" ) > gpurun_out/r2_pytest_2gpu.log 2>&1
grep -E "passed|failed" gpurun_out/r2_pytest_2gpu.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -3 gpurun_out/r2_bench_2gpu.err
python3 -c "
import json
for ln in open('gpurun_out/r2_bench_2gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(round(d['value']), round(d['ms_per_step']), round(d['e2e']['value'])); print(json.dumps(d.get('e2e_files'))[:2000])"
