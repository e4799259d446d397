set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L | head -8; nproc
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -3 gpurun_out/r2_bench_8gpu.err
( time timeout 420 python bench.py --config 2 --gpus 8 ) > gpurun_out/r2_config2_8gpu.json 2> gpurun_out/r2_config2_8gpu.err
python3 -c "
import json
for f in ('gpurun_out/r2_bench_8gpu.json','gpurun_out/r2_config2_8gpu.json'):
    for ln in open(f):
        if ln.startswith('{'):
            d=json.loads(ln); print(f, d.get('value'), d.get('ms_per_step'), json.dumps(d.get('e2e_files') or d.get('drop_in'))[:1800])"
