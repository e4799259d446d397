set -x
cd "$GRAFT_REPO_ROOT"
run() { # name, env...
  name=$1; shift
  env "$@" PSC_DEBUG_HOST=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/r2_v_$name.json 2> gpurun_out/r2_v_$name.err
  python3 -c "
import json; d=json.load(open('gpurun_out/r2_v_$name.json')); print('$name', round(d['value']), round(d['ms_per_step']), round(d['e2e']['value']), {k:round(x) for k,x in d['kernel_ms_per_step'].items()})"
}
run m32 CUDA_DEVICE_MAX_CONNECTIONS=32
run m32c0 CUDA_DEVICE_MAX_CONNECTIONS=32 PSC_CRIT_SMS=0
