set -x
cd "$GRAFT_REPO_ROOT"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/r2_v_$name.json 2> gpurun_out/r2_v_$name.err
  python3 -c "
import json; d=json.load(open('gpurun_out/r2_v_$name.json')); print('$name', round(d['value']), round(d['ms_per_step']), round(d['e2e']['value']), round(d['device_ms_per_step']))"
}
run s1
run s2 CUDA_DEVICE_MAX_CONNECTIONS=32
