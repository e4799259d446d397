set -x
cd "$GRAFT_REPO_ROOT"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/r2_v_$name.json 2> gpurun_out/r2_v_$name.err
  python3 -c "
import json; d=json.load(open('gpurun_out/r2_v_$name.json')); print('$name', round(d['value']), round(d['ms_per_step']), round(d['e2e']['value']), round(d['device_ms_per_step']))"
}
run q1
run q2 PSC_DEBUG_HOST=1
run q3
run q4 PSC_CRIT_SMS=0
