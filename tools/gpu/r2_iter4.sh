set -x
cd "$GRAFT_REPO_ROOT"
run() { # name, env...
  name=$1; shift
  env "$@" PSC_DEBUG_HOST=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-files > gpurun_out/r2_v_$name.json 2> gpurun_out/r2_v_$name.err
  python3 -c "
import json; d=json.load(open('gpurun_out/r2_v_$name.json')); print('$name', round(d['value']), round(d['ms_per_step']), round(d['e2e']['value']), {k:round(x) for k,x in d['kernel_ms_per_step'].items()})"
}
run c24h9 PSC_CRIT_SMS=24
run c24h20 PSC_CRIT_SMS=24 PSC_HEAVY_COMPS=20
run c24h40 PSC_CRIT_SMS=24 PSC_HEAVY_COMPS=40
run c32h30 PSC_CRIT_SMS=32 PSC_HEAVY_COMPS=30
run c32h60 PSC_CRIT_SMS=32 PSC_HEAVY_COMPS=60
run c40h60 PSC_CRIT_SMS=40 PSC_HEAVY_COMPS=60
