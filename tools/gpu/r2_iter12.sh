set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or two_stream or chunked or randomised" ) > gpurun_out/r2_iter_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2_iter_pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" PSC_DEBUG_HOST=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-files > gpurun_out/r2_v_$name.json 2> gpurun_out/r2_v_$name.err
  python3 -c "
import json; d=json.load(open('gpurun_out/r2_v_$name.json')); print('$name', round(d['value']), round(d['ms_per_step']), round(d['e2e']['value']), {k:round(x) for k,x in d['kernel_ms_per_step'].items()})"
}
run x2
run x3 PSC_THREE_CHAINS=1
