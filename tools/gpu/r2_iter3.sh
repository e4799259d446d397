set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or two_stream or chunked or workload_shape or randomised" ) > gpurun_out/r2_iter_pytest.log 2>&1
tail -3 gpurun_out/r2_iter_pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" PSC_DEBUG_HOST=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-files > gpurun_out/r2_v_$name.json 2> gpurun_out/r2_v_$name.err
  python3 -c "
import json; d=json.load(open('gpurun_out/r2_v_$name.json')); print('$name', round(d['value']), round(d['ms_per_step']), {k:round(x) for k,x in d['kernel_ms_per_step'].items()})"
}
run c16 PSC_CRIT_SMS=16
run c24 PSC_CRIT_SMS=24
run c24u6 PSC_CRIT_SMS=24 PSC_LIBPSC=$PWD/build/libpsc_u6.so
run c24u8 PSC_CRIT_SMS=24 PSC_LIBPSC=$PWD/build/libpsc_u8.so
run c0u8 PSC_CRIT_SMS=0 PSC_LIBPSC=$PWD/build/libpsc_u8.so
