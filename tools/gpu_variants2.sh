#!/bin/bash
# compare kernel build variants (build/variants/<X>/libpsc.so) on a bench-like workload:
# total DP cycles, heaviest DP unit (PSC_DEBUG_DP), then throughput and kernel ms of a normal run
cd "$(dirname "$0")/.."
G=${GENES:-6000}
cp psiclass_b200/libpsc.so /tmp/libpsc.keep
for v in "$@"; do
  cp build/variants/$v/libpsc.so psiclass_b200/libpsc.so
  PSC_DEBUG_DP=1 python bench.py --genes $G --steps 1 --warmup 1 --no-cpu-baseline > /tmp/v.json 2>/tmp/v.err
  echo "$v: $(grep '^\[dp\]' /tmp/v.err | tail -13 | head -2 | cut -c1-150 | tr '\n' ' ')"
  if [ -z "$DP_ONLY" ]; then
  python bench.py --genes $G --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);k=d['kernel_ms_per_step'];print('   value',round(d['value']),'dp',round(k['dp_units'],1),'pick1',round(k['units_pick1'],1),'pick2',round(k['units_pick2_refine_em'],1))"
  fi
done
cp /tmp/libpsc.keep psiclass_b200/libpsc.so
