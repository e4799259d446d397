#!/bin/bash
# compare kernel build variants (build/variants/<X>/libpsc.so) on the heavy-DP set; prints device ms per variant
cd "$(dirname "$0")/.."
d=/tmp/fzh; rm -rf $d
./build/synth -o $d --genes 30 --samples 4 --seed 61 --exon-mu 3.0 --exon-sigma 0.4 --iso-lambda 5 --p-alt 0.4 --pairs 1200 2>/dev/null
for v in "$@"; do
  for rep in 1 2 3; do
    echo -n "$v: "; LD_LIBRARY_PATH=build/variants/$v ./build/psc_check $d 2>&1 | grep -E "device ms|RESULT" | tr '\n' ' '; echo
  done
done
