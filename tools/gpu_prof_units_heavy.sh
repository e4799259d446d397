#!/bin/bash
# ncu source-level profile of k_units (phases 1 and 2) on the heavy-DP set: few, very large units
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
d=/tmp/fzh; rm -rf $d
./build/synth -o $d --genes 30 --samples 4 --seed 61 --exon-mu 3.0 --exon-sigma 0.4 --iso-lambda 5 --p-alt 0.4 --pairs 1200 2>/dev/null
PSC_NO_OVERLAP=1 PSC_DEBUG_HOST=1 ./build/psc_check $d 2>&1 | grep -E "units phase|RESULT"
PSC_NO_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:k_units -c 2 -o gpurun_out/prof_units_heavy ./build/psc_check $d > gpurun_out/heavy_units_ncu.log 2>&1
tail -2 gpurun_out/heavy_units_ncu.log
