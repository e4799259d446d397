#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
d=/tmp/fzh; rm -rf $d
./build/synth -o $d --genes 30 --samples 4 --seed 61 --exon-mu 3.0 --exon-sigma 0.4 --iso-lambda 5 --p-alt 0.4 --pairs 1200 2>/dev/null
PSC_DEBUG_DP=1 ./build/psc_check $d > gpurun_out/heavy_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_dp_units -c 1 -o gpurun_out/prof_dp ./build/psc_check $d > gpurun_out/heavy_ncu.log 2>&1
tail -12 gpurun_out/heavy_plain.log
