#!/bin/bash
# run on the GPU box: generate the fuzz sets and compare the CUDA engine with the oracle
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/gpu.txt
i=0
while read -r cfg; do
  i=$((i+1)); d=/tmp/fz$i; rm -rf $d
  ./build/synth -o $d $cfg 2>/dev/null
  for w in 32 1; do
    echo "== fz$i coop=$w: $cfg"
    timeout 600 ./build/psc_check $d --coop $w 2>&1 | tail -8
  done
done <<CFG | tee gpurun_out/check.log
--genes 40 --samples 3 --seed 7
--genes 300 --samples 5 --seed 11 --name-style 2 --p-noise 0.1 --p-overlap 0.15
--genes 200 --samples 12 --seed 12 --pairs 60 --iso-lambda 3 --p-alt 0.3
--genes 150 --samples 4 --seed 13 --pairs 300 --exon-mu 2.3 --iso-lambda 4 --p-alt 0.3 --p-ir 0.1
--genes 100 --samples 3 --seed 14 --single-end 1
--genes 100 --samples 6 --seed 15 --p-paralog 0.15 --stranded 1
CFG
