#!/bin/bash
# Regenerates the golden fixtures with the UNMODIFIED reference (oracle/_ref, built from
# /root/reference by oracle/Makefile).  Run in the build container only; the outputs are committed.
#   synth_small/   seeded synthetic set (tools/synth): subexon_combined.out, s0.bam, s1.bam
#                  ref_dump.txt   = library stats, components, constraint tables, totalReadCnt
#                  ref_sample_*.gtf = the reference's GTFs ('#' header lines dropped)
#   dp_small/      same, with richer alternative splicing so several components take the DP path
set -e
cd "$(dirname "$0")"
R=../../oracle/_ref
gen() { # name, synth args...
  name=$1; shift
  rm -rf $name; mkdir -p $name
  ../../build/synth -o $name "$@" 2>/dev/null
  ( cd $name
    ls s*.bam | sort -V > bamlist
    ../$R/ref_dump -s subexon_combined.out --lb bamlist -o ref > ref_dump.txt
    for f in ref_sample_*.gtf; do grep -v '^#' $f > $f.tmp; mv $f.tmp $f; done )
}
gen synth_small --genes 25 --samples 2 --seed 21 --pairs 40 --name-style 2 --p-noise 0.08 --p-overlap 0.1
gen dp_small --genes 14 --samples 3 --seed 22 --pairs 120 --exon-mu 2.2 --iso-lambda 4 --p-alt 0.3 --p-ir 0.1
du -sh synth_small dp_small
