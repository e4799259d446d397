#!/bin/bash
# Regenerates the golden fixtures with the UNMODIFIED reference (oracle/_ref, built from
# /root/reference by oracle/Makefile).  Run in the build container only; the outputs are committed.
#   synth_small/   seeded synthetic set (tools/synth): subexon_combined.out, s0.bam, s1.bam
#                  ref_dump.txt   = library stats, components, constraint tables, totalReadCnt
#                  ref_sample_*.gtf = the reference's GTFs ('#' header lines dropped)
#   dp_small/      same, with richer alternative splicing so several components take the DP path
#   nh3_small/     20 % NH=3 reads: the weight sums follow std::sort's order of equal constraints (Q9)
#   dense_down/    one component of 1587 subexons, 13 samples, --maxDpConstraintSize -1: the down-sampling
#                  branch of TranscriptDecider.cpp:3214-3235 (n >= 1500), iterBound = 10 and its skip rule
#   example/       the reference's own example/ (BASELINE.json configs[0]): the two BAMs (data, copied), the
#                  combined subexon file its stages 0-2 produce, the stage-3 GTFs and the stage-4 vote
set -e
cd "$(dirname "$0")"
R=../../oracle/_ref
CLASSES_ARGS=""
gen() { # name, synth args...
  name=$1; shift
  rm -rf $name; mkdir -p $name
  ../../build/synth -o $name "$@" 2>/dev/null
  ( cd $name
    ls s*.bam | sort -V > bamlist
    ../$R/ref_dump -s subexon_combined.out --lb bamlist -o ref $CLASSES_ARGS > ref_dump.txt
    for f in ref_sample_*.gtf; do grep -v '^#' $f > $f.tmp || true; mv $f.tmp $f; done )
}
gen synth_small --genes 25 --samples 2 --seed 21 --pairs 40 --name-style 2 --p-noise 0.08 --p-overlap 0.1
gen dp_small --genes 14 --samples 3 --seed 22 --pairs 120 --exon-mu 2.2 --iso-lambda 4 --p-alt 0.3 --p-ir 0.1
gen nh3_small --genes 30 --samples 3 --seed 31 --pairs 300 --p-nh3 0.2 --iso-lambda 3 --p-alt 0.3
CLASSES_ARGS="--maxDpConstraintSize -1"
gen dense_down --genes 1 --samples 13 --seed 1003 --dense 1 --exon-min 460 --exon-max 470 --pairs 12
grep -v '^[cm] ' dense_down/ref_dump.txt > dense_down/ref_dump.tmp; mv dense_down/ref_dump.tmp dense_down/ref_dump.txt  # tables are covered elsewhere
CLASSES_ARGS=""
# the reference's example through its own driver, stages 0-4
rm -rf example; mkdir -p example
( cd example
  cp /root/reference/example/s1.bam /root/reference/example/s2.bam .
  printf "s1.bam\ns2.bam\n" > bamlist
  ../$R/psiclass --lb bamlist -o ex_ > /dev/null 2>&1
  cp subexon/ex_subexon_combined.out subexon_combined.out
  ../$R/ref_dump -s subexon_combined.out --lb bamlist -o ref > ref_dump.txt
  for i in 0 1; do grep -v '^#' ref_sample_$i.gtf > ref_sample_$i.gtf.tmp; mv ref_sample_$i.gtf.tmp ref_sample_$i.gtf; cmp ref_sample_$i.gtf <(grep -v '^#' ex_sample_$i.gtf); done
  grep -v '^#' ex_vote.gtf > ref_vote.gtf
  rm -rf splice subexon ex_* )
du -sh synth_small dp_small nh3_small dense_down example
