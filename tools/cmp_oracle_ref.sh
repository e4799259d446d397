#!/bin/bash
# usage: cmp_oracle_ref.sh <dir with subexon_combined.out + bamlist> [extra classes args...]
# Runs the unmodified reference (ref_dump) and the CPU restatement (psc_oracle) and diffs tables + GTFs.
d=$1; shift
R=/root/repo/oracle/_ref
cd $d || exit 1
$R/ref_dump -s subexon_combined.out --lb bamlist -o ref "$@" > ref.txt 2> ref.err || { echo "ref_dump failed"; exit 2; }
$R/psc_oracle -s subexon_combined.out --lb bamlist -o orc --dump "$@" > orc.txt 2> orc.err || { echo "psc_oracle failed"; tail -3 orc.err; exit 3; }
grep -v atCnt ref.txt > ref.t; grep -v atCnt orc.txt > orc.t
grep atCnt ref.txt > ref.a; grep atCnt orc.txt > orc.a
if ! cmp -s ref.t orc.t; then echo "TABLE DIFF"; diff ref.t orc.t | head -10; fi
if ! cmp -s ref.a orc.a; then echo "ATCNT DIFF"; diff ref.a orc.a | head -10; fi
n=$(wc -l < bamlist); bad=0
for ((i=0;i<n;i++)); do
  if ! diff -q <(grep -v '^#' ref_sample_$i.gtf) <(grep -v '^#' orc_sample_$i.gtf) >/dev/null; then bad=$((bad+1)); echo "GTF DIFF sample $i"; diff <(grep -v '^#' ref_sample_$i.gtf) <(grep -v '^#' orc_sample_$i.gtf) | head -6; fi
done
echo "done: $(grep -c . ref.a) comps, useDP=$(awk '$4==1' ref.a | wc -l), gtf lines $(cat ref_sample_*.gtf | wc -l), bad=$bad"
