set -x
cd "$GRAFT_REPO_ROOT"
# --- config 3 diagnosis: 2 dense components, 10 samples, --maxDpConstraintSize -1
mkdir -p /tmp/d3 && ./build/synth -o /tmp/d3 --genes 2 --samples 10 --seed 1003 --dense 1 --pairs 300 --chr 2
( cd /tmp/d3 && time PSC_DEBUG_HOST=1 timeout 170 $GRAFT_REPO_ROOT/psiclass_b200/bin/classes -s subexon_combined.out --lb bamlist -o new --maxDpConstraintSize -1 -p 8 ) > gpurun_out/r2_d3_ours.log 2>&1
grep -E "dp |units|real|atCnt" gpurun_out/r2_d3_ours.log | head -30
# --- ncu
export PSC_CRIT_SMS=0
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-files"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 45 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_units|k_dp2_units|k_build_constraints|k_dp_merge" -c 13 -o /tmp/r2_prof $CMD > gpurun_out/r2_ncu_f.log 2>&1
ls -la /tmp/r2_prof.ncu-rep
ncu -i /tmp/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2> gpurun_out/r2_prof_raw.err
ncu -i /tmp/r2_prof.ncu-rep --page source --csv --kernel-name regex:k_build_constraints > gpurun_out/r2_prof_src_build.csv 2>/dev/null
gzip -f gpurun_out/r2_prof_src_build.csv
ls -la gpurun_out/
