set -x
cd "$GRAFT_REPO_ROOT"
export PSC_CRIT_SMS=0
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-files"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
$CMD > gpurun_out/r2_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_units|k_dp2_units|k_build_constraints|k_dp_merge" -c 13 -o gpurun_out/r2_prof $CMD > gpurun_out/r2_ncu_f.log 2>&1
tail -3 gpurun_out/r2_ncu_f.log; ls -la gpurun_out/
