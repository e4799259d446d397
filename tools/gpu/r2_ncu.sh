set -x
cd "$GRAFT_REPO_ROOT"
export PSC_CRIT_SMS=0
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-files"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
$CMD > gpurun_out/r2_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_units|k_dp2_units|k_build_constraints|k_dp_merge" -c 13 -o /tmp/r2_prof $CMD > gpurun_out/r2_ncu_f.log 2>&1
ls -la /tmp/r2_prof.ncu-rep
ncu -i /tmp/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2> gpurun_out/r2_prof_raw.err
ncu -i /tmp/r2_prof.ncu-rep --page details --csv > gpurun_out/r2_prof_details.csv 2>/dev/null
ncu -i /tmp/r2_prof.ncu-rep --page source --csv --kernel-name regex:k_units > gpurun_out/r2_prof_src_units.csv 2>/dev/null
ncu -i /tmp/r2_prof.ncu-rep --page source --csv --kernel-name regex:k_build_constraints > gpurun_out/r2_prof_src_build.csv 2>/dev/null
ncu -i /tmp/r2_prof.ncu-rep --page source --csv --kernel-name regex:k_dp2_units > gpurun_out/r2_prof_src_dp2.csv 2>/dev/null
gzip -f gpurun_out/r2_prof_src_*.csv
sz=$(stat -c %s /tmp/r2_prof.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/r2_prof.ncu-rep gpurun_out/; fi
du -sh gpurun_out; ls -la gpurun_out/
