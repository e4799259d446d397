set -x
cd "$GRAFT_REPO_ROOT"
for k in 4 2 3; do
( time timeout 1200 python bench.py --config $k --gpus 1 --with-reference ) > gpurun_out/r2_config$k.json 2> gpurun_out/r2_config$k.err
head -c 1800 gpurun_out/r2_config$k.json; tail -2 gpurun_out/r2_config$k.err
done
