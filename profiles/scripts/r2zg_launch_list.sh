set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2zg_launches.csv python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline > gpurun_out/r2zg_under_ncu.log 2>&1; echo "ncu rc=$?"
grep -c hived gpurun_out/r2zg_launches.csv
