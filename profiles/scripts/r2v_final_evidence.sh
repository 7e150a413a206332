set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2v_pytest.log
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2v_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference.json 2>/dev/null; head -c 600 gpurun_out/r2v_bench_reference.json; echo
HIVED_CUDA_LIB=$PWD/hivedscheduler_b200/csrc/libhived_cuda_profile.so timeout 300 python profiles/scripts/c5_phases.py > gpurun_out/r2v_c5_phases.json 2> gpurun_out/r2v_c5_phases.err; cat gpurun_out/r2v_c5_phases.json
HIVED_CUDA_LIB=$PWD/hivedscheduler_b200/csrc/libhived_cuda_profile.so timeout 300 python profiles/scripts/c4_phases.py > gpurun_out/r2v_c4_phases.json 2> gpurun_out/r2v_c4_phases.err; cat gpurun_out/r2v_c4_phases.json; tail -2 gpurun_out/r2v_c4_phases.err
HIVED_CUDA_LIB=$PWD/hivedscheduler_b200/csrc/libhived_cuda_profile.so timeout 600 python bench.py --steps 3 --warmup 2 --no-other-configs --no-cpu-baseline > gpurun_out/r2v_bench_profile.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2v_bench_profile.json')); print(d['value'], json.dumps(d.get('phase_cycles_per_step')))"
# launch list of the bench command (per-launch gpu time, cold and serialised) and one full capture of the C3 launch
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline > gpurun_out/r2v_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:hived_events_kernel -s 2 -c 1 -o gpurun_out/ncu_r2v python bench.py --steps 1 --warmup 1 --no-other-configs --no-cpu-baseline > gpurun_out/r2v_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/ncu_r2v.ncu-rep
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitizer_driver.py 8 3000 > gpurun_out/sanitizer3_${tool}_8cta.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer3_${tool}_8cta.log
  grep -v "Host Frame\|Device Frame" gpurun_out/sanitizer3_${tool}_8cta.log | tail -6
done
bash profiles/scripts/percall_latency.sh > gpurun_out/r2v_percall.log 2>&1; tail -8 gpurun_out/r2v_percall.log
python profiles/micro/ingest_bench.py > gpurun_out/r2v_ingest.json 2>&1; tail -3 gpurun_out/r2v_ingest.json
