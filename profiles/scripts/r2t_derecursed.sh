set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2t_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); print('r2t', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity'], json.dumps(d['other_configs']))"
HIVED_CUDA_LIB=$PWD/hivedscheduler_b200/csrc/libhived_cuda_profile.so timeout 300 python profiles/scripts/c5_phases.py > gpurun_out/r2t_c5_phases.json 2> gpurun_out/r2t_c5_phases.err; cat gpurun_out/r2t_c5_phases.json; tail -2 gpurun_out/r2t_c5_phases.err
