set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/r2zh_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2zh_pytest.log
timeout 160 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2zh_bench.json 2> gpurun_out/r2zh_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2zh_bench.json')); print('r2zh', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity']); oc=d['other_configs']; print({k:{kk:vv for kk,vv in oc[k].items() if kk in ('decisions_per_s','gangs_per_s','seconds_e2e','seconds','matches_oracle','us_per_call','seconds_e2e_first_pass')} for k in ('C2','C5','C4')})"
