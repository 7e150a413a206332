set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); print('r2r', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity'], json.dumps(d['other_configs']))"
HIVED_NCTA=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline > gpurun_out/r2r_bench_1cta.json 2>/dev/null; echo "bench 1cta rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_1cta.json')); print('1cta', d['value'], d['e2e']['value'])"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2r_pytest.log
bash profiles/scripts/percall_latency.sh > gpurun_out/r2r_percall.log 2>&1; tail -9 gpurun_out/r2r_percall.log
