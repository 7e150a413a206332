set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2zd_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2zd_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/r2zd_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2zd_bench.json')); print('r2zd', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity'])"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
