set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2za_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2za_pytest.log
FUZZ_FILTERING=1 timeout 600 python tests/fuzz_api.py cuda 1 12 300 > gpurun_out/r2za_apifuzz_filtering.log 2>&1; tail -2 gpurun_out/r2za_apifuzz_filtering.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/r2za_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2za_bench.json')); print('r2za', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity'])"
python -c "import __graft_entry__ as g; g.smoke()"
