set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2w_pytest.log
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print('r2w', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity'], json.dumps(d['other_configs']))"
timeout 300 python - <<'PY'
import json, sys, time
sys.path.insert(0, '.')
from hivedscheduler_b200 import _cabi, trace
lib = _cabi.load_cuda_library()
for k in range(3):
    t = trace.trace_c2()
    bc = trace.BatchContext(lib, t["config"], t["n_groups"], t["n_pods"], t["max_group_leaves"], t["max_group_pods"])
    bc.set_all_nodes_healthy()
    t0 = time.perf_counter(); bc.process(t["events"], 3 * 8 * len(t["events"]) + 4096); dt = time.perf_counter() - t0
    print("C2 again", k, len(t["events"]) / dt, dt, "%016x" % bc.result_hash())
    bc.close()
PY
python -c "import __graft_entry__ as g; g.smoke()"
