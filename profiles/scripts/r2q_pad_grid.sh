set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pad in 0 1 2; do
  HIVED_PAD_GRID=$pad timeout 600 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/r2q_bench_pad$pad.json 2> gpurun_out/r2q_bench_pad$pad.err; echo "bench pad=$pad rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_pad$pad.json')); print('pad$pad', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity'])"
done
for pad in 0 2; do
  HIVED_NCTA=1 HIVED_PAD_GRID=$pad timeout 600 python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline > gpurun_out/r2q_bench_1cta_pad$pad.json 2>/dev/null; echo "bench 1cta pad=$pad rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_1cta_pad$pad.json')); print('1cta pad$pad', d['value'], d['e2e']['value'])"
done
for pad in 0 2; do
HIVED_PAD_GRID=$pad timeout 900 python - > gpurun_out/r2q_other_pad$pad.json 2> gpurun_out/r2q_other_pad$pad.err <<'PY'
import json, sys
sys.path.insert(0, '.')
import bench
from hivedscheduler_b200 import _cabi
print(json.dumps(bench.other_configs(_cabi.load_cuda_library())))
PY
echo "other pad=$pad rc=$?"; cat gpurun_out/r2q_other_pad$pad.json; tail -3 gpurun_out/r2q_other_pad$pad.err
done
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "api_fuzz or compiled or golden or c5" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest.log
bash profiles/scripts/percall_latency.sh > gpurun_out/r2q_percall.log 2>&1; tail -12 gpurun_out/r2q_percall.log
