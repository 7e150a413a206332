set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2z_bench.json')); print('r2z', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity'], d['roofline']['frac'], d['roofline']['traffic']); print(json.dumps(d['other_configs']))"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | head -c 400
