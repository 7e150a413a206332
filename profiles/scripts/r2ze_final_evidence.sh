set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# memcheck on the API-fuzz seeds behind the last three fixes (duplicate free-list entries, twice-deleted group, ghost reached through a reservation)
( FUZZ_FILTERING=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python tests/fuzz_api.py cuda 782 1 300
  FUZZ_FILTERING=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python tests/fuzz_api.py cuda 1059 1 300
  timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python tests/fuzz_api.py cuda 2385 1 300 ) > gpurun_out/sanitizer5_memcheck_apifuzz.log 2>&1
grep -v "Host Frame\|Device Frame" gpurun_out/sanitizer5_memcheck_apifuzz.log | grep "ERROR SUMMARY\|api fuzz"
# the default bench line (every leg) and the reference arm, as the driver runs them
timeout 600 python bench.py > gpurun_out/r2ze_bench.json 2> gpurun_out/r2ze_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2ze_bench.json')); print('r2ze', d['value'], d['e2e']['value'], d['per_call']['us_per_event'], d['parity'], d['roofline']['frac']); print(json.dumps(d['other_configs'])[:1500]); print(json.dumps(d.get('cpu_flat'))[:400])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | head -c 600
