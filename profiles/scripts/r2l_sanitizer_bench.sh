set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "frontend or litmus or c_driver or partition" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log
for tool in synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitizer_driver.py 8 4000 > gpurun_out/sanitizer2_${tool}_8cta.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer2_${tool}_8cta.log
  grep -v "Host Frame\|Device Frame" gpurun_out/sanitizer2_${tool}_8cta.log | tail -12
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tests/sanitizer_driver.py 16 4000 > gpurun_out/sanitizer2_racecheck_16cta.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer2_racecheck_16cta.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tests/sanitizer_driver.py 16 4000 > gpurun_out/sanitizer2_synccheck_16cta.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer2_synccheck_16cta.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python profiles/micro/ingest_bench.py > gpurun_out/r2l_ingest.json 2>&1
head -c 1500 gpurun_out/r2l_bench.json
