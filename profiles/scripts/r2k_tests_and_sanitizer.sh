set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
for tool in memcheck racecheck synccheck initcheck; do
  vcs=8; [ $tool = racecheck ] && vcs=16
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitizer_driver.py $vcs 4000 > gpurun_out/sanitizer_${tool}_${vcs}cta.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_${vcs}cta.log
  tail -6 gpurun_out/sanitizer_${tool}_${vcs}cta.log
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tests/sanitizer_driver.py 16 4000 > gpurun_out/sanitizer_memcheck_16cta.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck_16cta.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tests/sanitizer_driver.py 8 4000 > gpurun_out/sanitizer_racecheck_8cta.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_racecheck_8cta.log
tail -4 gpurun_out/sanitizer_memcheck_16cta.log gpurun_out/sanitizer_racecheck_8cta.log
