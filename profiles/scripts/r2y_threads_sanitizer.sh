set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2y_pytest.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tests/sanitizer_driver.py 8 3000 > gpurun_out/sanitizer4_memcheck_8cta.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer4_memcheck_8cta.log; grep -v "Host Frame\|Device Frame" gpurun_out/sanitizer4_memcheck_8cta.log | tail -5
timeout 1200 compute-sanitizer --tool memcheck --print-limit 10 python tests/fuzz_api.py cuda 16 3 300 > gpurun_out/sanitizer4_memcheck_apifuzz.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer4_memcheck_apifuzz.log; grep -v "Host Frame\|Device Frame" gpurun_out/sanitizer4_memcheck_apifuzz.log | tail -5
timeout 600 python tests/fuzz_api.py cuda 100 30 300 > gpurun_out/r2y_apifuzz_cuda.log 2>&1; tail -2 gpurun_out/r2y_apifuzz_cuda.log
timeout 600 python tests/fuzz_parity.py cuda 4001 6 > gpurun_out/r2y_fuzzparity_cuda.log 2>&1; tail -1 gpurun_out/r2y_fuzzparity_cuda.log
