set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
HIVED_STACK_BYTES=16384 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or c1 or preemption_c4" > gpurun_out/r2u_pytest_stack16k.log 2>&1; echo "stack16k rc=$?"; tail -3 gpurun_out/r2u_pytest_stack16k.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden" > gpurun_out/r2u_pytest_default.log 2>&1; echo "default rc=$?"; tail -3 gpurun_out/r2u_pytest_default.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden" > gpurun_out/r2u_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -A25 "Invalid\|========= Error" gpurun_out/r2u_memcheck.log | head -60
