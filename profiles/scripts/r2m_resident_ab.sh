set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -15 gpurun_out/r2m_pytest.log
for v in product ab_volpub ab_alignedbar ab_both; do
  if [ $v = product ]; then unset HIVED_CUDA_LIB; else export HIVED_CUDA_LIB=$PWD/tests/_build/libhived_cuda_$v.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/r2m_bench_$v.json 2> gpurun_out/r2m_bench_$v.err; echo "bench $v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], d['per_call'])"
done
unset HIVED_CUDA_LIB
HIVED_NO_RESIDENT=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/r2m_bench_noresident.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_noresident.json')); print('noresident', d['value'], d['per_call'])"
