set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/r2s_topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2s_bench_gpus$N.json 2> gpurun_out/r2s_bench_gpus$N.err
echo "rc=$?"; tail -3 gpurun_out/r2s_bench_gpus$N.err; head -c 3000 gpurun_out/r2s_bench_gpus$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 3 --warmup 3 --replicas --no-other-configs --no-cpu-baseline > gpurun_out/r2s_bench_replicas$N.json 2> gpurun_out/r2s_bench_replicas$N.err
echo "rc=$?"; head -c 500 gpurun_out/r2s_bench_replicas$N.json
