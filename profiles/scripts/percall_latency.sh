# per-call latency from plain C, resident kernel on / off  ->  gpurun_out/percall_*.jsonl
set -e
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
python -c "
from hivedscheduler_b200 import config
open('gpurun_out/c3.spec','w').write(config.to_spec_text(config.config_c3()))"
gcc -O2 -std=c99 -I include -o gpurun_out/percall profiles/micro/percall_latency.c -L hivedscheduler_b200/csrc -lhived_cuda -Wl,-rpath,$PWD/hivedscheduler_b200/csrc
./gpurun_out/percall gpurun_out/c3.spec 5000 > gpurun_out/percall_resident.jsonl
HIVED_NO_RESIDENT=1 ./gpurun_out/percall gpurun_out/c3.spec 3000 > gpurun_out/percall_launch_per_call.jsonl
cat gpurun_out/percall_resident.jsonl gpurun_out/percall_launch_per_call.jsonl
