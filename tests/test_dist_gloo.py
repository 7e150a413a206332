"""world_size-2 gloo test of the N>1 plumbing used by bench.py (barrier + max-over-ranks timing)."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from hivedscheduler_b200 import dist as hd
rank, world, local = hd.dist_env()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["MASTER_PORT"], rank=rank, world_size=world)
dist.barrier()
mx = hd.max_over_ranks([1.0 + rank, 5.0 - rank])
assert mx == [2.0, 5.0], mx
v = hd.aggregate_throughput(100000, 3, mx[0], world)
assert abs(v - 2 * 100000 * 3 / 2.0) < 1e-9
assert [hd.vc_owner(v, 2) for v in range(4)] == [0, 1, 0, 1]
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
''' % ROOT


def test_max_over_ranks_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
