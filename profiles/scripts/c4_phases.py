#!/usr/bin/env python
"""Where C4 (preemption, call by call) goes: the profiling build's SM-cycle counters per phase while the compiled
closed-loop player (tests/harness/c4_player.cpp) drives the per-call path.
    HIVED_CUDA_LIB=hivedscheduler_b200/csrc/libhived_cuda_profile.so python profiles/scripts/c4_phases.py [n_gangs]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hivedscheduler_b200 import _cabi, trace  # noqa: E402
from hivedscheduler_b200.config import config_c3  # noqa: E402

n_gangs = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
lib = _cabi.load_cuda_library()
# the counters are per context; run_c4_compiled owns its context, so read them through a hook on close
orig_close = trace.BatchContext.close
out = {}


def close(self):
    cyc = (C.c_int64 * 16)()
    lib.hived_bench_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.hived_bench_phase_cycles(self.ctx, cyc)
    pc = (C.c_int64 * 16)()
    lib.hived_bench_path_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    n = lib.hived_bench_path_counters(self.ctx, pc)
    names = ["view", "leaf_search", "map_v2p", "emit", "commit", "delete", "all_events", "shared_wait", "shared_sections",
             "sched_existing_cyc", "sched_existing_n", "delete_pod_cyc", "delete_pod_n", "commit_pod_cyc", "commit_pod_n"]
    out["phase_cycles"] = {k: int(cyc[i]) for i, k in enumerate(names)}
    out["paths"] = [int(x) for x in list(pc)[:n]]
    orig_close(self)


trace.BatchContext.close = close
h, log, st, tm = trace.run_c4_compiled(lib, config=config_c3(), n_gangs=n_gangs, n_vcs=8, vc_gpus=7168, total_gpus=65536)
out.update({"hash": "%016x" % h, "timing": tm, "gangs_per_s": n_gangs / tm["seconds"], "us_per_call": 1e6 * tm["seconds"] / tm["calls"],
            "stats": st})
print(json.dumps(out))
