#!/usr/bin/env python
"""Where a C5 (churn) pass goes: the profiling build's SM-cycle counters per phase + path counters.
    HIVED_CUDA_LIB=hivedscheduler_b200/csrc/libhived_cuda_profile.so python profiles/scripts/c5_phases.py"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hivedscheduler_b200 import _cabi, trace  # noqa: E402

lib = _cabi.load_cuda_library()
lib.hived_bench_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
lib.hived_bench_path_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
t = trace.trace_c5()
bc = trace.BatchContext(lib, t["config"], t["n_groups"], t["n_pods"], t["max_group_leaves"], t["max_group_pods"])
bc.set_all_nodes_healthy()
cyc0 = (C.c_int64 * 16)()
lib.hived_bench_phase_cycles(bc.ctx, cyc0)
t0 = time.perf_counter()
bc.process(t["events"], 3 * 64 * len(t["events"]) + 4096)
dt = time.perf_counter() - t0
cyc = (C.c_int64 * 16)()
lib.hived_bench_phase_cycles(bc.ctx, cyc)
pc = (C.c_int64 * 16)()
n = lib.hived_bench_path_counters(bc.ctx, pc)
names = ["view", "leaf_search", "map_v2p", "emit", "commit", "delete", "all_events", "shared_wait", "shared_sections",
         "sched_existing_cyc", "sched_existing_n", "delete_pod_cyc", "delete_pod_n", "commit_pod_cyc", "commit_pod_n"]
ev = t["events"]
print(json.dumps({"seconds_e2e": dt, "gangs_per_s": int(t["decision"].sum()) / dt,
                  "events": {"health": int((ev["type"] == 3).sum()), "schedule": int((ev["type"] == 0).sum()), "delete": int((ev["type"] == 1).sum())},
                  "phase_cycles": {k: int(cyc[i] - cyc0[i]) for i, k in enumerate(names)},
                  "paths": [int(x) for x in list(pc)[:n]], "hash": "%016x" % bc.result_hash()}))
