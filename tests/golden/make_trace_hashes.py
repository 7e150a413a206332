#!/usr/bin/env python
"""Runs the CPU oracle over the full-size benchmark traces and records the parity hash
(FNV-1a over every SCHEDULE result, include/hived_hash.h) after each chunk of events in
tests/golden/trace_hashes.json.  The full C3 trace takes the oracle ~20-30 minutes, so the GPU
parity test at BASELINE size compares against these committed checkpoints instead of re-running it.

    python tests/golden/make_trace_hashes.py C3 [n_chunks]
    python tests/golden/make_trace_hashes.py C5 [n_chunks]      (full-size churn trace, ~10 min of oracle time)
    python tests/golden/make_trace_hashes.py C4 [n_gangs]       (call-by-call preemption harness, ~1 h at 100 000 gangs)
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hivedscheduler_b200 import _cabi, trace  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "trace_hashes.json")


def log_digest(log) -> str:
    """sha256 over the decision log of the interactive harness (gang, pod, kind, node / victims / wait code)."""
    m = hashlib.sha256()
    for entry in log:
        m.update(repr(entry).encode())
    return m.hexdigest()


def c4_kwargs(n_gangs: int):
    from hivedscheduler_b200.config import config_c3
    return dict(config=config_c3(), n_gangs=n_gangs, n_vcs=8, vc_gpus=7168, total_gpus=65536)


def main_c4(lib, n_gangs: int):
    t0 = time.time()
    h, log, stats = trace.run_c4_interactive(lib, **c4_kwargs(n_gangs))
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    from collections import Counter
    data["C4"] = {"n_gangs": n_gangs, "hash": "%016x" % h, "log_sha256": log_digest(log), "log_entries": len(log),
                  "kinds": dict(Counter(e[2] for e in log)), "stats": stats, "oracle_seconds": time.time() - t0}
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print("C4", data["C4"], flush=True)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    lib = _cabi.load_library(os.path.join(ROOT, "oracle", "libhived_oracle.so"))
    if name == "C4":
        return main_c4(lib, int(sys.argv[2]) if len(sys.argv) > 2 else 100000)
    t = {"C2": trace.trace_c2, "C3": trace.trace_c3, "C5": trace.trace_c5}[name]()
    ev = t["events"]
    bc = trace.BatchContext(lib, t["config"], t["n_groups"], t["n_pods"], t["max_group_leaves"], t["max_group_pods"])
    bc.set_all_nodes_healthy()
    n = len(ev)
    bounds = [n * (i + 1) // n_chunks for i in range(n_chunks)]
    checkpoints = []
    start = 0
    t0 = time.time()
    for b in bounds:
        chunk = ev[start:b]
        res, _ = bc.process(chunk, 3 * 64 * len(chunk) + 4096)
        sched = chunk["type"] == _cabi.EV_SCHEDULE
        checkpoints.append({"events": int(b), "hash": "%016x" % bc.result_hash(),
                            "binds": int((res["kind"][sched] == 1).sum()),
                            "waits": int((res["kind"][sched] == 0).sum())})
        start = b
        print(name, b, n, checkpoints[-1], "%.0fs" % (time.time() - t0), flush=True)
        data = json.load(open(OUT)) if os.path.exists(OUT) else {}
        data[name] = {"n_events": n, "n_decisions": int(t["decision"].sum()), "checkpoints": checkpoints,
                      "stats": bc.stats(), "oracle_seconds": time.time() - t0}
        json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
