#!/usr/bin/env python
"""Seeded fuzz of the parity property: the same trace families as the tests, re-seeded, replayed on two implementations
of include/hived.h and compared byte-wise (results, pool, work counters, every cell's state).

    python tests/fuzz_parity.py emu  [first_seed n_seeds]     device program (host emulation) vs oracle, CPU only
    python tests/fuzz_parity.py cuda [first_seed n_seeds]     libhived_cuda.so vs oracle, on a GPU box
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from conftest import run_trace  # noqa: E402
from hivedscheduler_b200 import _cabi, config, trace  # noqa: E402


def families():
    small = config.config_c3(n_pods=4, n_vcs=2, racks_per_vc=6)
    return [
        ("multi-member", lambda: trace.trace_multi_member(n_gangs=500)),
        ("heterogeneous", lambda: trace.trace_heterogeneous(n_gangs=700)),
        ("suggested-nodes", lambda: trace.trace_suggested_nodes(n_gangs=300)),
        ("churn", lambda: trace.trace_c5(n_steps=3, gangs_per_step=150, n_nodes=4 * 16 * 32, n_vcs=2,
                                         vc_gpus=(16 + 6) * 32 * 8, config=small)),
        ("c3-small", lambda: dict(trace.trace_c3(n_gangs=600, n_vcs=2, vc_gpus=(16 + 6) * 32 * 8), config=small)),
    ]


def run_seeds(lib, oracle, first: int, count: int, verbose: bool = True) -> int:
    """Replays every family for seeds first .. first+count-1 on both libraries; returns the number of divergences."""
    base = trace.seed_for
    bad = 0
    t0 = time.time()
    try:
        for seed in range(first, first + count):
            trace.seed_for = lambda n, s=seed: base(n) ^ ((0x9E3779B97F4A7C15 * s) & 0xFFFFFFFFFFFFFFFF)
            for name, gen in families():
                t = gen()
                snaps = []
                ha, ra, sa = run_trace(lib, t, chunks=2, snapshots=snaps)
                hb, rb, sb = run_trace(oracle, t, chunks=2, snapshots=snaps)
                same = (ha == hb and sa == sb and snaps[0] == snaps[1] and
                        all(x.tobytes() == y.tobytes() for (x, _), (y, _) in zip(ra, rb)))
                if not same:
                    bad += 1
                    print("DIVERGED seed %d family %s (hash %s stats %s cells %s)" % (seed, name, ha == hb, sa == sb, snaps[0] == snaps[1]),
                          flush=True)
            if verbose:
                print("seed %d done, %d divergences so far, %.0f s" % (seed, bad, time.time() - t0), flush=True)
    finally:
        trace.seed_for = base
    return bad


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "emu"
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    count = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    import __graft_entry__ as g
    oracle = _cabi.load_library(g.build_oracle())
    if which == "cuda":
        lib = _cabi.load_cuda_library()
    else:
        import subprocess
        emu_path = os.path.join(HERE, "_build", "libhived_emu.so")
        os.makedirs(os.path.dirname(emu_path), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-o", emu_path,
                               os.path.join(HERE, "emu", "hived_emu.cpp")])
        lib = _cabi.load_library(emu_path)
    bad = run_seeds(lib, oracle, first, count)
    print("fuzz %s vs oracle: %d seeds x %d families, %d divergences" % (which, count, len(families()), bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
