"""The DEVICE PROGRAM with its real SIMT geometry, on the host: 32-lane warps, a leader warp plus worker warps per
CTA and one CTA per group of VCs, executed by the functional emulator of tests/emu/simt_rt.h (every CUDA thread a
fiber; collectives and barriers block until all peers arrived; the lanes of a warp run one after the other between
two collectives).  test_device_program_emu.py checks the scheduling logic with one lane; these tests check what one
lane cannot see — ballots / match / shuffles / lane-partitioned loops, the barrier protocol with the worker warps and
the ordered shared sections between CTAs — against the oracle, in the CPU-only tier."""
import numpy as np
import pytest

from conftest import run_trace
from golden_scenario import Scenario
from hivedscheduler_b200 import trace
from test_device_program_emu import small_c3


def _same(lib, oracle, t, chunks=2):
    snaps = []
    ha, ra, sa = run_trace(lib, t, chunks=chunks, snapshots=snaps)
    hb, rb, sb = run_trace(oracle, t, chunks=chunks, snapshots=snaps)
    assert ha == hb and sa == sb
    assert snaps[0] == snaps[1]
    for (a, pa), (b, pb) in zip(ra, rb):
        assert a.tobytes() == b.tobytes()
        n = int((a["leaf_off"] + 3 * a["n_leaves"]).max()) if len(a) else 0
        assert pa[:n].tobytes() == pb[:n].tobytes()


def test_simt_reproduces_reference_golden_vectors(simt_lib, oracle_lib):
    sc = Scenario(simt_lib)
    assert sc.run() == []
    so = Scenario(oracle_lib)
    assert so.run() == []
    assert sc.decisions == so.decisions


@pytest.mark.parametrize("name", ["C1", "C2", "C3-small", "multi-member", "heterogeneous", "suggested-nodes", "bad-requests"])
def test_simt_matches_oracle_on_trace(simt_lib, oracle_lib, name):
    t = {"C1": trace.trace_c1, "C2": lambda: trace.trace_c2(n_pods=700), "C3-small": lambda: small_c3(500),
         "multi-member": lambda: trace.trace_multi_member(n_gangs=300),
         "heterogeneous": lambda: trace.trace_heterogeneous(n_gangs=500),
         "suggested-nodes": lambda: trace.trace_suggested_nodes(n_gangs=250),
         "bad-requests": trace.trace_bad_requests}[name]()
    _same(simt_lib, oracle_lib, t)
