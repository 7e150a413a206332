"""Parity tests proper: the CUDA path (libhived_cuda.so, through the C ABI) against the oracle on the
same seeded inputs, against the reference's golden vectors, and — at BASELINE size — against the
committed oracle checkpoints plus size-independent properties.  Bit-exact: all work is integer."""
import json
import os

import numpy as np
import pytest

from conftest import run_trace
from golden_scenario import Scenario
from hivedscheduler_b200 import _cabi, trace
from test_device_program_emu import small_c3

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_cuda_backend_is_the_one_loaded(cuda_lib):
    assert cuda_lib.hived_backend() == b"cuda-sm100a"


def test_cuda_reproduces_reference_golden_vectors(cuda_lib, oracle_lib):
    sc = Scenario(cuda_lib)
    assert sc.run() == []
    so = Scenario(oracle_lib)
    assert so.run() == []
    assert sc.decisions == so.decisions


@pytest.mark.parametrize("name", ["C1", "C2", "C3-small"])
def test_cuda_matches_oracle_on_trace(cuda_lib, oracle_lib, name):
    t = {"C1": trace.trace_c1, "C2": trace.trace_c2, "C3-small": small_c3}[name]()
    hc, rc, sc = run_trace(cuda_lib, t, chunks=2)
    ho, ro, so = run_trace(oracle_lib, t, chunks=2)
    assert hc == ho
    assert sc == so
    for (a, pa), (b, pb) in zip(rc, ro):
        assert a.tobytes() == b.tobytes()
        n = int((a["leaf_off"] + 3 * a["n_leaves"]).max()) if len(a) else 0
        assert pa[:n].tobytes() == pb[:n].tobytes()


def _replay_occupancy(t, results):
    """Size-independent properties of a bind-only trace: every GPU is held by at most one alive gang,
    a pod's GPUs sit on one node, and later pods of a gang repeat the gang's stored placement."""
    ev = t["events"]
    owner = {}
    placement = {}
    k = 0
    for res, pool in results:
        for i in range(len(res)):
            e = ev[k]
            k += 1
            g = int(e["spec"]["group"])
            if e["type"] == _cabi.EV_SCHEDULE:
                r = res[i]
                assert r["error"] == 0
                if r["kind"] != _cabi.KIND_BIND:
                    continue
                leaves = pool[r["leaf_off"]:r["leaf_off"] + 3 * r["n_leaves"]].reshape(-1, 3)
                mine = pool[r["this_off"]:r["this_off"] + 3 * r["this_n"]].reshape(-1, 3)
                assert (mine[:, 0] == r["node"]).all()
                key = [(int(a), int(b)) for a, b, _ in leaves]
                if g in placement:
                    assert placement[g][0] == key
                    placement[g][1] += 1
                else:
                    for gpu in key:
                        assert gpu not in owner, "GPU %r double-allocated" % (gpu,)
                        owner[gpu] = g
                    placement[g] = [key, 1]
            elif e["type"] == _cabi.EV_DELETE_ALLOCATED and g in placement:
                placement[g][1] -= 1
                if placement[g][1] == 0:
                    for gpu in placement[g][0]:
                        del owner[gpu]
                    del placement[g]
    return len(owner)


def test_cuda_full_size_c3_against_oracle_checkpoints(cuda_lib):
    """BASELINE size: 64k-GPU tree, 100k gangs.  The oracle needs ~20 minutes for this trace, so its
    parity hashes per chunk are committed in tests/golden/trace_hashes.json (make_trace_hashes.py)."""
    t = trace.trace_c3()
    golden = json.load(open(os.path.join(HERE, "golden", "trace_hashes.json")))["C3"]
    assert golden["n_events"] == len(t["events"])
    bc = trace.BatchContext(cuda_lib, t["config"], t["n_groups"], t["n_pods"], t["max_group_leaves"], t["max_group_pods"])
    bc.set_all_nodes_healthy()
    start, results = 0, []
    for cp in golden["checkpoints"]:
        chunk = t["events"][start:cp["events"]]
        res, pool = bc.process(chunk, 3 * 64 * len(chunk) + 4096)
        sched = chunk["type"] == _cabi.EV_SCHEDULE
        assert "%016x" % bc.result_hash() == cp["hash"], "diverged from the oracle before event %d" % cp["events"]
        assert int((res["kind"][sched] == 1).sum()) == cp["binds"]
        results.append((res, pool))
        start = cp["events"]
    stats = bc.stats()
    for key in ("view_nodes_scanned", "leaves_committed", "free_cells_scanned", "pods_placed", "algorithmic_bytes"):
        assert stats[key] == golden["stats"][key]
    in_use = _replay_occupancy(t, results)
    assert 0 < in_use <= 8 * 7168
    bc.close()


def test_schedule_is_idempotent_without_commit(cuda_lib):
    """Schedule alone does not allocate (hived_algorithm.go:180-224): asking twice gives the same answer."""
    from hivedscheduler_b200 import algorithm as alg
    from hivedscheduler_b200.config import config_c1
    h = alg.HivedAlgorithm(config_c1(), lib=cuda_lib, max_groups=16, max_pods=16, max_group_leaves=8, max_group_pods=8)
    for n in h.node_names:
        h.setHealthyNode(n)
    pod = alg.Pod("p0", "ns")
    pod.annotations[alg.ANNOTATION_POD_SCHEDULING_SPEC] = alg.to_yaml(
        {"virtualCluster": "default", "priority": 0, "leafCellType": "K80", "leafCellNumber": 2})
    a = h.Schedule(pod, h.node_names, alg.FILTERING_PHASE).pod_bind_info
    b = h.Schedule(pod, h.node_names, alg.FILTERING_PHASE).pod_bind_info
    assert a == b and a["leafCellIsolation"] == [0, 1]
    h.close()
