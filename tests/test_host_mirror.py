"""Host-side mirror of the reference's Go helpers (annotation parsing, defaulting, config inference)."""
import pytest

from hivedscheduler_b200 import algorithm as alg
from hivedscheduler_b200 import config


def pod_with(spec):
    p = alg.Pod("p", "ns", "uid-p")
    p.annotations[alg.ANNOTATION_POD_SCHEDULING_SPEC] = alg.to_yaml(spec)
    return p


def test_spec_defaulting_matches_reference():
    # pkg/internal/utils.go:235-253: ignoreK8sSuggestedNodes defaults to true, group defaults to ns/name
    s = alg.extract_pod_scheduling_spec(pod_with({"virtualCluster": "VC1", "priority": 0, "leafCellNumber": 2}))
    assert s["ignoreK8sSuggestedNodes"] is True
    assert s["affinityGroup"] == {"name": "ns/p", "members": [{"podNumber": 1, "leafCellNumber": 2}]}
    s = alg.extract_pod_scheduling_spec(pod_with({"virtualCluster": "VC1", "leafCellNumber": 1, "ignoreK8sSuggestedNodes": False}))
    assert s["ignoreK8sSuggestedNodes"] is False


@pytest.mark.parametrize("spec", [
    {"priority": 0, "leafCellNumber": 1},                                   # VirtualCluster is empty
    {"virtualCluster": "a", "priority": -2, "leafCellNumber": 1},
    {"virtualCluster": "a", "priority": 1001, "leafCellNumber": 1},
    {"virtualCluster": "a", "priority": 0, "leafCellNumber": 0},
    {"virtualCluster": "a", "leafCellNumber": 2, "affinityGroup": {"name": "g", "members": [{"podNumber": 1, "leafCellNumber": 8}]}},
    {"virtualCluster": "a", "leafCellNumber": 2, "affinityGroup": {"name": "", "members": [{"podNumber": 1, "leafCellNumber": 2}]}},
])
def test_spec_validation_is_a_400(spec):
    with pytest.raises(alg.WebServerError) as ei:
        alg.extract_pod_scheduling_spec(pod_with(spec))
    assert ei.value.code == 400


def test_old_annotation_keys_are_converted():
    p = alg.Pod("p", "ns")
    p.annotations[alg.ANNOTATION_POD_SCHEDULING_SPEC] = "virtualCluster: a\ngpuType: K80\ngpuNumber: 2\n"
    s = alg.extract_pod_scheduling_spec(p)
    assert s["leafCellType"] == "K80" and s["leafCellNumber"] == 2


def test_binding_pod_round_trip():
    info = {"node": "n1", "leafCellIsolation": [3, 1], "cellChain": "C",
            "affinityGroupBindInfo": [{"podPlacements": [
                {"physicalNode": "n0", "physicalLeafCellIndices": [0, 2], "preassignedCellTypes": ["T", "T"]},
                {"physicalNode": "n1", "physicalLeafCellIndices": [3, 1], "preassignedCellTypes": ["T", "T"]}]}]}
    b = alg.new_binding_pod(alg.Pod("p"), info)
    assert b.node_name == "n1" and b.annotations[alg.ANNOTATION_POD_LEAF_CELL_ISOLATION] == "3,1"
    assert alg.extract_pod_bind_info(b) == info
    assert alg.get_allocated_pod_index(info, 2) == 1
    assert alg.get_allocated_pod_index(info, 4) == -1


def test_physical_cell_address_inference():
    # pkg/api/config.go:134-167 on the C1 fixture: addresses default to parent*fanout+i, reset at node level
    c = config.config_c1()
    top = c["physicalCluster"]["physicalCells"][0]
    assert top["cellAddress"] == "0"
    n0, n1 = top["cellChildren"]
    assert n0["cellAddress"] == "0/10.151.41.23" and n1["cellAddress"] == "0/10.151.41.24"
    assert [g["cellAddress"] for sw in n1["cellChildren"] for g in sw["cellChildren"]] == \
        ["0/10.151.41.24/0/0", "0/10.151.41.24/0/1", "0/10.151.41.24/1/2", "0/10.151.41.24/1/3"]
    text = config.to_spec_text(c)
    assert text.startswith("HIVEDSPEC 1\ncelltypes 3\n") and text.endswith("end\n")


def test_trace_generators_are_deterministic():
    from hivedscheduler_b200 import trace
    a, b = trace.trace_c3(n_gangs=500), trace.trace_c3(n_gangs=500)
    assert a["events"].tobytes() == b["events"].tobytes()
    assert int(a["decision"].sum()) == 500
    r1, r2 = trace.XorShift64Star(trace.seed_for(3)), trace.XorShift64Star(trace.seed_for(3))
    seq = [r1.next() for _ in range(4)]
    assert seq == [r2.next() for _ in range(4)] and len(set(seq)) == 4 and all(0 <= v < 2 ** 64 for v in seq)
    # xorshift64* known answer: one step from state 1
    r = trace.XorShift64Star(1)
    x = 1
    x ^= x >> 12
    x ^= (x << 25) & (2 ** 64 - 1)
    x ^= x >> 27
    assert r.next() == (x * 0x2545F4914F6CDD1D) % 2 ** 64


def test_user_error_messages_are_the_references(emu_lib):
    """The HTTP 400 body: the shims format the reference's own message (the names live above the ABI) —
    hived_algorithm.go:684-686, 785-787, 824-826, 858-868; cases of casesThatShouldFail (hived_algorithm_test.go)."""
    import copy
    from golden_scenario import load_fixture
    from hivedscheduler_b200 import algorithm as alg
    from hivedscheduler_b200.config import new_config
    fx = load_fixture()
    h = alg.HivedAlgorithm(new_config(copy.deepcopy(fx["design_config"])), lib=emu_lib, max_groups=256, max_pods=1024,
                           max_group_leaves=128, max_group_pods=16)
    for n in h.node_names:
        h.setHealthyNode(n)

    def sched(spec, uid):
        pod = alg.Pod(name=uid, namespace="test", uid=uid, annotations={alg.ANNOTATION_POD_SCHEDULING_SPEC: alg.to_yaml(spec)})
        return pod, h.Schedule(pod, list(h.node_names), alg.PREEMPTING_PHASE)

    base = copy.deepcopy(fx["pss"]["pod1"])
    cases = []
    s = copy.deepcopy(base); s["virtualCluster"] = "surprise!"
    cases.append((s, "[u1(test/u1)]: VC surprise! does not exists!"))
    s = copy.deepcopy(base); s["pinnedCellId"] = "surprise!"
    cases.append((s, "[u2(test/u2)]: VC %s does not have pinned cell surprise!" % base["virtualCluster"]))
    s = copy.deepcopy(base); s["leafCellType"] = "no-such-gpu"
    cases.append((s, "[u3(test/u3)]: Pod requesting leaf cell type no-such-gpu which the whole cluster does not have"))
    for i, (spec, want) in enumerate(cases):
        with pytest.raises(alg.WebServerError) as ei:
            sched(spec, "u%d" % (i + 1))
        assert ei.value.code == 400 and ei.value.message == want
    # more pods than the group declares
    spec = copy.deepcopy(base)
    pod, psr = sched(spec, "u10")
    assert psr.pod_bind_info is not None
    h.AddAllocatedPod(alg.new_binding_pod(pod, psr.pod_bind_info))
    declared = sum(m["podNumber"] for m in spec["affinityGroup"]["members"] if m["leafCellNumber"] == spec["leafCellNumber"])
    for k in range(declared - 1):
        pod, psr = sched(spec, "u1%d" % (k + 1))
        h.AddAllocatedPod(alg.new_binding_pod(pod, psr.pod_bind_info))
    with pytest.raises(alg.WebServerError) as ei:
        sched(spec, "u19")
    assert ei.value.message == ("Requesting more pods than the configured number for %d leaf cells (%d pods) in affinity group %s"
                                % (spec["leafCellNumber"], declared, spec["affinityGroup"]["name"]))
    h.close()
