"""Inspect surface of internal.SchedulerAlgorithm (hived_algorithm.go:298-363): GetAffinityGroup(s) and the cluster
status forests, materialised by the host mirror from the library's raw snapshots.  CPU tier: the device program
(host emulation) against the oracle after the reference's own "casesThatShouldSucceed" sequence; GPU tier: the CUDA
library against the oracle."""
import json

import pytest

from golden_scenario import Scenario
from hivedscheduler_b200 import algorithm as alg
from hivedscheduler_b200.config import new_config


def _state_after_success_cases(lib):
    sc = Scenario(lib)
    h = sc.new_algorithm(new_config(sc.raw_config()))
    sc.all_nodes = list(h.node_names)
    sc.set_healthy_nodes(h)
    sc.test_cases_that_should_succeed(h)
    assert sc.errors == []
    return sc, h


def _check_inspect(lib, oracle_lib):
    sc, h = _state_after_success_cases(lib)
    so, ho = _state_after_success_cases(oracle_lib)
    groups = h.GetAllAffinityGroups()
    assert json.dumps(groups, sort_keys=True) == json.dumps(ho.GetAllAffinityGroups(), sort_keys=True)
    assert json.dumps(h.GetClusterStatus(), sort_keys=True) == json.dumps(ho.GetClusterStatus(), sort_keys=True)
    # every bound pod's decision is what its group's status reports (nodeToLeafCellIndices, types.go:223-237)
    by_name = {g["metadata"]["name"]: g["status"] for g in groups["items"]}
    assert by_name, "the scenario allocates groups"
    for pod in sc.allocated_pods:
        info = alg.extract_pod_bind_info(pod)
        spec = alg.extract_pod_scheduling_spec(pod)
        st = by_name[spec["affinityGroup"]["name"]]
        assert st["state"] in ("Allocated", "BeingPreempted")
        assert st["vc"] == spec["virtualCluster"] and st["priority"] == spec["priority"]
        assert pod.uid in st["allocatedPods"]
        assert set(info["leafCellIsolation"]) <= set(st["physicalPlacement"][info["node"]])
        assert h.GetAffinityGroup(spec["affinityGroup"]["name"])["status"] == st
    for pod in sc.preempting_pods:
        spec = alg.extract_pod_scheduling_spec(pod)
        st = by_name[spec["affinityGroup"]["name"]]
        assert st["state"] == "Preempting" and pod.uid in st["preemptingPods"]
    with pytest.raises(alg.WebServerError) as ei:
        h.GetAffinityGroup("no-such-group")
    assert ei.value.code == 400
    with pytest.raises(alg.WebServerError):
        h.GetVirtualClusterStatus("no-such-vc")
    # forests: every cell once; a Used leaf is reported Used in both trees with the same priority
    phys = h.GetPhysicalClusterStatus()

    def walk(c):
        yield c
        for ch in c.get("cellChildren", []):
            yield from walk(ch)
    cells = [c for top in phys for c in walk(top)]
    assert len(cells) == len(h.physical_snapshot())
    used = [c for c in cells if c["cellState"] == "Used" and "cellChildren" not in c]
    assert used
    for c in used:
        if "virtualCell" in c:
            assert c["virtualCell"]["cellState"] == "Used" and c["virtualCell"]["cellPriority"] == c["cellPriority"]
            assert c["vc"] in h.vc_names
    vcs = h.GetAllVirtualClustersStatus()
    assert set(vcs) == set(h.vc_names)
    nvirt = sum(1 for tops in vcs.values() for top in tops for _ in walk(top))
    assert nvirt == len(h.virtual_snapshot())
    assert h.GetVirtualClusterStatus(h.vc_names[0]) == vcs[h.vc_names[0]]
    h.close()
    ho.close()


def test_inspect_device_program_matches_oracle(emu_lib, oracle_lib):
    _check_inspect(emu_lib, oracle_lib)


@pytest.mark.gpu
def test_inspect_cuda_matches_oracle(cuda_lib, oracle_lib):
    _check_inspect(cuda_lib, oracle_lib)


# ---- recovery: a pod whose cells left the spec is "insisted on" from the other pods' annotations
# (generateAffinityGroupBindInfo + retrieveMissingPodPlacement, pkg/algorithm/utils.go:126-141, 250-265)
def _raw_c1(node_a="10.151.41.23", node_b="10.151.41.24"):
    return {
        "physicalCluster": {
            "cellTypes": {
                "K80-2GPU": {"childCellType": "K80", "childCellNumber": 2},
                "K80-NODE": {"childCellType": "K80-2GPU", "childCellNumber": 2, "isNodeLevel": True},
                "2-K80-NODE": {"childCellType": "K80-NODE", "childCellNumber": 2},
            },
            "physicalCells": [{"cellType": "2-K80-NODE", "cellChildren": [{"cellAddress": node_a}, {"cellAddress": node_b}]}],
        },
        "virtualClusters": {"default": {"virtualCells": [{"cellType": "2-K80-NODE", "cellNumber": 1}]}},
    }


def _check_missing_placement_is_retrieved(lib):
    opts = dict(lib=lib, max_groups=16, max_pods=16, max_group_leaves=16, max_group_pods=8)
    spec = {"virtualCluster": "default", "priority": 0, "leafCellType": "K80", "leafCellNumber": 4,
            "affinityGroup": {"name": "gang", "members": [{"podNumber": 2, "leafCellNumber": 4}]}}

    def pod(name):
        p = alg.Pod(name, "ns", name)
        p.annotations[alg.ANNOTATION_POD_SCHEDULING_SPEC] = alg.to_yaml(spec)
        return p
    h = alg.HivedAlgorithm(new_config(_raw_c1()), **opts)
    for n in h.node_names:
        h.setHealthyNode(n)
    first = h.Schedule(pod("p0"), h.node_names, alg.FILTERING_PHASE).pod_bind_info
    assert first is not None and len(first["leafCellIsolation"]) == 4
    bound0 = alg.new_binding_pod(pod("p0"), first)
    h.close()
    # the cluster is reconfigured: the node of pod 0 is gone from the spec (renamed)
    other = [n for n in ("10.151.41.23", "10.151.41.24") if n != first["node"]][0]
    raw = _raw_c1("10.9.9.9", other) if first["node"] == "10.151.41.23" else _raw_c1(other, "10.9.9.9")
    h = alg.HivedAlgorithm(new_config(raw), **opts)
    for n in h.node_names:
        h.setHealthyNode(n)
    h.AddAllocatedPod(bound0)  # recovery: the cells of pod 0 are not found, the group exists without them
    g = h.GetAffinityGroup("gang")["status"]
    assert g["state"] == "Allocated" and g["allocatedPods"] == ["p0"]
    assert first["node"] not in g.get("physicalPlacement", {})
    second = h.Schedule(pod("p1"), h.node_names, alg.FILTERING_PHASE).pod_bind_info
    assert second is not None
    # the decision is insisted on: pod 0's placement comes back from its annotation, pod 1 gets the node that is left
    assert second["affinityGroupBindInfo"][0]["podPlacements"][0] == first["affinityGroupBindInfo"][0]["podPlacements"][0]
    assert second["node"] == other and second["node"] == second["affinityGroupBindInfo"][0]["podPlacements"][1]["physicalNode"]
    assert second["cellChain"] == first["cellChain"] and sorted(second["leafCellIsolation"]) == [0, 1, 2, 3]
    h.AddAllocatedPod(alg.new_binding_pod(pod("p1"), second))
    assert sorted(h.GetAffinityGroup("gang")["status"]["allocatedPods"]) == ["p0", "p1"]
    h.close()
    return second


def test_missing_placement_is_retrieved_device_program_matches_oracle(emu_lib, oracle_lib):
    assert _check_missing_placement_is_retrieved(emu_lib) == _check_missing_placement_is_retrieved(oracle_lib)


@pytest.mark.gpu
def test_missing_placement_is_retrieved_cuda_matches_oracle(cuda_lib, oracle_lib):
    assert _check_missing_placement_is_retrieved(cuda_lib) == _check_missing_placement_is_retrieved(oracle_lib)
