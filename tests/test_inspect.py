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
