"""SURVEY.md section 8 row f4 — the extender's pod state machine with batch draining (include/hived_frontend.h).
CPU tier: the front end is host code above the public ABI, exercised here over the test-only emulation of the device
program.  Oracle of the answers: the reference-shaped mirror (hivedscheduler_b200/algorithm.py: Schedule +
AddAllocatedPod one call at a time, pkg/scheduler/scheduler.go:485-583) running on the CPU checker library."""
import ctypes as C
import json
import threading
import time

import pytest

from hivedscheduler_b200 import _cabi, config, frontend as fe_mod, trace
from hivedscheduler_b200.algorithm import (ANNOTATION_POD_SCHEDULING_SPEC, HivedAlgorithm, Pod, new_binding_pod,
                                           FILTERING_PHASE, PREEMPTING_PHASE)

MAXG, MAXP = 4096, 16384


def cluster():
    return config.config_c3(n_pods=4, n_vcs=2, racks_per_vc=6)


def annotation(vc, prio, leaf_num, group=None, pods=1, lazy=False):
    s = "virtualCluster: vc%d\npriority: %d\nleafCellType: B200\nleafCellNumber: %d\nlazyPreemptionEnable: %s\n" % (
        vc, prio, leaf_num, "true" if lazy else "false")
    if group:
        s += "affinityGroup:\n  name: %s\n  members:\n  - podNumber: %d\n    leafCellNumber: %d\n" % (group, pods, leaf_num)
    return s


@pytest.fixture()
def fe(emu_mt_lib):
    bc = trace.BatchContext(emu_mt_lib, cluster(), MAXG, MAXP, 64, 8)
    bc.set_all_nodes_healthy()
    f = fe_mod.FrontEnd(emu_mt_lib, bc.ctx, MAXG, MAXP)
    yield f, bc
    f.close()
    bc.close()


def workload(n=60):
    """(uid, key, annotation) of n pods: singles, 8-GPU pods and 4-pod gangs over 2 VCs."""
    out = []
    for i in range(n):
        vc = i % 2
        if i % 5 == 4:
            for j in range(4):
                out.append(("uid-g%d-%d" % (i, j), "default/g%d-%d" % (i, j), annotation(vc, 10, 8, "default/gang%d" % i, 4)))
        else:
            out.append(("uid-%d" % i, "default/p%d" % i, annotation(vc, 10, [1, 4, 8][i % 3])))
    return out


def mirror_answers(oracle_lib, pods, deletes_after=()):
    """The reference's shape: one filterRoutine at a time (Schedule, then AddAllocatedPod on a bind result)."""
    alg = HivedAlgorithm(cluster(), lib=oracle_lib, max_groups=MAXG, max_pods=MAXP, max_group_leaves=64, max_group_pods=8)
    for n in alg.node_names:
        alg.AddNode({"name": n, "healthy": True})
    out, bound = {}, {}
    for k, (uid, key, ann) in enumerate(pods):
        ns, nm = key.split("/")
        pod = Pod(name=nm, namespace=ns, uid=uid, annotations={ANNOTATION_POD_SCHEDULING_SPEC: ann})
        r = alg.Schedule(pod, alg.node_names, FILTERING_PHASE)
        if r.pod_bind_info is not None:
            b = new_binding_pod(pod, r.pod_bind_info)
            alg.AddAllocatedPod(b)
            bound[uid] = b
            out[uid] = ("bind", r.pod_bind_info["node"], tuple(r.pod_bind_info["leafCellIsolation"]))
        elif r.pod_preempt_info is not None:
            out[uid] = ("preempt",)
        else:
            out[uid] = ("wait",)
        if k in deletes_after:
            victim = deletes_after[k]
            alg.DeleteAllocatedPod(bound.pop(victim))
    alg.close()
    return out


def fe_answer(f, lib, bc, r):
    if r.kind == fe_mod.FE_BIND:
        return ("bind", lib.hived_node_name(bc.ctx, r.node).decode(), tuple(r.leaf_index[:r.n_leaves]))
    return {fe_mod.FE_WAIT: ("wait",), fe_mod.FE_PREEMPT: ("preempt",)}.get(r.kind, ("error", r.message.decode()))


def test_one_batch_answers_like_the_reference_one_by_one(fe, emu_mt_lib, oracle_lib):
    f, bc = fe
    pods = workload()
    deletes = {10: "uid-3", 25: "uid-7", 40: "uid-g4-1"}
    want = mirror_answers(oracle_lib, pods, deletes)
    for uid, key, ann in pods:
        assert f.add_unbound_pod(uid, key, ann) == 0
    tickets = {}
    for k, (uid, key, ann) in enumerate(pods):
        tickets[uid] = f.enqueue_filter(uid)
        if k in deletes:
            f.delete_pod(deletes[k])
    assert f.drain() == 0
    got = {uid: fe_answer(f, emu_mt_lib, bc, f.take(t)) for uid, t in tickets.items()}
    assert got == want
    st = f.stats()
    # the deletion of uid-3 follows its own filter call: its kind depends on that answer, so it opens a second batch
    assert st["drains"] == 2 and st["events"] == len(pods) + len(deletes) and st["answered"] == len(pods)
    assert st["largest_batch"] >= len(pods) - 11
    assert f.pod_state("uid-0") == fe_mod.POD_BINDING and f.pod_state("uid-3") == fe_mod.POD_UNKNOWN


def test_blocking_calls_one_at_a_time_give_the_same_answers(fe, emu_mt_lib, oracle_lib):
    f, bc = fe
    pods = workload(30)
    want = mirror_answers(oracle_lib, pods)
    got = {}
    for uid, key, ann in pods:
        f.add_unbound_pod(uid, key, ann)
        got[uid] = fe_answer(f, emu_mt_lib, bc, f.filter(uid))
    assert got == want
    assert f.stats()["drains"] == len(pods)


def test_pod_state_machine(fe, emu_mt_lib):
    f, bc = fe
    lib = emu_mt_lib
    r = f.filter("nobody")
    assert r.kind == fe_mod.FE_ERROR and r.error == 1 or b"does not exist" in r.message
    assert b"Pod does not exist, completed or has not been informed to the scheduler" in r.message
    f.add_unbound_pod("u1", "default/a", annotation(0, 5, 4))
    assert f.pod_state("u1") == fe_mod.POD_WAITING
    r1 = f.filter("u1")
    assert r1.kind == fe_mod.FE_BIND and not r1.insisted and r1.n_leaves == 4 and not r1.force_bind
    assert f.pod_state("u1") == fe_mod.POD_BINDING
    # the same pod again: the previous decision is insisted on, attempts counted, force bind at the threshold
    for attempt in (1, 2, 3):
        r = f.filter("u1")
        assert r.kind == fe_mod.FE_BIND and r.insisted and r.node == r1.node and r.bind_attempts == attempt
        assert list(r.leaf_index[:4]) == list(r1.leaf_index[:4])
        assert bool(r.force_bind) == (attempt >= 3)
    # ... or at once when the decided node is not among the suggested ones (validatePodBindInfo)
    other = lib.hived_node_name(bc.ctx, (r1.node + 1) % bc.n_nodes).decode()
    f.add_unbound_pod("u2", "default/b", annotation(0, 5, 4))
    r2 = f.filter("u2", json.dumps([other]).encode())
    assert r2.kind == fe_mod.FE_BIND
    if lib.hived_node_name(bc.ctx, r2.node).decode() != other:
        assert r2.force_bind
    # bindRoutine's checks
    assert f.bind_check("u1", r1.node) == (0, "")
    rc, msg = f.bind_check("u1", (r1.node + 1) % bc.n_nodes)
    assert rc != 0 and msg.startswith("Pod binding node mismatch: expected ")
    f.add_unbound_pod("u3", "default/c", annotation(0, 5, 4))
    rc, msg = f.bind_check("u3", 0)
    assert rc != 0 and "cannot be bound without a scheduling placement" in msg
    # the informer reports u1 bound: further filter calls are refused
    assert f.lib.hived_fe_add_bound_pod(f.h, b"u1", b"default/a", None, 0, None, None, 0) == 0
    assert f.pod_state("u1") == fe_mod.POD_BOUND
    r = f.filter("u1")
    assert r.kind == fe_mod.FE_ERROR and b"Pod has already been bound to node " in r.message
    # a bad annotation surfaces when the pod is scheduled, as a 400
    f.add_unbound_pod("u4", "default/d", "virtualCluster: vc0\npriority: 5000\nleafCellNumber: 1\n")
    r = f.filter("u4")
    assert r.kind == fe_mod.FE_ERROR and 1 <= r.error < 100 and b"Priority is greater than 1000" in r.message
    # deleting an allocated pod frees its cells: the next pod of the same shape lands on them
    f.delete_pod("u1")
    f.add_unbound_pod("u5", "default/e", annotation(0, 5, 4))
    r5 = f.filter("u5")
    assert r5.kind == fe_mod.FE_BIND and f.pod_state("u1") == fe_mod.POD_UNKNOWN


def test_a_pod_appears_once_per_batch(fe):
    f, bc = fe
    f.add_unbound_pod("u1", "default/a", annotation(0, 5, 8))
    f.add_unbound_pod("u2", "default/b", annotation(0, 5, 8))
    t1, t2, t3 = f.enqueue_filter("u1"), f.enqueue_filter("u2"), f.enqueue_filter("u1")
    f.drain()
    a, b, c = f.take(t1), f.take(t2), f.take(t3)
    assert a.kind == b.kind == c.kind == fe_mod.FE_BIND
    assert not a.insisted and c.insisted and c.node == a.node and c.bind_attempts == 1
    assert a.batch_events == 2 and f.stats()["drains"] == 1  # the repeated request was answered from the state


def test_concurrent_filter_calls_are_drained_in_batches(emu_mt_lib):
    """Callers that arrive while the scheduler is busy queue up and are answered by ONE batch.  (The emulated device
    answers a single event in microseconds, so the test keeps the scheduler busy with the reference's own means:
    a pod that has to wait holds it for waiting_block_ms.)"""
    bc = trace.BatchContext(emu_mt_lib, cluster(), MAXG, MAXP, 64, 8)
    bc.set_all_nodes_healthy()
    f = fe_mod.FrontEnd(emu_mt_lib, bc.ctx, MAXG, MAXP, waiting_block_ms=300)
    pods = [("uid-%d" % i, "default/p%d" % i, annotation(i % 2, 10, [1, 4, 8][i % 3])) for i in range(48)]
    for uid, key, ann in pods:
        f.add_unbound_pod(uid, key, ann)
    f.add_unbound_pod("too-big", "default/too-big", annotation(0, 10, 16))  # no node has 16 GPUs: this pod waits
    answers = {}

    def client(uid):
        answers[uid] = f.filter(uid)

    first = threading.Thread(target=client, args=("too-big",))
    first.start()
    time.sleep(0.1)  # the scheduler is now held by the WAIT answer
    threads = [threading.Thread(target=client, args=(uid,)) for uid, _, _ in pods]
    for t in threads:
        t.start()
    for t in [first] + threads:
        t.join()
    assert answers["too-big"].kind == fe_mod.FE_WAIT
    assert all(answers[uid].kind == fe_mod.FE_BIND for uid, _, _ in pods)
    st = f.stats()
    assert st["answered"] == len(pods) + 1 and st["events"] == len(pods) + 1
    assert st["drains"] < 10 and st["largest_batch"] >= 24, st
    assert max(r.batch_events for r in answers.values()) == st["largest_batch"]
    # whatever the interleaving was, no GPU was handed out twice
    used = set()
    for uid, _, _ in pods:
        r = answers[uid]
        for k in range(r.n_leaves):
            cell = (r.node, r.leaf_index[k])
            assert cell not in used
            used.add(cell)
    f.close()
    bc.close()


def test_waiting_block_holds_the_scheduler(emu_mt_lib):
    bc = trace.BatchContext(emu_mt_lib, cluster(), MAXG, MAXP, 64, 8)
    bc.set_all_nodes_healthy()
    f = fe_mod.FrontEnd(emu_mt_lib, bc.ctx, MAXG, MAXP, waiting_block_ms=40)
    # vc0 owns 1 POD + 6 racks = 22 racks x 32 nodes: ask for more 8-GPU pods than it has nodes
    n = 22 * 32 + 3
    for i in range(n):
        f.add_unbound_pod("u%d" % i, "default/p%d" % i, annotation(0, 10, 8))
    tickets = [f.enqueue_filter("u%d" % i) for i in range(n)]
    t0 = time.perf_counter()
    f.drain()
    dt = time.perf_counter() - t0
    kinds = [f.take(t).kind for t in tickets]
    assert kinds.count(fe_mod.FE_BIND) == 22 * 32 and kinds.count(fe_mod.FE_WAIT) == 3
    st = f.stats()
    assert st["wait_answers"] == 3 and st["held_ms"] == 120 and dt >= 0.12
    assert f.pod_state("u%d" % (n - 1)) == fe_mod.POD_WAITING
    f.close()
    bc.close()


def test_preemption_through_the_front_end(fe, emu_mt_lib):
    f, bc = fe
    # fill vc0 with low-priority 8-GPU pods, then a high-priority pod arrives
    n = 22 * 32
    for i in range(n):
        f.add_unbound_pod("low%d" % i, "default/low%d" % i, annotation(0, 1, 8))
    tickets = [f.enqueue_filter("low%d" % i) for i in range(n)]
    f.drain()
    assert all(f.take(t).kind == fe_mod.FE_BIND for t in tickets)
    f.add_unbound_pod("high", "default/high", annotation(0, 100, 8))
    r = f.filter("high")
    assert r.kind == fe_mod.FE_PREEMPT and r.n_victims >= 1  # FailedNodes: preemption may help
    assert f.pod_state("high") == fe_mod.POD_WAITING
    rp = f.preempt("high")
    assert rp.kind == fe_mod.FE_PREEMPT and rp.n_victims >= 1 and f.pod_state("high") == fe_mod.POD_PREEMPTING
    assert f.stats()["per_call"] == 1
    # the victims are deleted (kube-scheduler evicts them); the preemptor binds onto the freed node
    victims = {rp.victim_pod[k] for k in range(rp.n_victims)}
    node = rp.victim_node[0]
    for i in range(n):
        if f.ingest.lookup(1, b"low%d" % i) in victims:
            f.delete_pod("low%d" % i)
    r2 = f.filter("high")
    assert r2.kind == fe_mod.FE_BIND and r2.node == node and f.pod_state("high") == fe_mod.POD_BINDING
    rb = f.preempt("high")
    assert rb.kind == fe_mod.FE_ERROR and b"Pod has already been binding to node " in rb.message
