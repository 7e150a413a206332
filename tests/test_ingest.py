"""SURVEY.md section 8 row f1 — request ingest as built code (include/hived_ingest.h): node-name interning, the
NodeNames JSON array -> node bitmap (+ the cached previous request), the scheduling-spec annotation -> pod spec.
Host code: the same source is compiled into every backend, so the CPU tier exercises it through the emulation
library; tests/test_cabi.py checks that the product library exports the symbols."""
import ctypes as C
import json

import numpy as np
import pytest
import yaml

from hivedscheduler_b200 import _cabi, config, trace
from hivedscheduler_b200.algorithm import Pod, extract_pod_scheduling_spec, ANNOTATION_POD_SCHEDULING_SPEC
from hivedscheduler_b200.ingest import Ingest


@pytest.fixture(scope="module")
def ctx8k(emu_mt_lib):
    bc = trace.BatchContext(emu_mt_lib, config.config_c3(), 64, 64)
    ing = Ingest(emu_mt_lib, bc.ctx)
    names = [emu_mt_lib.hived_node_name(bc.ctx, i) for i in range(bc.n_nodes)]
    yield emu_mt_lib, bc, ing, names
    ing.close()
    bc.close()


def _bits(bm, words):
    a = np.frombuffer(bm, dtype=np.uint32, count=words)
    return {int(i) for i in np.flatnonzero(np.unpackbits(a.view(np.uint8), bitorder="little"))}


def test_node_names_to_bitmap(ctx8k):
    lib, bc, ing, names = ctx8k
    assert len(names) == 8192 and ing.words == 256
    rng = np.random.default_rng(7)
    pick = [int(i) for i in rng.choice(len(names), 3000, replace=False)]
    req = [names[i] for i in pick] + [b"not-a-node", names[pick[0]], b""]
    bm, cnt, is_all = ing.node_names(req)
    assert cnt == 3000 and not is_all
    assert _bits(bm, ing.words) == set(pick)
    bm, cnt, is_all = ing.node_names(list(reversed(names)))
    assert cnt == 8192 and is_all
    for i in (0, 17, 8191):
        assert lib.hived_ingest_node_id(ing.h, names[i], -1) == i
    assert lib.hived_ingest_node_id(ing.h, b"n99999", -1) == -1


def test_node_names_json_and_cache(ctx8k):
    lib, bc, ing, names = ctx8k
    sel = [n.decode() for n in names[5:4000:3]]
    body = json.dumps({"Pod": {"metadata": {"name": "p", "annotations": {"NodeNames": "decoy"}}}, "NodeNames": sel + ["zz\\u0041", "a\"b"],
                       "Nodes": None}).encode()
    off = ing.json_find(body, "NodeNames")
    assert body[off:off + 1] == b"["
    bm, cnt, is_all, used, cached = ing.node_names_json(body, off)
    assert cnt == len(sel) and not is_all and not cached
    assert body[off + used - 1:off + used] == b"]"
    assert _bits(bm, ing.words) == {5 + 3 * k for k in range(len(sel))}
    # the same array again (another request body around it): answered from the cache
    body2 = b'{"NodeNames":   ' + body[off:off + used] + b', "x": 1}'
    off2 = ing.json_find(body2, "NodeNames")
    bm2, cnt2, _, used2, cached2 = ing.node_names_json(body2, off2)
    assert cached2 and cnt2 == cnt and bytes(bm2) == bytes(bm) and used2 == used
    # every node, with whitespace between the items
    all_body = ("[ " + " ,\n ".join(json.dumps(n.decode()) for n in names) + " ]").encode()
    bm3, cnt3, is_all3, used3, cached3 = ing.node_names_json(all_body)
    assert cnt3 == 8192 and is_all3 and not cached3 and used3 == len(all_body)
    assert ing.node_names_json(b"[]")[1] == 0
    with pytest.raises(ValueError):
        ing.node_names_json(b'["n0001", 5]')
    with pytest.raises(ValueError):
        ing.node_names_json(b'["n0001"')
    assert ing.json_find(b'{"a": {"NodeNames": 1}, "b": 2}', "NodeNames") == -1  # nested keys are not top-level


def _mirror_spec(annotation: str, name: str = "default/pod-a"):
    ns, nm = name.split("/")
    pod = Pod(name=nm, namespace=ns, uid=name, annotations={ANNOTATION_POD_SCHEDULING_SPEC: annotation})
    return extract_pod_scheduling_spec(pod)


SPECS = [
    # what common.ToYaml emits for a PodSchedulingSpec (block style, sequences at the indentation of their key)
    """virtualCluster: vc1
priority: 1000
pinnedCellId: ""
leafCellType: B200
leafCellNumber: 8
gangReleaseEnable: false
lazyPreemptionEnable: true
ignoreK8sSuggestedNodes: false
affinityGroup:
  name: default/group1
  members:
  - podNumber: 2
    leafCellNumber: 8
  - podNumber: 1
    leafCellNumber: 4
""",
    # indented sequence, comments, quoted scalars, v1 field names
    """# a job submitted by an old client
virtualCluster: "vc0"   # quoted
priority: -1
gpuType: 'B200'
gpuNumber: 4
affinityGroup:
    name: "ns/with: colon"
    members:
        - podNumber: 3
          gpuNumber: 4
""",
    # no affinityGroup: a gang of its own
    "virtualCluster: vc3\npriority: 5\nleafCellNumber: 1\n",
    # flow style / JSON
    '{"virtualCluster": "vc2", "priority": 10, "leafCellNumber": 2, "lazyPreemptionEnable": true, '
    '"affinityGroup": {"name": "j/g", "members": [{"podNumber": 4, "leafCellNumber": 2}, {"podNumber": 1, "leafCellNumber": 16}]}}',
    "virtualCluster: vc1\npriority: 1\nleafCellNumber: 2\naffinityGroup: {name: g2, members: [{podNumber: 1, leafCellNumber: 2}]}\n",
]


@pytest.mark.parametrize("k", range(len(SPECS)))
def test_annotation_to_pod_spec_matches_the_mirror(ctx8k, k):
    lib, bc, ing, names = ctx8k
    ann = SPECS[k]
    want = _mirror_spec(ann)
    rc, sp, err = ing.pod_spec_yaml(ann.encode(), b"default/pod-a", 64, 64)
    assert rc == 0, err
    vcs = [lib.hived_vc_name(bc.ctx, i).decode() for i in range(lib.hived_num_vcs(bc.ctx))]
    assert vcs[sp.vc] == want["virtualCluster"]
    assert sp.priority == want["priority"] and sp.leaf_num == want["leafCellNumber"]
    assert bool(sp.flags & _cabi.SPEC_LAZY_PREEMPTION) == want["lazyPreemptionEnable"]
    assert bool(sp.flags & _cabi.SPEC_IGNORE_SUGGESTED) == want["ignoreK8sSuggestedNodes"]
    assert sp.pinned == -1
    assert sp.leaf_type == (-1 if want["leafCellType"] == "" else 0)
    members = want["affinityGroup"]["members"]
    assert sp.n_members == len(members)
    assert [(sp.member_pod_num[i], sp.member_leaf_num[i]) for i in range(sp.n_members)] == [(m["podNumber"], m["leafCellNumber"]) for m in members]
    assert sp.group == ing.lookup(Ingest.GROUPS, want["affinityGroup"]["name"].encode())
    assert sp.pod == ing.lookup(Ingest.PODS, b"default/pod-a")


@pytest.mark.parametrize("ann,msg", [
    ("", "Annotation does not exist or is empty"),
    ("priority: 1\nleafCellNumber: 1\n", "VirtualCluster is empty"),
    ("virtualCluster: vc0\npriority: -2\nleafCellNumber: 1\n", "Priority is less than -1"),
    ("virtualCluster: vc0\npriority: 1001\nleafCellNumber: 1\n", "Priority is greater than 1000"),
    ("virtualCluster: vc0\npriority: 1\n", "LeafCellNumber is non-positive"),
    ("virtualCluster: vc0\npriority: 1\nleafCellNumber: 1\naffinityGroup:\n  name: \"\"\n  members:\n  - podNumber: 1\n    leafCellNumber: 1\n",
     "AffinityGroup.Name is empty"),
    ("virtualCluster: vc0\npriority: 1\nleafCellNumber: 1\naffinityGroup:\n  name: g\n  members:\n  - podNumber: 0\n    leafCellNumber: 1\n",
     "AffinityGroup.Members has non-positive PodNumber"),
    ("virtualCluster: vc0\npriority: 1\nleafCellNumber: 1\naffinityGroup:\n  name: g\n  members:\n  - podNumber: 1\n    leafCellNumber: 2\n",
     "AffinityGroup.Members does not contains current Pod"),
])
def test_annotation_validation_messages(ctx8k, ann, msg):
    """internal/utils.go:244-287: the reference's messages, in the reference's order; the mirror raises the same."""
    lib, bc, ing, names = ctx8k
    rc, sp, err = ing.pod_spec_yaml(ann.encode(), b"default/p", 64, 64)
    assert rc != 0 and err.endswith(msg), err
    with pytest.raises(Exception) as ei:
        _mirror_spec(ann)
    assert msg in str(ei.value)


def test_unknown_names_and_id_recycling(ctx8k):
    lib, bc, ing, names = ctx8k
    rc, sp, err = ing.pod_spec_yaml(b"virtualCluster: nope\npriority: 1\nleafCellNumber: 1\nleafCellType: H100\npinnedCellId: pc9\n",
                                    b"default/q", 64, 64)
    assert rc == 0 and sp.vc == -1 and sp.leaf_type == -2 and sp.pinned == -2
    ids = [ing.intern(Ingest.GROUPS, b"grp%d" % i, 1000) for i in range(10)]
    assert len(set(ids)) == 10
    assert ing.intern(Ingest.GROUPS, b"grp3", 1000) == ids[3]
    assert ing.release(Ingest.GROUPS, b"grp7") == ids[7] and ing.release(Ingest.GROUPS, b"grp2") == ids[2]
    assert ing.lookup(Ingest.GROUPS, b"grp7") == -1
    assert ing.intern(Ingest.GROUPS, b"new-a", 1000) == min(ids[2], ids[7])  # lowest free id first
    assert ing.intern(Ingest.GROUPS, b"new-b", 1000) == max(ids[2], ids[7])
    full = Ingest(lib, bc.ctx)
    assert [full.intern(Ingest.PODS, b"p%d" % i, 3) for i in range(4)] == [0, 1, 2, -1]
    full.close()
