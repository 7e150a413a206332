"""Scheduler config: YAML -> defaulted spec -> the line-oriented "HIVEDSPEC" text the C-ABI ingests.

Host-side mirror of the reference's config layer (kept in the host language in the reference too):

* ``load_raw_config``   ~ api.InitRawConfig            (reference pkg/api/config.go:174-187)
* ``new_config``        ~ api.NewConfig + defaulting   (reference pkg/api/config.go:87-167)
* ``to_spec_text``      serialises the defaulted config for ``hived_create`` (include/hived.h).
  The reference hands ``*api.Config`` to ``algorithm.NewHivedAlgorithm`` (hived_algorithm.go:108);
  a cgo shim would emit exactly this text from its ``api.Config``.

The synthetic cluster generators (C1..C5 of SURVEY.md section 8d) live here as well so that tests,
bench and oracle all see byte-identical specs.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, List, Optional

import yaml


class ConfigError(Exception):
    """Platform-class error raised for malformed configs (the reference panics)."""


def load_raw_config(path: str) -> Dict[str, Any]:
    """Parse the YAML file into plain dicts/lists (api.InitRawConfig, config.go:174-187)."""
    with open(path, "r") as f:
        return yaml.safe_load(f) or {}


def _infer_physical_cell_spec(spec: Dict[str, Any], cts: Dict[str, Any], cell_type: str,
                              default_address: int, address_prefix: str) -> None:
    """api.inferPhysicalCellSpec (reference pkg/api/config.go:134-167)."""
    if not spec.get("cellType"):
        spec["cellType"] = cell_type
    addr = spec.get("cellAddress")
    if addr is None or addr == "":
        spec["cellAddress"] = address_prefix + str(default_address)
    else:
        spec["cellAddress"] = address_prefix + str(addr)
    ct = cts.get(cell_type)
    if ct is None:
        return  # leaf cell type
    if ct.get("isNodeLevel", False):
        default_address = 0
    n_child = int(ct.get("childCellNumber", 0))
    children = spec.get("cellChildren")
    if n_child > 0 and not children:
        children = [dict() for _ in range(n_child)]
        spec["cellChildren"] = children
    children = children or []
    for i, child in enumerate(children):
        if child is None:
            child = {}
            children[i] = child
        _infer_physical_cell_spec(child, cts, ct["childCellType"],
                                  default_address * n_child + i, spec["cellAddress"] + "/")


def new_config(raw: Dict[str, Any]) -> Dict[str, Any]:
    """api.NewConfig (reference pkg/api/config.go:87-118): defaulting of the physical cells.

    Mutates and returns ``raw`` like the reference does (the reference's tests rely on editing the
    returned object afterwards, hived_algorithm_test.go:820-821, 1051-1074).
    """
    c = raw
    if c.get("physicalCluster") is None:
        c["physicalCluster"] = {}
    if c.get("virtualClusters") is None:
        c["virtualClusters"] = {}
    pc = c["physicalCluster"]
    pc.setdefault("cellTypes", {})
    if pc["cellTypes"] is None:
        pc["cellTypes"] = {}
    pc.setdefault("physicalCells", [])
    if pc["physicalCells"] is None:
        pc["physicalCells"] = []
    cts = pc["cellTypes"]
    for idx, cell in enumerate(pc["physicalCells"]):
        ct = cell.get("cellType")
        if ct not in cts:
            raise ConfigError("physicalCells contains unknown cellType: %s" % ct)
        _infer_physical_cell_spec(cell, cts, ct, idx, "")
    return c


def load_config(path: str) -> Dict[str, Any]:
    return new_config(load_raw_config(path))


def _tok(s: Any) -> str:
    s = str(s)
    if s == "" or any(ch.isspace() for ch in s):
        raise ConfigError("names in the spec must be non-empty and free of whitespace: %r" % s)
    return s


def to_spec_text(config: Dict[str, Any]) -> str:
    """Serialise a defaulted config into HIVEDSPEC text (format documented in include/hived.h)."""
    pc = config["physicalCluster"]
    cts = pc.get("cellTypes") or {}
    out: List[str] = ["HIVEDSPEC 1"]
    out.append("celltypes %d" % len(cts))
    for name in sorted(cts):
        ct = cts[name]
        out.append("%s %s %d %d" % (_tok(name), _tok(ct["childCellType"]),
                                    int(ct["childCellNumber"]), 1 if ct.get("isNodeLevel") else 0))
    cells = pc.get("physicalCells") or []
    out.append("physicalcells %d" % len(cells))

    def emit(cell: Dict[str, Any], depth: int) -> None:
        children = cell.get("cellChildren") or []
        pid = cell.get("pinnedCellId") or "-"
        out.append("%d %s %s %s %d" % (depth, _tok(cell["cellType"]), _tok(cell["cellAddress"]),
                                       _tok(pid), len(children)))
        for ch in children:
            emit(ch, depth + 1)

    for cell in cells:
        emit(cell, 0)
    vcs = config.get("virtualClusters") or {}
    out.append("virtualclusters %d" % len(vcs))
    for vcn in sorted(vcs):
        spec = vcs[vcn] or {}
        vcells = spec.get("virtualCells") or []
        pinned = spec.get("pinnedCells") or []
        out.append("vc %s %d %d" % (_tok(vcn), len(vcells), len(pinned)))
        for v in vcells:
            out.append("%s %d" % (_tok(v["cellType"]), int(v["cellNumber"])))
        for p in pinned:
            out.append("%s" % _tok(p["pinnedCellId"]))
    out.append("end")
    return "\n".join(out) + "\n"


# ----------------------------------------------------------------------------------------------
# Synthetic clusters of SURVEY.md section 8(d)
# ----------------------------------------------------------------------------------------------

def config_c1() -> Dict[str, Any]:
    """C1: 2 nodes x 4 GPU, chain 2-K80-NODE (4 levels), 1 VC (cell types as in the reference's
    example/feature/file/hived-config-1.yaml); expected placements in SURVEY.md section 8(c)."""
    raw = {
        "physicalCluster": {
            "cellTypes": {
                "K80-2GPU": {"childCellType": "K80", "childCellNumber": 2},
                "K80-NODE": {"childCellType": "K80-2GPU", "childCellNumber": 2, "isNodeLevel": True},
                "2-K80-NODE": {"childCellType": "K80-NODE", "childCellNumber": 2},
            },
            "physicalCells": [
                {"cellType": "2-K80-NODE",
                 "cellChildren": [{"cellAddress": "10.151.41.23"}, {"cellAddress": "10.151.41.24"}]},
            ],
        },
        "virtualClusters": {
            "default": {"virtualCells": [{"cellType": "2-K80-NODE", "cellNumber": 1}]},
        },
    }
    return new_config(raw)


def _synthetic(levels: List[tuple], n_top: int, node_level_name: str, vcs: Dict[str, List[tuple]],
               node_fmt: str) -> Dict[str, Any]:
    """levels: [(type, child_type, fanout)] bottom-up above the leaf type."""
    cts = {}
    for name, child, fanout in levels:
        cts[name] = {"childCellType": child, "childCellNumber": fanout}
        if name == node_level_name:
            cts[name]["isNodeLevel"] = True
    top = levels[-1][0]
    # explicit node names: walk down from the top to the node level
    chain = [top]
    while chain[-1] != node_level_name:
        chain.append(cts[chain[-1]]["childCellType"])
    counter = [0]

    def build(depth: int) -> Dict[str, Any]:
        if chain[depth] == node_level_name:
            d = {"cellAddress": node_fmt % counter[0]}
            counter[0] += 1
            return d
        return {"cellChildren": [build(depth + 1) for _ in range(cts[chain[depth]]["childCellNumber"])]}

    cells = []
    for _ in range(n_top):
        c = build(0)
        c["cellType"] = top
        cells.append(c)
    raw = {
        "physicalCluster": {"cellTypes": cts, "physicalCells": cells},
        "virtualClusters": {vcn: {"virtualCells": [{"cellType": t, "cellNumber": n} for t, n in vl]}
                            for vcn, vl in vcs.items()},
    }
    return new_config(raw)


def config_c2() -> Dict[str, Any]:
    """C2: 1024 nodes x 8 GPU, 4 levels (GPU/HALF/NODE/RACK), 32 racks, 4 VCs x 7 racks."""
    levels = [("HALF", "B200", 4), ("NODE", "HALF", 2), ("RACK", "NODE", 32)]
    vcs = {"vc%d" % i: [("RACK", 7)] for i in range(4)}
    return _synthetic(levels, 32, "NODE", vcs, "n%04d")


def config_c3(n_pods: int = 16, n_vcs: int = 8, racks_per_vc: int = 12) -> Dict[str, Any]:
    """C3/C4/C5: 8192 nodes x 8 GPU = 65536 GPUs, 5 levels (GPU/HALF/NODE/RACK/POD), 16 PODs;
    8 VCs each POD x1 + POD.RACK x12 (28 racks = 7168 GPUs per VC)."""
    levels = [("HALF", "B200", 4), ("NODE", "HALF", 2), ("RACK", "NODE", 32), ("POD", "RACK", 16)]
    vcs = {"vc%d" % i: [("POD", 1), ("POD.RACK", racks_per_vc)] for i in range(n_vcs)}
    return _synthetic(levels, n_pods, "NODE", vcs, "n%04d")


def deepcopy_config(c: Dict[str, Any]) -> Dict[str, Any]:
    return copy.deepcopy(c)
