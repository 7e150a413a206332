"""Ingest micro-benchmark: one scheduling request naming 8192 nodes (the 64k-GPU
cluster of BASELINE configs[2]), host work only.  Legs, microseconds per request (median of `reps`):

  json.loads + python set      what a Python front end would do (and the shape of the reference: JSON decode into
                               []string, webserver.go:173-182, then one set insert per name, hived_algorithm.go:190-193)
  python dict -> bitmap        names already decoded, the pre-round-2 mirror
  C names[] -> bitmap          hived_ingest_node_names on decoded strings (the cgo shim's form)
  C JSON array -> bitmap       hived_ingest_node_names_json on the raw request body: no string is materialised
  C JSON array, cached         the same body again (kube-scheduler resends the same feasible-node list)
  annotation YAML -> pod spec  hived_ingest_pod_spec_yaml  vs  yaml.safe_load + the mirror's defaulting/validation

usage: python profiles/micro/ingest_bench.py [path/to/lib.so]     (default: the CUDA library; any backend has the same host code)
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hivedscheduler_b200 import _cabi, config, trace  # noqa: E402
from hivedscheduler_b200.algorithm import ANNOTATION_POD_SCHEDULING_SPEC, Pod, extract_pod_scheduling_spec  # noqa: E402
from hivedscheduler_b200.ingest import Ingest  # noqa: E402


def med_us(fn, reps=200):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e6 * statistics.median(ts)


def main():
    lib = _cabi.load_library(sys.argv[1]) if len(sys.argv) > 1 else _cabi.load_cuda_library()
    bc = trace.BatchContext(lib, config.config_c3(), 64, 64)
    ing = Ingest(lib, bc.ctx)
    names_b = [lib.hived_node_name(bc.ctx, i) for i in range(bc.n_nodes)]
    names_s = [n.decode() for n in names_b]
    body = json.dumps({"Pod": {"metadata": {"name": "p"}}, "NodeNames": names_s}).encode()
    body_b = json.dumps({"Pod": {"metadata": {"name": "q"}}, "NodeNames": list(reversed(names_s))}).encode()
    off, off_b = ing.json_find(body, "NodeNames"), ing.json_find(body_b, "NodeNames")
    ids = {n: i for i, n in enumerate(names_s)}
    words = ing.words
    bm = ing.new_bitmap()

    def py_json_set():
        return set(json.loads(body)["NodeNames"])

    def py_bitmap():
        w = [0] * words
        for n in names_s:
            i = ids.get(n)
            if i is not None:
                w[i >> 5] |= 1 << (i & 31)
        return w

    flip = [0]

    def c_json():  # alternate two bodies so that the cache never hits
        flip[0] ^= 1
        return ing.node_names_json(body_b if flip[0] else body, off_b if flip[0] else off, bm)

    ann = ("virtualCluster: vc1\npriority: 1000\nleafCellType: B200\nleafCellNumber: 8\nlazyPreemptionEnable: true\n"
           "affinityGroup:\n  name: default/group1\n  members:\n  - podNumber: 2\n    leafCellNumber: 8\n  - podNumber: 1\n    leafCellNumber: 4\n")
    ann_b = ann.encode()
    pod = Pod(name="p", namespace="default", uid="default/p", annotations={ANNOTATION_POD_SCHEDULING_SPEC: ann})
    out = {
        "nodes_per_request": len(names_s), "request_body_bytes": len(body),
        "us_per_request": {
            "json.loads + python set": med_us(py_json_set, 50),
            "python dict -> bitmap": med_us(py_bitmap, 50),
            "C names[] -> bitmap (incl. ctypes array build)": med_us(lambda: ing.node_names(names_b, bm), 50),
            "C JSON array -> bitmap": med_us(c_json),
            "C JSON array, cached": med_us(lambda: ing.node_names_json(body, off, bm)),
            "C json_find(NodeNames)": med_us(lambda: ing.json_find(body, "NodeNames")),
            "yaml.safe_load + mirror validation": med_us(lambda: extract_pod_scheduling_spec(pod)),
            "C annotation YAML -> pod spec": med_us(lambda: ing.pod_spec_yaml(ann_b, b"default/p", 64, 64)),
        },
        "backend": lib.hived_backend().decode(),
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
