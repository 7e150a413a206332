"""Would evaluating the next gang of a VC speculatively (on the state BEFORE the current gang commits) give the right
answer?  For every pair of consecutive scheduling decisions (i, j) of one VC in the C3 trace with no release of that VC
in between: Schedule(j) without commit on the state before i, then compare with what j really gets after i.
The share that differs is the replay rate of a 2-wide speculation window — the number DESIGN.md section 4 quotes.
Runs on any library of the ABI (default: the host build of the device program, which answers a call in ~1 us).

    python profiles/micro/speculation_replay_rate.py [n_gangs] [lib.so]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from hivedscheduler_b200 import _cabi, trace  # noqa: E402


def main():
    n_gangs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    lib = _cabi.load_library(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "_build", "libhived_emu_mt.so"))
    t = trace.trace_c3(n_gangs=n_gangs)
    ev, dec = np.ascontiguousarray(t["events"]), t["decision"]
    bc = trace.BatchContext(lib, t["config"], t["n_groups"], t["n_pods"], t["max_group_leaves"], t["max_group_pods"])
    bc.set_all_nodes_healthy()
    n = len(ev)
    # next decision of the same VC, if no DELETE of that VC lies in between
    nxt = [-1] * n
    last_dec, dirty = {}, {}
    group_vc = {}
    for i in range(n):
        e = ev[i]
        if e["type"] == _cabi.EV_SCHEDULE:
            vc = int(e["spec"]["vc"])
            group_vc[int(e["spec"]["group"])] = vc
            if dec[i]:
                if vc in last_dec and not dirty.get(vc):
                    nxt[last_dec[vc]] = i
                last_dec[vc] = i
                dirty[vc] = False
        else:
            dirty[group_vc.get(int(e["spec"]["group"]), -1)] = True
    cap = 3 * 64 + 4096
    pool, pool2 = (C.c_int32 * cap)(), (C.c_int32 * cap)()
    res, res2 = _cabi.Result(), _cabi.Result()
    spec_pick = {}
    pairs = same = 0
    by_shape = {}
    evp = ev.ctypes.data_as(C.POINTER(_cabi.Event))
    one = C.sizeof(_cabi.Event)
    for i in range(n):
        if ev[i]["type"] == _cabi.EV_SCHEDULE and dec[i] and nxt[i] >= 0:
            j = nxt[i]
            sp = _cabi.PodSpec.from_buffer_copy(ev[j]["spec"].tobytes())
            rc = lib.hived_schedule(bc.ctx, C.byref(sp), None, _cabi.PHASE_PREEMPTING, C.byref(res2), pool2, cap)
            assert rc == 0
            spec_pick[j] = (res2.kind, tuple(pool2[res2.leaf_off:res2.leaf_off + 3 * res2.n_leaves]))
        rc = lib.hived_process_events(bc.ctx, C.cast(C.addressof(evp.contents) + i * one, C.POINTER(_cabi.Event)), 1, None, 0,
                                      C.byref(res), pool, cap)
        assert rc == 0
        if i in spec_pick:
            real = (res.kind, tuple(pool[res.leaf_off:res.leaf_off + 3 * res.n_leaves]))
            ok = real == spec_pick.pop(i)
            pairs += 1
            same += ok
            shape = (int(ev[i]["spec"]["leaf_num"]), int(ev[i]["spec"]["member_pod_num"][0]))
            a, b = by_shape.get(shape, (0, 0))
            by_shape[shape] = (a + 1, b + ok)
    print("pairs of consecutive same-VC decisions (no release in between): %d" % pairs)
    print("speculative answer still right after the previous gang committed: %.1f %%  -> replay rate %.1f %%" % (
        100.0 * same / pairs, 100.0 * (1 - same / pairs)))
    for shape, (a, b) in sorted(by_shape.items()):
        print("  gang of %d pod(s) x %d GPU: %6d pairs, replay rate %.1f %%" % (shape[1], shape[0], a, 100.0 * (1 - b / a)))
    bc.close()


if __name__ == "__main__":
    main()
