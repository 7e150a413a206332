"""ctypes binding of include/hived_frontend.h: the extender's pod state machine with
batch draining.  The state machine and the drain live in the library (csrc/hived_frontend.hpp); this file types the
entry points and offers a thin convenience class for tests, bench.py and a Python extender."""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _cabi
from .ingest import Ingest

POD_UNKNOWN, POD_WAITING, POD_PREEMPTING, POD_BINDING, POD_BOUND = range(5)
FE_BIND, FE_WAIT, FE_PREEMPT, FE_ERROR, FE_NONE = 1, 2, 3, 4, 5
MAX_LEAVES = MAX_VICTIMS = 64


class Config(C.Structure):
    _fields_ = [("waiting_block_ms", C.c_int32), ("force_bind_threshold", C.c_int32), ("max_batch", C.c_int32),
                ("max_groups", C.c_int32), ("max_pods", C.c_int32), ("reserved", C.c_int32 * 3)]


class Response(C.Structure):
    _fields_ = [("kind", C.c_int32), ("error", C.c_int32), ("node", C.c_int32), ("insisted", C.c_int32),
                ("force_bind", C.c_int32), ("bind_attempts", C.c_int32), ("chain", C.c_int32), ("n_leaves", C.c_int32),
                ("leaf_index", C.c_int32 * MAX_LEAVES), ("wait_code", C.c_int32), ("wait_cell", C.c_int32),
                ("n_victims", C.c_int32), ("victim_pod", C.c_int32 * MAX_VICTIMS), ("victim_node", C.c_int32 * MAX_VICTIMS),
                ("batch_events", C.c_int32), ("message", C.c_char * 160)]


_P = C.c_void_p
SYMBOLS = [
    ("hived_fe_create", C.c_int, [_P, _P, C.POINTER(Config), C.POINTER(_P)]),
    ("hived_fe_destroy", None, [_P]),
    ("hived_fe_add_unbound_pod", C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64]),
    ("hived_fe_add_bound_pod", C.c_int,
     [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(_cabi.BindInfo), C.POINTER(C.c_int32), C.c_int32]),
    ("hived_fe_delete_pod", C.c_int, [_P, C.c_char_p]),
    ("hived_fe_pod_state", C.c_int32, [_P, C.c_char_p]),
    ("hived_fe_filter", C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(Response)]),
    ("hived_fe_preempt", C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(Response)]),
    ("hived_fe_bind_check", C.c_int, [_P, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]),
    ("hived_fe_enqueue_filter", C.c_int64, [_P, C.c_char_p, C.c_char_p, C.c_int64]),
    ("hived_fe_drain", C.c_int, [_P]),
    ("hived_fe_take", C.c_int, [_P, C.c_int64, C.POINTER(Response)]),
    ("hived_fe_stats", C.c_int, [_P, C.POINTER(C.c_int64), C.c_int32]),
    ("hived_fe_last_error", C.c_char_p, [_P]),
]
STAT_NAMES = ["answered", "drains", "events", "largest_batch", "wait_answers", "held_ms", "per_call"]


def bind(lib: C.CDLL) -> None:
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


class FrontEnd:
    def __init__(self, lib: C.CDLL, ctx, max_groups: int, max_pods: int, waiting_block_ms: int = 0,
                 force_bind_threshold: int = 3, max_batch: int = 4096):
        bind(lib)
        self.lib = lib
        self.ingest = Ingest(lib, ctx)
        cfg = Config(waiting_block_ms=waiting_block_ms, force_bind_threshold=force_bind_threshold, max_batch=max_batch,
                     max_groups=max_groups, max_pods=max_pods)
        self.h = _P()
        rc = lib.hived_fe_create(ctx, self.ingest.h, C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise RuntimeError("hived_fe_create failed (%d)" % rc)

    def close(self):
        if self.h:
            self.lib.hived_fe_destroy(self.h)
            self.h = None
            self.ingest.close()

    def add_unbound_pod(self, uid: str, key: str, annotation: str) -> int:
        a = annotation.encode()
        return self.lib.hived_fe_add_unbound_pod(self.h, uid.encode(), key.encode(), a, len(a))

    def delete_pod(self, uid: str) -> int:
        return self.lib.hived_fe_delete_pod(self.h, uid.encode())

    def pod_state(self, uid: str) -> int:
        return int(self.lib.hived_fe_pod_state(self.h, uid.encode()))

    def filter(self, uid: str, node_names_json: Optional[bytes] = None) -> Response:
        r = Response()
        self.lib.hived_fe_filter(self.h, uid.encode(), node_names_json, len(node_names_json) if node_names_json else 0, C.byref(r))
        return r

    def preempt(self, uid: str, node_names_json: Optional[bytes] = None) -> Response:
        r = Response()
        self.lib.hived_fe_preempt(self.h, uid.encode(), node_names_json, len(node_names_json) if node_names_json else 0, C.byref(r))
        return r

    def enqueue_filter(self, uid: str, node_names_json: Optional[bytes] = None) -> int:
        return int(self.lib.hived_fe_enqueue_filter(self.h, uid.encode(), node_names_json, len(node_names_json) if node_names_json else 0))

    def drain(self) -> int:
        return self.lib.hived_fe_drain(self.h)

    def take(self, ticket: int) -> Optional[Response]:
        r = Response()
        return r if self.lib.hived_fe_take(self.h, ticket, C.byref(r)) == 0 else None

    def bind_check(self, uid: str, node: int):
        buf = C.create_string_buffer(256)
        rc = self.lib.hived_fe_bind_check(self.h, uid.encode(), node, buf, 256)
        return rc, buf.value.decode()

    def stats(self) -> dict:
        out = (C.c_int64 * 8)()
        n = self.lib.hived_fe_stats(self.h, out, 8)
        return {STAT_NAMES[i]: int(out[i]) for i in range(n)}
