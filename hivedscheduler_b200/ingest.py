"""ctypes binding of include/hived_ingest.h: node-name interning, the NodeNames JSON
array -> node bitmap (with the cached previous request), the scheduling-spec annotation -> hived_pod_spec_t.  The work
is done by the C helpers inside the library; this file only types them."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

from . import _cabi

_P = C.c_void_p
SYMBOLS = [
    ("hived_ingest_create", C.c_int, [_P, C.POINTER(_P)]),
    ("hived_ingest_destroy", None, [_P]),
    ("hived_ingest_bitmap_words", C.c_int32, [_P]),
    ("hived_ingest_node_id", C.c_int32, [_P, C.c_char_p, C.c_int32]),
    ("hived_ingest_node_names", C.c_int32, [_P, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    ("hived_ingest_node_names_json", C.c_int32,
     [_P, C.c_char_p, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    ("hived_ingest_json_find", C.c_int64, [C.c_char_p, C.c_int64, C.c_char_p]),
    ("hived_ingest_intern", C.c_int32, [_P, C.c_int32, C.c_char_p, C.c_int32, C.c_int32]),
    ("hived_ingest_lookup", C.c_int32, [_P, C.c_int32, C.c_char_p, C.c_int32]),
    ("hived_ingest_release", C.c_int32, [_P, C.c_int32, C.c_char_p, C.c_int32]),
    ("hived_ingest_pod_spec_yaml", C.c_int,
     [_P, C.c_char_p, C.c_int64, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(_cabi.PodSpec)]),
    ("hived_ingest_last_error", C.c_char_p, [_P]),
    ("hived_ingest_last_group_name", C.c_char_p, [_P]),
]


def bind(lib: C.CDLL) -> None:
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


class Ingest:
    GROUPS, PODS = 0, 1

    def __init__(self, lib: C.CDLL, ctx):
        bind(lib)
        self.lib = lib
        self.h = _P()
        rc = lib.hived_ingest_create(ctx, C.byref(self.h))
        if rc != 0:
            raise RuntimeError("hived_ingest_create failed (%d)" % rc)
        self.words = int(lib.hived_ingest_bitmap_words(self.h))

    def close(self):
        if self.h:
            self.lib.hived_ingest_destroy(self.h)
            self.h = None

    def new_bitmap(self):
        return (C.c_uint32 * max(1, self.words))()

    def node_names(self, names: Sequence[bytes], bitmap=None) -> Tuple[object, int, bool]:
        """names (bytes) -> (bitmap, distinct known nodes, names every node of the cluster)."""
        bm = bitmap if bitmap is not None else self.new_bitmap()
        arr = (C.c_char_p * max(1, len(names)))(*names)
        all_ = C.c_int32(0)
        cnt = self.lib.hived_ingest_node_names(self.h, arr, len(names), bm, C.byref(all_))
        return bm, int(cnt), bool(all_.value)

    def node_names_json(self, body: bytes, offset: int = 0, bitmap=None):
        """The JSON array at body[offset:] -> (bitmap, count, is_all, bytes consumed, answered from the cache)."""
        bm = bitmap if bitmap is not None else self.new_bitmap()
        all_, used, cached = C.c_int32(0), C.c_int64(0), C.c_int32(0)
        buf = C.c_char_p(body)
        ptr = C.cast(C.cast(buf, C.c_void_p).value + offset, C.c_char_p)
        cnt = self.lib.hived_ingest_node_names_json(self.h, ptr, len(body) - offset, bm, C.byref(all_), C.byref(used), C.byref(cached))
        if cnt < 0:
            raise ValueError(self.lib.hived_ingest_last_error(self.h).decode())
        return bm, int(cnt), bool(all_.value), int(used.value), bool(cached.value)

    def json_find(self, body: bytes, key: str) -> int:
        return int(self.lib.hived_ingest_json_find(body, len(body), key.encode()))

    def pod_spec_yaml(self, annotation: bytes, pod_name: bytes, max_groups: int, max_pods: int) -> Tuple[int, _cabi.PodSpec, str]:
        sp = _cabi.PodSpec()
        rc = self.lib.hived_ingest_pod_spec_yaml(self.h, annotation, len(annotation), pod_name, max_groups, max_pods, C.byref(sp))
        return rc, sp, (self.lib.hived_ingest_last_error(self.h) or b"").decode()

    def intern(self, kind: int, name: bytes, capacity: int) -> int:
        return int(self.lib.hived_ingest_intern(self.h, kind, name, len(name), capacity))

    def lookup(self, kind: int, name: bytes) -> int:
        return int(self.lib.hived_ingest_lookup(self.h, kind, name, len(name)))

    def release(self, kind: int, name: bytes) -> int:
        return int(self.lib.hived_ingest_release(self.h, kind, name, len(name)))
