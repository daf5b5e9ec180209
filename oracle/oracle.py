"""ctypes wrapper of the CPU oracle (oracle/qw_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never from quickwit_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libqworacle.so")


class OHit(C.Structure):
    _fields_ = [("v1", C.c_uint64), ("v2", C.c_uint64), ("doc_id", C.c_uint32),
                ("flags", C.c_uint32), ("score", C.c_float), ("reserved", C.c_uint32)]


class OCell(C.Structure):
    _fields_ = [("count", C.c_uint64), ("sum_bits", C.c_uint64),
                ("min_mapped", C.c_uint64), ("max_mapped", C.c_uint64)]


def build() -> None:
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "qw_oracle.c")
        hdr = os.path.join(_HERE, "..", "include", "qwgpu_format.h")
        if (not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
            build()
        L = C.CDLL(_LIB)
        L.qwo_split_search.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(OHit),
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(OCell),
                                       C.c_uint64, C.POINTER(C.c_uint64)]
        L.qwo_split_search_fast.argtypes = L.qwo_split_search.argtypes
        L.qwo_search_many.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                      C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        L.qwo_decode_postings.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.qwo_column_first.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OracleResult:
    def __init__(self, num_hits: int, hits: List[Tuple[int, int, int, int, float]], cells, visited: int):
        self.num_hits = num_hits
        self.hits = hits      # (doc_id, flags, v1, v2, score)
        self.cells = cells    # list of (count, sum_bits, min_mapped, max_mapped)
        self.postings_visited = visited


class ManySearch:
    """Pre-marshalled (split, plan) pairs for qwo_search_many: the CPU reference arm of bench.py runs the
    searches from a pool of C threads (no Python in the timed region)."""

    def __init__(self, imgs, plans):
        n = len(imgs)
        self.n = n
        self.bufs = [C.create_string_buffer(p, len(p)) for p in plans]
        self.imgs = (C.c_void_p * n)(*[im.ptr for im in imgs])
        self.img_lens = (C.c_uint64 * n)(*[im.nbytes for im in imgs])
        self.plans = (C.c_void_p * n)(*[C.addressof(b) for b in self.bufs])
        self.plan_lens = (C.c_uint64 * n)(*[len(p) for p in plans])
        self._keep = list(imgs)

    def run(self, threads: int, fast: bool = True):
        tot = (C.c_uint64 * 2)()
        rc = lib().qwo_search_many(self.imgs, self.img_lens, self.plans, self.plan_lens, self.n, threads, int(fast), tot)
        if rc != 0:
            raise RuntimeError(f"oracle failed: {rc}")
        return int(tot[0]), int(tot[1])  # hits, postings visited


def split_search(img, plan: bytes, max_cells: int = 1 << 22, fast: bool = False) -> OracleResult:
    """img: quickwit_b200.splitgen.SplitImage (only .ptr/.nbytes are used). fast=True: the windowed-union /
    SIMD-unpack organisation of the same arithmetic (the CPU baseline bench.py times)."""
    L = lib()
    k = int.from_bytes(plan[16:20], "little")
    num_aggs = int.from_bytes(plan[12:16], "little")
    hits = (OHit * max(k, 1))()
    ncap = max_cells if num_aggs else 1
    cells = (OCell * ncap)()
    n, nh, vis = C.c_uint32(), C.c_uint64(), C.c_uint64()
    pbuf = C.create_string_buffer(plan, len(plan))
    fn = L.qwo_split_search_fast if fast else L.qwo_split_search
    rc = fn(img.ptr, img.nbytes, C.addressof(pbuf), len(plan), hits, C.byref(n), C.byref(nh), cells, ncap, C.byref(vis))
    if rc != 0:
        raise RuntimeError(f"oracle failed: {rc}")
    out_hits = [(h.doc_id, h.flags, h.v1, h.v2, h.score) for h in hits[: n.value]]
    out_cells = []
    if num_aggs:
        from quickwit_b200 import ffi
        hdr = C.sizeof(ffi.QwPlanHeader)
        num_nodes = int.from_bytes(plan[8:12], "little")
        aggs = (ffi.QwAggNode * num_aggs).from_buffer_copy(plan[hdr + num_nodes * C.sizeof(ffi.QwPlanNode):])
        total = 0
        for a in aggs:
            c = 1 if a.kind == ffi.AGG_STATS else a.num_buckets
            p = a.parent
            while p != 0xFFFFFFFF:
                c *= aggs[p].num_buckets
                p = aggs[p].parent
            total += c
        out_cells = [(x.count, x.sum_bits, x.min_mapped, x.max_mapped) for x in cells[:total]]
    return OracleResult(nh.value, out_hits, out_cells, vis.value)


def decode_postings(img, term_ord: int, cap: int):
    docs = np.zeros(cap, dtype=np.uint32)
    tfs = np.zeros(cap, dtype=np.uint32)
    n = lib().qwo_decode_postings(img.ptr, img.nbytes, term_ord, docs.ctypes.data, tfs.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("decode failed")
    return docs[:n], tfs[:n]


def column_first(img, col: int):
    n = img.num_docs
    vals = np.zeros(n, dtype=np.uint64)
    present = np.zeros(n, dtype=np.uint8)
    if lib().qwo_column_first(img.ptr, img.nbytes, col, vals.ctypes.data, present.ctypes.data) != 0:
        raise RuntimeError("column read failed")
    return vals, present.astype(bool)
