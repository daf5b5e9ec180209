"""Pure-Python wire codec for the quickwit.search messages of this path
(quickwit-proto/protos/quickwit/search.proto). No protoc exists in the image, so requests are
assembled by hand; the C++ side has its own independent codec (csrc/proto.cpp) and the tests
cross-check the two.

Messages are plain dicts / dataclasses-free tuples to stay close to the wire:
  SortValue       ("u64"|"i64"|"f64"|"bool", value) or None
  PartialHit      dict(split_id, segment_ord, doc_id, sort_value, sort_value2) where sort_value is
                  missing (no SortByValue), None-valued (SortByValue with empty oneof) or a SortValue
"""
from __future__ import annotations

import struct
from typing import Any, Dict, List, Optional, Sequence, Tuple

ASC, DESC = 0, 1


def _varint(v: int) -> bytes:
    v &= 0xFFFFFFFFFFFFFFFF
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _tag(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _u(field: int, v: int, always=False) -> bytes:
    return _tag(field, 0) + _varint(v) if (v or always) else b""


def _b(field: int, data: bytes, always=True) -> bytes:
    return _tag(field, 2) + _varint(len(data)) + data if (data or always) else b""


def _s(field: int, s: str) -> bytes:
    return _b(field, s.encode(), always=False)


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.p = data, 0

    def done(self):
        return self.p >= len(self.d)

    def varint(self) -> int:
        v, sh = 0, 0
        while True:
            c = self.d[self.p]
            self.p += 1
            v |= (c & 0x7F) << sh
            if not c & 0x80:
                return v
            sh += 7

    def field(self):
        key = self.varint()
        f, wt = key >> 3, key & 7
        if wt == 0:
            return f, wt, self.varint()
        if wt == 1:
            v = self.d[self.p:self.p + 8]
            self.p += 8
            return f, wt, v
        if wt == 2:
            n = self.varint()
            v = self.d[self.p:self.p + n]
            self.p += n
            return f, wt, v
        if wt == 5:
            v = self.d[self.p:self.p + 4]
            self.p += 4
            return f, wt, v
        raise ValueError(f"wire type {wt}")


def _i64(v: int) -> int:
    return v - (1 << 64) if v >> 63 else v


# ---- SortByValue / PartialHit ------------------------------------------------------------------
def enc_sort_by_value(sv) -> bytes:
    if sv is None:
        return b""
    kind, v = sv
    if kind == "u64":
        return _u(1, v, True)
    if kind == "i64":
        return _u(2, v, True)
    if kind == "f64":
        return _tag(3, 1) + struct.pack("<d", v)
    if kind == "bool":
        return _u(4, int(v), True)
    raise ValueError(kind)


def dec_sort_by_value(data: bytes):
    r = _Reader(data)
    out = None
    while not r.done():
        f, wt, v = r.field()
        if f == 1:
            out = ("u64", v)
        elif f == 2:
            out = ("i64", _i64(v))
        elif f == 3:
            out = ("f64", struct.unpack("<d", v)[0])
        elif f == 4:
            out = ("bool", bool(v))
    return out


def enc_partial_hit(h: Dict[str, Any]) -> bytes:
    out = _s(2, h.get("split_id", "")) + _u(3, h.get("segment_ord", 0)) + _u(4, h.get("doc_id", 0))
    if "sort_value" in h:
        out += _b(10, enc_sort_by_value(h["sort_value"]))
    if "sort_value2" in h:
        out += _b(11, enc_sort_by_value(h["sort_value2"]))
    return out


def dec_partial_hit(data: bytes) -> Dict[str, Any]:
    r = _Reader(data)
    h: Dict[str, Any] = {"split_id": "", "segment_ord": 0, "doc_id": 0}
    while not r.done():
        f, wt, v = r.field()
        if f == 2:
            h["split_id"] = v.decode()
        elif f == 3:
            h["segment_ord"] = v
        elif f == 4:
            h["doc_id"] = v
        elif f == 10:
            h["sort_value"] = dec_sort_by_value(v)
        elif f == 11:
            h["sort_value2"] = dec_sort_by_value(v)
    return h


# ---- SearchRequest / LeafSearchRequest -------------------------------------------------------------
def enc_search_request(query_ast: str, max_hits: int = 0, start_offset: int = 0,
                       sort_fields: Sequence[Tuple[str, int]] = (), aggregation_request: Optional[str] = None,
                       start_timestamp: Optional[int] = None, end_timestamp: Optional[int] = None,
                       search_after: Optional[Dict[str, Any]] = None, index_id_patterns: Sequence[str] = ("idx",),
                       count_hits: int = 0) -> bytes:
    out = b"".join(_b(1, p.encode()) for p in index_id_patterns)
    if start_timestamp is not None:
        out += _u(4, start_timestamp, True)
    if end_timestamp is not None:
        out += _u(5, end_timestamp, True)
    out += _u(6, max_hits) + _u(7, start_offset)
    if aggregation_request is not None:
        out += _b(11, aggregation_request.encode())
    out += _s(13, query_ast)
    for name, order in sort_fields:
        out += _b(14, _s(1, name) + _u(2, order))
    if search_after is not None:
        out += _b(16, enc_partial_hit(search_after))
    out += _u(17, count_hits)
    return out


def enc_split_offsets(split_id: str, num_docs: int = 0, timestamp_start: Optional[int] = None,
                      timestamp_end: Optional[int] = None) -> bytes:
    out = _s(1, split_id)
    if timestamp_start is not None:
        out += _u(4, timestamp_start, True)
    if timestamp_end is not None:
        out += _u(5, timestamp_end, True)
    return out + _u(6, num_docs)


def enc_leaf_search_request(search_request: bytes, split_offsets: Sequence[bytes], doc_mapper_json: str,
                            index_uri: str = "ram:///idx") -> bytes:
    leaf_ref = _u(1, 0) + _u(2, 0) + b"".join(_b(3, s) for s in split_offsets)
    return _b(1, search_request) + _b(7, leaf_ref) + _b(8, doc_mapper_json.encode()) + _b(9, index_uri.encode())


# ---- LeafSearchResponse ----------------------------------------------------------------------------
def enc_leaf_search_response(num_hits: int = 0, partial_hits: Sequence[Dict[str, Any]] = (),
                             failed_splits: Sequence[Tuple[str, str, bool]] = (), num_attempted_splits: int = 0,
                             num_successful_splits: int = 0, intermediate_aggregation_result: Optional[bytes] = None,
                             resource_stats: Optional[bytes] = None) -> bytes:
    out = _u(1, num_hits) + b"".join(_b(2, enc_partial_hit(h)) for h in partial_hits)
    for err, split_id, retry in failed_splits:
        out += _b(3, _s(1, err) + _s(2, split_id) + _u(3, int(retry)))
    out += _u(4, num_attempted_splits)
    if intermediate_aggregation_result is not None:
        out += _b(6, intermediate_aggregation_result)
    out += _u(7, num_successful_splits)
    if resource_stats is not None:
        out += _b(9, resource_stats)
    return out


def enc_leaf_resource_stats(cpu_sum: int, cpu_worst: int, localexec_num_splits: int) -> bytes:
    """Only the fields the reference's merge test sets (collector.rs:1976-1987)."""
    return _u(3, localexec_num_splits) + _b(5, _u(9, cpu_worst)) + _b(6, _u(9, cpu_sum))


def dec_leaf_resource_stats(data: bytes) -> Dict[str, Any]:
    r = _Reader(data)
    out: Dict[str, Any] = {}
    names = {1: "partial_result_cache_num_splits", 2: "partial_result_cache_num_docs", 3: "localexec_num_splits",
             4: "localexec_num_docs", 9: "wall_time_microsecs"}
    split_names = {1: "split_num_docs", 2: "input_memory_bytes", 5: "matched_num_docs", 7: "warmup_microsecs",
                   9: "cpu_search_microsecs"}
    while not r.done():
        f, wt, v = r.field()
        if f in names:
            out[names[f]] = v
        elif f in (5, 6):
            rr = _Reader(v)
            d = {}
            while not rr.done():
                ff, _, vv = rr.field()
                d[split_names.get(ff, f"f{ff}")] = vv
            out["split_resources_worst" if f == 5 else "split_resources_sum"] = d
    return out


def dec_leaf_search_response(data: bytes) -> Dict[str, Any]:
    r = _Reader(data)
    out: Dict[str, Any] = {"num_hits": 0, "partial_hits": [], "failed_splits": [], "num_attempted_splits": 0,
                           "num_successful_splits": 0, "intermediate_aggregation_result": None, "resource_stats": None}
    while not r.done():
        f, wt, v = r.field()
        if f == 1:
            out["num_hits"] = v
        elif f == 2:
            out["partial_hits"].append(dec_partial_hit(v))
        elif f == 3:
            rr = _Reader(v)
            e = {"error": "", "split_id": "", "retryable_error": False}
            while not rr.done():
                ff, _, vv = rr.field()
                if ff == 1:
                    e["error"] = vv.decode()
                elif ff == 2:
                    e["split_id"] = vv.decode()
                elif ff == 3:
                    e["retryable_error"] = bool(vv)
            out["failed_splits"].append(e)
        elif f == 4:
            out["num_attempted_splits"] = v
        elif f == 6:
            out["intermediate_aggregation_result"] = bytes(v)
        elif f == 7:
            out["num_successful_splits"] = v
        elif f == 9:
            out["resource_stats"] = dec_leaf_resource_stats(v)
    return out


def dec_lambda_responses(data: bytes) -> List[Dict[str, Any]]:
    r = _Reader(data)
    out = []
    while not r.done():
        f, wt, v = r.field()
        if f != 2:
            continue
        rr = _Reader(v)
        item: Dict[str, Any] = {"split_id": "", "response": None, "error": None}
        while not rr.done():
            ff, _, vv = rr.field()
            if ff == 1:
                item["split_id"] = vv.decode()
            elif ff == 2:
                item["response"] = dec_leaf_search_response(vv)
            elif ff == 3:
                item["error"] = vv.decode()
        out.append(item)
    return out
