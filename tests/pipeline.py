"""CPU reference pipeline for tests: the oracle stands where the GPU engine stands, everything
else (plan compilation, response building, merging, finalisation) is the product's host code."""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional, Sequence

from quickwit_b200 import proto, service
from oracle import oracle as O


def search_request(query_ast: Any, **kw) -> bytes:
    if not isinstance(query_ast, str):
        query_ast = json.dumps(query_ast)
    aggs = kw.pop("aggs", None)
    if aggs is not None and not isinstance(aggs, str):
        aggs = json.dumps(aggs)
    return proto.enc_search_request(query_ast, aggregation_request=aggs, **kw)


def leafify(req_kw: Dict[str, Any]) -> Dict[str, Any]:
    """jobs_to_leaf_request (root.rs:1775-1777): start_offset := 0, max_hits += start_offset."""
    kw = dict(req_kw)
    kw["max_hits"] = kw.get("max_hits", 0) + kw.get("start_offset", 0)
    kw["start_offset"] = 0
    return kw


def cpu_split_response(img, req_pb: bytes, doc_mapper: Dict[str, Any]) -> bytes:
    dm = json.dumps(doc_mapper)
    plan = service.compile_plan(img, req_pb, dm)
    r = O.split_search(img, plan)
    return service.build_leaf_response(img, req_pb, dm, r.num_hits, r.hits, r.cells)


def cpu_root_search(imgs: Sequence, query_ast: Any, doc_mapper: Dict[str, Any], **req_kw) -> Dict[str, Any]:
    """root_search minus fetch_docs: per-split leaf responses (oracle) -> leaf merge -> root merge ->
    aggregation finalisation. Returns the decoded merged response + final aggregation JSON."""
    leaf_pb = search_request(query_ast, **leafify(req_kw))
    root_pb = search_request(query_ast, **req_kw)
    parts = [cpu_split_response(img, leaf_pb, doc_mapper) for img in imgs]
    leaf_merged = service.merge_leaf_responses(leaf_pb, parts) if parts else b""
    root_merged = service.merge_leaf_responses(root_pb, [leaf_merged])
    out = proto.dec_leaf_search_response(root_merged)
    aggs = req_kw.get("aggs")
    if aggs is not None:
        agg_json = aggs if isinstance(aggs, str) else json.dumps(aggs)
        out["aggregations"] = json.loads(service.finalize_aggregation(agg_json, out["intermediate_aggregation_result"] or b""))
    return out


def term(field, value):
    return {"type": "term", "field": field, "value": value}


def full_text(field, text, operator="Or", mode="bool"):
    m = {"type": mode}
    if mode == "bool":
        m["operator"] = operator
    return {"type": "full_text", "field": field, "text": text, "params": {"mode": m}, "lenient": False}


def bool_(**kw):
    return {"type": "bool", **kw}


MATCH_ALL = {"type": "match_all"}
