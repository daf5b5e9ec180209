"""Seam A / B on the GPU: LeafSearchRequest bytes -> LeafSearchResponse bytes through
`qwgpu_leaf_search` / `qwgpu_invoke_leaf_search`, compared with the CPU pipeline (oracle + the
same host code) and with the reference goldens."""
import json

import numpy as np
import pytest

from quickwit_b200 import ffi, proto, service, splitgen as S
from quickwit_b200.proto import ASC, DESC
from oracle import oracle as O
from pipeline import MATCH_ALL, bool_, cpu_root_search, cpu_split_response, full_text, leafify, search_request, term
from test_oracle_goldens import (AGG_MAPPING, AGG_SPLIT1, AGG_SPLIT2, BM25_DOCS, BM25_MAPPING, SORT_DATA, SORT_MAPPING,
                                 _expected_order)

pytestmark = pytest.mark.gpu

FRACS = [0.2, 0.1, 0.05, 0.05, 0.02, 0.02, 0.01, 0.01, 0.005, 0.001]
SYNTH_MAPPING = {"field_mappings": [{"name": "body", "type": "text", "record": "freq", "fieldnorms": True},
                                    {"name": "severity_text", "type": "text", "tokenizer": "raw", "fast": True},
                                    {"name": "timestamp", "type": "datetime", "fast": True, "fast_precision": "seconds"},
                                    {"name": "tenant_id", "type": "u64", "fast": True}], "timestamp_field": "timestamp"}


def gpu_root_search(ctx, imgs, query_ast, doc_mapper, **req_kw):
    """root_search with the GPU leaf: one LeafSearchRequest over all splits, then the root merge."""
    leaf_pb = search_request(query_ast, **leafify(req_kw))
    root_pb = search_request(query_ast, **req_kw)
    offsets = [proto.enc_split_offsets(im.split_id, im.num_docs) for im in imgs]
    lreq = proto.enc_leaf_search_request(leaf_pb, offsets, json.dumps(doc_mapper))
    leaf_resp = ctx.leaf_search(lreq)
    out = proto.dec_leaf_search_response(service.merge_leaf_responses(root_pb, [leaf_resp]))
    aggs = req_kw.get("aggs")
    if aggs is not None:
        out["aggregations"] = json.loads(service.finalize_aggregation(json.dumps(aggs), out["intermediate_aggregation_result"] or b""))
    return out, proto.dec_leaf_search_response(leaf_resp)


def same(a, b):
    assert a["num_hits"] == b["num_hits"]
    assert a["partial_hits"] == b["partial_hits"]
    assert a.get("aggregations") == b.get("aggregations")


@pytest.fixture(scope="module")
def synth(gpu_ctx):
    imgs = [S.synth_split(40_000 + 1000 * i, i, FRACS, split_id=f"leaf-{i}", ts_start_secs=1_700_000_000 + 86_400 * i) for i in range(4)]
    for im in imgs:
        gpu_ctx.register_split(im)
    return imgs


def test_bm25_golden_through_leaf_search(gpu_ctx):
    img = S.build_split(BM25_DOCS, BM25_MAPPING, "bm25-gpu")
    gpu_ctx.register_split(img)
    f32 = np.float32
    out, leaf = gpu_root_search(gpu_ctx, [img], term("title", "one"), BM25_MAPPING, max_hits=1000, sort_fields=[("_score", DESC)])
    assert [(f32(h["sort_value"][1]), h["doc_id"]) for h in out["partial_hits"]] == [(f32(0.1738279), 2), (f32(0.15965714), 1), (f32(0.12343242), 0)]
    assert leaf["num_attempted_splits"] == 1 and leaf["num_successful_splits"] == 1
    assert leaf["resource_stats"]["localexec_num_splits"] == 1 and leaf["resource_stats"]["split_resources_sum"]["matched_num_docs"] == 3
    both = bool_(must=[full_text("title", "one"), full_text("nofreq", "two")])
    out, _ = gpu_root_search(gpu_ctx, [img], both, BM25_MAPPING, max_hits=1000, sort_fields=[("_score", DESC)])
    assert [(f32(h["sort_value"][1]), h["doc_id"]) for h in out["partial_hits"]] == [(f32(0.31931427), 1), (f32(0.2972603), 2), (f32(0.24686484), 0)]


def test_sort_matrix_through_leaf_search(gpu_ctx):
    docs = [{k: v for k, v in (("sort1", a), ("sort2", b)) if v is not None} for a, b in SORT_DATA]
    img = S.build_split(docs, SORT_MAPPING, "sortmatrix-leaf")
    gpu_ctx.register_split(img)
    for spec in ([], [("sort1", DESC)], [("sort1", ASC), ("sort2", DESC)], [("sort1", DESC), ("sort2", ASC)]):
        want = _expected_order(spec)
        for k in (0, 1, 5, 16):
            out, _ = gpu_root_search(gpu_ctx, [img], MATCH_ALL, SORT_MAPPING, max_hits=k, sort_fields=spec)
            assert [h["doc_id"] for h in out["partial_hits"]] == want[:k]


def test_config_queries_match_cpu_pipeline(gpu_ctx, synth):
    or10 = bool_(should=[term("body", f"t{i}") for i in range(10)])
    cases = [
        (or10, dict(max_hits=1000, sort_fields=[("_score", DESC)])),                                   # C2
        (or10, dict(max_hits=20, start_offset=30, sort_fields=[("_score", DESC)])),
        (bool_(must=[term("body", "t2")]), dict(max_hits=1000, sort_fields=[("timestamp", DESC)],
                                                start_timestamp=1_700_000_000 + 21_600, end_timestamp=1_700_000_000 + 3 * 86_400 - 21_600)),  # C3
        (term("severity_text", "ERROR"), dict(max_hits=10)),                                            # C1
        (MATCH_ALL, dict(max_hits=0, aggs={"by_sev": {"terms": {"field": "severity_text"}},
                                           "over_time": {"date_histogram": {"field": "timestamp", "fixed_interval": "1h"}}})),  # C4
        (MATCH_ALL, dict(max_hits=3, sort_fields=[("tenant_id", ASC), ("timestamp", DESC)],
                         aggs={"by_sev": {"terms": {"field": "severity_text"}, "aggs": {"over_time": {"date_histogram": {"field": "timestamp", "fixed_interval": "6h"}},
                                                                                          "tenants": {"stats": {"field": "tenant_id"}}}},
                               "tenants": {"terms": {"field": "tenant_id", "size": 5}}})),
    ]
    for ast, kw in cases:
        got, leaf = gpu_root_search(gpu_ctx, synth, ast, SYNTH_MAPPING, **kw)
        want = cpu_root_search(synth, ast, SYNTH_MAPPING, **kw)
        same(got, want)
        assert leaf["num_successful_splits"] == len(synth) and not leaf["failed_splits"]


def test_aggregation_goldens_on_gpu(gpu_ctx):
    imgs = [S.build_split(AGG_SPLIT1, AGG_MAPPING, "agg-gpu-1"), S.build_split(AGG_SPLIT2, AGG_MAPPING, "agg-gpu-2")]
    for im in imgs:
        gpu_ctx.register_split(im)
    aggs = {"date_histo": {"date_histogram": {"field": "date", "fixed_interval": "30d", "offset": "-4d"},
                           "aggs": {"response": {"stats": {"field": "response"}}}},
            "hosts": {"terms": {"field": "host"}}, "tags": {"terms": {"field": "tags"}},
            "names": {"terms": {"field": "name", "size": 1, "split_size": 1}}}
    got, _ = gpu_root_search(gpu_ctx, imgs, MATCH_ALL, AGG_MAPPING, max_hits=0, aggs=aggs)
    a = got["aggregations"]
    assert [(b["doc_count"], b["key"]) for b in a["date_histo"]["buckets"]] == [(5, 1420070400000.0), (2, 1422662400000.0)]
    assert a["date_histo"]["buckets"][0]["response"] == {"avg": 85.0, "count": 4, "max": 120.0, "min": 20.0, "sum": 340.0}
    assert [(b["doc_count"], b["key"]) for b in a["hosts"]["buckets"]] == [(4, "192.168.0.10"), (2, "192.168.0.1"), (1, "192.168.0.11")]
    assert [(b["doc_count"], b["key"]) for b in a["tags"]["buckets"]] == [(5, "nice"), (2, "cool")]
    assert a["names"] == {"buckets": [{"doc_count": 2, "key": "Fritz"}], "sum_other_doc_count": 8, "doc_count_error_upper_bound": 2}
    same(got, cpu_root_search(imgs, MATCH_ALL, AGG_MAPPING, max_hits=0, aggs=aggs))


def test_failed_split_is_reported_not_fatal(gpu_ctx, synth):
    leaf_pb = search_request(term("body", "t0"), max_hits=5)
    offsets = [proto.enc_split_offsets(synth[0].split_id, synth[0].num_docs), proto.enc_split_offsets("not-resident", 1)]
    resp = proto.dec_leaf_search_response(gpu_ctx.leaf_search(proto.enc_leaf_search_request(leaf_pb, offsets, json.dumps(SYNTH_MAPPING))))
    assert resp["num_attempted_splits"] == 2 and resp["num_successful_splits"] == 1
    assert resp["failed_splits"] == [{"error": "split `not-resident` is not resident on this GPU", "split_id": "not-resident", "retryable_error": True}]
    assert len(resp["partial_hits"]) == 5
    with pytest.raises(ffi.QwGpuError) as e:
        gpu_ctx.leaf_search(proto.enc_leaf_search_request(search_request(term("nope", "x"), max_hits=5), offsets[:1], json.dumps(SYNTH_MAPPING)))
    assert e.value.code == ffi.EINVALID_QUERY
    # a request this library does not execute (top-K above the 4096 cap) fails the same way on every node:
    # the split is reported failed but NOT retryable, so the root does not run it twice
    big = proto.enc_leaf_search_request(search_request(term("body", "t0"), max_hits=5000), offsets[:1], json.dumps(SYNTH_MAPPING))
    resp = proto.dec_leaf_search_response(gpu_ctx.leaf_search(big))
    assert resp["num_successful_splits"] == 0 and len(resp["failed_splits"]) == 1
    assert resp["failed_splits"][0]["retryable_error"] is False and resp["failed_splits"][0]["split_id"] == synth[0].split_id


def test_invoke_leaf_search_per_split_results(gpu_ctx, synth):
    leaf_pb = search_request(bool_(should=[term("body", "t0"), term("body", "t3")]), max_hits=7, sort_fields=[("_score", DESC)])
    offsets = [proto.enc_split_offsets(im.split_id, im.num_docs) for im in synth] + [proto.enc_split_offsets("ghost", 1)]
    results = proto.dec_lambda_responses(gpu_ctx.invoke_leaf_search(proto.enc_leaf_search_request(leaf_pb, offsets, json.dumps(SYNTH_MAPPING))))
    assert [r["split_id"] for r in results] == [im.split_id for im in synth] + ["ghost"]
    assert results[-1]["error"] and results[-1]["response"] is None
    for im, r in zip(synth, results):
        want = proto.dec_leaf_search_response(cpu_split_response(im, leaf_pb, SYNTH_MAPPING))
        assert r["response"]["num_hits"] == want["num_hits"] and r["response"]["partial_hits"] == want["partial_hits"]


def test_concurrent_leaf_searches_match_sequential(gpu_ctx, synth):
    """qwgpu_leaf_search is callable from several host threads at once (one stream + staging slot per
    in-flight call); interleaved requests must return what they return alone."""
    from concurrent.futures import ThreadPoolExecutor
    offsets = [proto.enc_split_offsets(im.split_id, im.num_docs) for im in synth]
    dm = json.dumps(SYNTH_MAPPING)
    reqs = []
    for q in range(8):
        ast = bool_(should=[term("body", f"t{(q + i) % 10}") for i in range(1 + q % 4)]) if q % 2 == 0 else bool_(must=[term("body", f"t{q % 5}")])
        kw = dict(max_hits=50 + 10 * q, sort_fields=[("_score", DESC)] if q % 2 == 0 else [("timestamp", DESC)])
        if q == 3:
            kw["aggs"] = {"by_sev": {"terms": {"field": "severity_text"}}}
        reqs.append(proto.enc_leaf_search_request(search_request(ast, **kw), offsets, dm))

    def hits(resp):
        d = proto.dec_leaf_search_response(resp)
        return d["num_hits"], d["partial_hits"], d["intermediate_aggregation_result"]

    want = [hits(gpu_ctx.leaf_search(r)) for r in reqs]
    with ThreadPoolExecutor(max_workers=4) as ex:
        for _ in range(5):
            got = list(ex.map(lambda r: hits(gpu_ctx.leaf_search(r)), reqs))
            assert got == want


def test_more_callers_than_the_admission_gate_admits(gpu_ctx, synth):
    """40 host threads against the default gate of 16 searches in flight: the surplus callers wait inside the
    library and every response equals the sequential one."""
    import threading
    offsets = [proto.enc_split_offsets(im.split_id, im.num_docs) for im in synth]
    dm = json.dumps(SYNTH_MAPPING)
    reqs = [proto.enc_leaf_search_request(search_request(bool_(should=[term("body", f"t{(q + i) % 10}") for i in range(1 + q % 3)]),
                                                         max_hits=20 + q, sort_fields=[("_score", DESC)]), offsets, dm) for q in range(5)]
    key = lambda resp: (lambda d: (d["num_hits"], d["partial_hits"]))(proto.dec_leaf_search_response(resp))
    want = [key(gpu_ctx.leaf_search(r)) for r in reqs]
    got, errs = [[] for _ in range(40)], []

    def worker(t):
        try:
            for k in range(3):
                q = (t + k) % len(reqs)
                got[t].append((q, gpu_ctx.leaf_search(reqs[q])))
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    ths = [threading.Thread(target=worker, args=(t,)) for t in range(40)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errs, errs
    assert all(key(resp) == want[q] for g in got for q, resp in g) and sum(len(g) for g in got) == 120


def test_pre_search_pruning_keeps_the_response(gpu_ctx):
    """a16: with the split metadata in the request (time ranges, doc counts) match-all top-K requests demote the
    splits that cannot reach the top to count-only / metadata-count requests (leaf.rs:1141-1242, 525-528).
    The merged response must not change; the device sees fewer splits."""
    mapping = {"field_mappings": [{"name": "ts", "type": "datetime", "fast": True}, {"name": "body", "type": "text"}], "timestamp_field": "ts"}
    t0 = 1_700_000_000
    imgs, ranges = [], []
    for i in range(5):
        lo, hi = t0 + 100 * i, t0 + 100 * i + 60 + (80 if i == 2 else 0)   # split 2 overlaps split 3
        docs = [{"ts": lo + (j * 7) % (hi - lo + 1), "body": "hello"} for j in range(40)]
        imgs.append(S.build_split(docs, mapping, f"prune-{i}"))
        ranges.append((min(d["ts"] for d in docs), max(d["ts"] for d in docs)))
        gpu_ctx.register_split(imgs[-1])
    offsets = [proto.enc_split_offsets(im.split_id, im.num_docs, a, b) for im, (a, b) in zip(imgs, ranges)]
    for kw in (dict(max_hits=5, sort_fields=[("ts", DESC)]), dict(max_hits=50, sort_fields=[("ts", ASC)]), dict(max_hits=7),
               dict(max_hits=45, start_offset=40), dict(max_hits=0)):
        lreq = proto.enc_leaf_search_request(search_request(MATCH_ALL, **leafify(kw)), offsets, json.dumps(mapping))
        plan = service.optimize_leaf_request(lreq)
        leaf = proto.dec_leaf_search_response(gpu_ctx.leaf_search(lreq))
        got = proto.dec_leaf_search_response(service.merge_leaf_responses(search_request(MATCH_ALL, **kw), [gpu_ctx.leaf_search(lreq)]))
        want = cpu_root_search(imgs, MATCH_ALL, mapping, **kw)   # no pruning on this side
        assert got["num_hits"] == want["num_hits"] == 200 and got["partial_hits"] == want["partial_hits"], kw
        assert leaf["num_attempted_splits"] == 5 and leaf["num_successful_splits"] == 5 and not leaf["failed_splits"]
        n_meta = sum(r["metadata_count"] for r in plan)
        assert n_meta >= 1, (kw, plan)
        on_device = leaf["resource_stats"]["localexec_num_splits"] if leaf["resource_stats"] else 0
        assert on_device == 5 - n_meta, (kw, plan)
    # CountHits::Underestimate (= 1): the demoted splits are not searched at all (simplify_search_request returns None,
    # leaf.rs:1399-1433): same hits, a smaller count, fewer splits attempted
    kw = dict(max_hits=5, sort_fields=[("ts", DESC)])
    lreq = proto.enc_leaf_search_request(search_request(MATCH_ALL, count_hits=1, **kw), offsets, json.dumps(mapping))
    plan = service.optimize_leaf_request(lreq)
    leaf_bytes = gpu_ctx.leaf_search(lreq)
    leaf = proto.dec_leaf_search_response(leaf_bytes)
    got = proto.dec_leaf_search_response(service.merge_leaf_responses(search_request(MATCH_ALL, count_hits=1, **kw), [leaf_bytes]))
    n_skip = sum(r["skipped"] for r in plan)
    assert n_skip >= 1 and not any(r["metadata_count"] for r in plan)
    assert got["partial_hits"] == cpu_root_search(imgs, MATCH_ALL, mapping, **kw)["partial_hits"]
    assert leaf["num_hits"] == 40 * (5 - n_skip) and leaf["num_attempted_splits"] == 5 - n_skip and not leaf["failed_splits"]
    # a non-resident split that only has to be counted is answered from its metadata
    offs = offsets + [proto.enc_split_offsets("prune-ghost", 123, t0 - 500, t0 - 400)]
    lreq = proto.enc_leaf_search_request(search_request(MATCH_ALL, max_hits=5, sort_fields=[("ts", DESC)]), offs, json.dumps(mapping))
    leaf = proto.dec_leaf_search_response(gpu_ctx.leaf_search(lreq))
    assert leaf["num_hits"] == 323 and not leaf["failed_splits"] and leaf["num_successful_splits"] == 6


def test_residency_budget_lru_and_background_upload():
    """Residency manager (qwgpu.h): a byte budget evicts the least recently SEARCHED splits, an evicted split
    shows up as a retryable failed split, background uploads are waited for by the searches that need them."""
    ctx = service.SearcherContext(0)
    imgs = [S.synth_split(30_000, 100 + i, FRACS[:4], split_id=f"lru-{i}", ts_start_secs=1_700_000_000) for i in range(5)]
    one = None
    for im in imgs[:3]:
        ctx.register_split(im)
        one = ctx.resident_bytes() if one is None else one
    assert ctx.residency_info()["num_splits"] == 3 and ctx.resident_bytes() >= 2.9 * one
    leaf = lambda ids: proto.dec_leaf_search_response(ctx.leaf_search(proto.enc_leaf_search_request(
        search_request(term("body", "t0"), max_hits=3), [proto.enc_split_offsets(i, 30_000) for i in ids], json.dumps(SYNTH_MAPPING))))
    leaf(["lru-0"])                                    # lru-0 is now the most recently used, lru-1 the least
    ctx.set_residency_budget(int(3.5 * one))
    ctx.register_split(imgs[3])                        # needs room: evicts lru-1
    assert [ctx.is_resident(f"lru-{i}") for i in range(4)] == [True, False, True, True]
    assert ctx.residency_info()["evictions"] == 1 and ctx.resident_bytes() <= int(3.5 * one)
    resp = leaf(["lru-0", "lru-1"])
    assert resp["num_successful_splits"] == 1 and resp["failed_splits"][0]["split_id"] == "lru-1" and resp["failed_splits"][0]["retryable_error"]
    # background upload of two splits: the search that names them waits for the upload and sees every doc
    ctx.register_split_async(imgs[4])
    ctx.register_split_async(imgs[1])                  # back in: evicts the least recently used of the rest
    resp = leaf(["lru-4", "lru-1"])
    assert resp["num_successful_splits"] == 2 and not resp["failed_splits"]
    want = [proto.dec_leaf_search_response(cpu_split_response(imgs[i], search_request(term("body", "t0"), max_hits=3), SYNTH_MAPPING)) for i in (4, 1)]
    assert resp["num_hits"] == sum(w["num_hits"] for w in want)
    ctx.wait_split("lru-4"); ctx.wait_split("lru-1")
    info = ctx.residency_info()
    assert info["resident_bytes"] <= info["budget_bytes"] and info["num_splits"] == 3 and info["evictions"] == 3
    with pytest.raises(ffi.QwGpuError):
        ctx.set_residency_budget(one // 2) or ctx.register_split(imgs[0])   # a split larger than the whole budget


def test_wildcard_queries_on_the_device(gpu_ctx):
    """Wildcard queries compile (on the host) into a filter over the union of the matching dictionary terms: the device
    result must equal the oracle's for the same plan, through seam C and through qwgpu_leaf_search."""
    mapping = {"field_mappings": [{"name": "body", "type": "text", "record": "freq", "fieldnorms": True},
                                  {"name": "tag", "type": "text", "tokenizer": "raw"}, {"name": "n", "type": "u64", "fast": True}]}
    words = ["alpha", "alpine", "beta", "betamax", "gamma", "Alphabet", "al", "delta", "epsilon", "zeta"]
    docs = [{"body": f"{words[i % 10]} {words[(i * 7) % 10]} filler", "tag": ["Prod-EU", "prod-us", "Dev", "staging"][i % 4], "n": i} for i in range(3000)]
    img = S.build_split(docs, mapping, "wc-gpu-0")
    gpu_ctx.register_split(img)
    try:
        dm = json.dumps(mapping)
        wc = lambda field, value, **kw: {"type": "wildcard", "field": field, "value": value, **kw}
        for ast, kw in [(wc("body", "al*"), dict(max_hits=50)), (wc("body", "*eta*"), dict(max_hits=20, sort_fields=[("n", ASC)])),
                        (wc("tag", "prod*", case_insensitive=True), dict(max_hits=0)),
                        (bool_(must=[term("body", "filler")], filter=[wc("body", "?eta")]), dict(max_hits=30, sort_fields=[("_score", DESC)]))]:
            sreq = search_request(ast, **kw)
            plan = service.compile_plan(img, sreq, dm)
            got = gpu_ctx.split_search([img.split_id], [plan])[0]
            want = O.split_search(img, plan)
            assert got.num_hits == want.num_hits > 0 and [h[:4] for h in got.hits] == [h[:4] for h in want.hits], ast
            lreq = proto.enc_leaf_search_request(sreq, [proto.enc_split_offsets(img.split_id, img.num_docs)], dm)
            leaf = proto.dec_leaf_search_response(gpu_ctx.leaf_search(lreq))
            assert leaf["num_hits"] == want.num_hits and not leaf["failed_splits"]
    finally:
        gpu_ctx.unregister_split(img.split_id)
