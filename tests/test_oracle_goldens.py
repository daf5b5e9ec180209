"""Pins the CPU oracle (and the host-side compile / merge / finalise code around it) against every
golden vector the reference's own tests hold for this path. CPU only."""
import json

import numpy as np
import time

import pytest

from quickwit_b200 import ffi, proto, service, splitgen as S
from quickwit_b200.proto import ASC, DESC
from oracle import oracle as O
from pipeline import MATCH_ALL, bool_, cpu_root_search, cpu_split_response, full_text, search_request, term


# ---- quickwit-search/src/tests.rs:600-691 test_sort_bm25 ---------------------------------------------
BM25_MAPPING = {"field_mappings": [
    {"name": "title", "type": "text", "record": "freq", "fieldnorms": True},
    {"name": "body", "type": "text", "record": "freq", "fieldnorms": True},
    {"name": "nofreq", "type": "text", "record": "basic", "fieldnorms": True},
    {"name": "nofreq_nofieldnorms", "type": "text", "fieldnorms": False}]}
BM25_DOCS = [{"title": "one pad", "nofreq": "two pad"}, {"title": "one", "nofreq": "two"},
             {"title": "one one", "nofreq": "two two"}]


def _scores(res):
    return [(np.float32(h["sort_value"][1]), h["doc_id"]) for h in res["partial_hits"]]


def test_sort_bm25_exact_f32_scores():
    img = S.build_split(BM25_DOCS, BM25_MAPPING, "bm25")
    run = lambda ast: _scores(cpu_root_search([img], ast, BM25_MAPPING, max_hits=1000, sort_fields=[("_score", DESC)]))
    f32 = np.float32
    assert run(term("title", "one")) == [(f32(0.1738279), 2), (f32(0.15965714), 1), (f32(0.12343242), 0)]
    # tf forced to 1 for `record: basic`; the tie is broken doc id descending
    assert run(term("nofreq", "two")) == [(f32(0.15965714), 1), (f32(0.12343242), 2), (f32(0.12343242), 0)]
    # user text "title:one nofreq:two" with default operator AND: both must, scores add
    both = bool_(must=[full_text("title", "one"), full_text("nofreq", "two")])
    assert run(both) == [(f32(0.31931427), 1), (f32(0.2972603), 2), (f32(0.24686484), 0)]


# ---- quickwit-search/src/collector.rs:1391-1637 test_single_split_sorting ----------------------------
SORT_DATA = [(2, 1), (0, 1), (1, 1), (0, 0), (None, 1), (None, 2), (2, 1), (1, 2), (0, None), (None, 0), (2, 0), (2, 2),
             (0, 2), (2, None), (None, None), (1, 0), (1, None)]
SORT_MAPPING = {"field_mappings": [{"name": "sort1", "type": "u64", "fast": True}, {"name": "sort2", "type": "u64", "fast": True}]}


def _sort_img(split_id="fake_split_id"):
    docs = [{k: v for k, v in (("sort1", a), ("sort2", b)) if v is not None} for a, b in SORT_DATA]
    return S.build_split(docs, SORT_MAPPING, split_id)


def _expected_order(spec):
    """The comparator table of the reference test: None last in both directions, doc id direction
    follows the first sort order."""
    def key(doc):
        i, (a, b) = doc
        parts = []
        for name, v in (("sort1", a), ("sort2", b)):
            order = next((o for f, o in spec if f == name), None)
            if order is None:
                continue
            parts.append((1, 0) if v is None else (0, -v if order == DESC else v))
        first_order = spec[0][1] if spec else DESC
        parts.append((0, -i if first_order == DESC else i))
        return parts
    return [i for i, _ in sorted(enumerate(SORT_DATA), key=key)]


@pytest.mark.parametrize("spec", [[], [("sort1", DESC)], [("sort1", ASC)], [("sort1", DESC), ("sort2", DESC)],
                                  [("sort1", ASC), ("sort2", DESC)], [("sort1", DESC), ("sort2", ASC)],
                                  [("sort1", ASC), ("sort2", ASC)]])
def test_single_split_sorting(spec):
    img = _sort_img()
    want = _expected_order(spec)
    for k in range(0, len(SORT_DATA)):
        res = cpu_root_search([img], MATCH_ALL, SORT_MAPPING, max_hits=k, sort_fields=spec)
        assert [h["doc_id"] for h in res["partial_hits"]] == want[:k], (spec, k)
        assert res["num_hits"] == len(SORT_DATA)


# ---- collector.rs:1639-1771 test_search_after --------------------------------------------------------------
def test_search_after():
    img = _sort_img()
    spec = [("sort1", DESC), ("sort2", ASC)]
    order = _expected_order(spec)
    for i, doc in enumerate(order):
        a, b = SORT_DATA[doc]
        sa = {"split_id": "fake_split_id", "segment_ord": 0, "doc_id": doc,
              "sort_value": ("u64", a) if a is not None else None, "sort_value2": ("u64", b) if b is not None else None}
        res = cpu_root_search([img], MATCH_ALL, SORT_MAPPING, max_hits=1000, sort_fields=spec, search_after=sa)
        assert res["num_hits"] == len(SORT_DATA)  # hits removed by search_after are still counted
        assert [h["doc_id"] for h in res["partial_hits"]] == order[i + 1:]
    # elimination on split id (sort by _shard_doc desc)
    sa = {"split_id": "fake_split_id2", "segment_ord": 0, "doc_id": 5}
    for split_id, n in (("fake_split_id1", 17), ("fake_split_id2", 5), ("fake_split_id3", 0)):
        res = cpu_root_search([_sort_img(split_id)], MATCH_ALL, SORT_MAPPING, max_hits=1000, sort_fields=[("_shard_doc", DESC)], search_after=sa)
        assert res["num_hits"] == 17 and len(res["partial_hits"]) == n, split_id


# ---- collector.rs:1332-1389 merge_partial_hits / quickwit-common binary_heap.rs:201-275 ----------------------
def _merge_hits(hits, order, k):
    req = proto.enc_search_request("{}", max_hits=k, sort_fields=[("f", order)])
    parts = [proto.enc_leaf_search_response(partial_hits=[h], num_attempted_splits=1, num_successful_splits=1) for h in hits]
    return proto.dec_leaf_search_response(service.merge_leaf_responses(req, parts))["partial_hits"]


def test_merge_partial_hits_no_tie():
    mk = lambda v: {"split_id": "split1", "segment_ord": 0, "doc_id": 0, "sort_value": ("u64", v)}
    assert _merge_hits([mk(1), mk(3), mk(2)], ASC, 2) == [mk(1), mk(2)]


def test_merge_partial_hits_with_tie():
    mk = lambda s: {"split_id": f"split_{s}", "segment_ord": 0, "doc_id": 0, "sort_value": ("u64", 0)}
    assert _merge_hits([mk(1), mk(3), mk(2)], DESC, 2) == [mk(3), mk(2)]
    assert _merge_hits([mk(1), mk(3), mk(2)], ASC, 2) == [mk(1), mk(2)]


# ---- collector.rs:1793-2059 test_merge_collectors ---------------------------------------------------------------
def test_merge_collectors():
    hit = lambda split, doc, v: {"split_id": split, "segment_ord": 0, "doc_id": doc, "sort_value": ("i64", v)}
    req = lambda order: proto.enc_search_request("{}", max_hits=2, sort_fields=[("timestamp", order)])
    single = proto.enc_leaf_search_response(1234, [hit("1", 123, 1234)], num_attempted_splits=3, num_successful_splits=3)
    out = proto.dec_leaf_search_response(service.merge_leaf_responses(req(DESC), [single]))
    assert (out["num_hits"], out["partial_hits"], out["num_attempted_splits"], out["num_successful_splits"]) == \
        (1234, [hit("1", 123, 1234)], 3, 3)
    stats = lambda cpu: proto.enc_leaf_resource_stats(cpu, cpu, 1)
    a = lambda st: proto.enc_leaf_search_response(1234, [hit("1", 123, 1234), hit("1", 125, 1236)], num_attempted_splits=3,
                                                  num_successful_splits=3, resource_stats=st)
    b = lambda st: proto.enc_leaf_search_response(10, [hit("2", 3, 1235)], [("fake error", "3", True)], 2, 1, resource_stats=st)
    out = proto.dec_leaf_search_response(service.merge_leaf_responses(req(DESC), [a(None), b(None)]))
    assert out["num_hits"] == 1244 and out["partial_hits"] == [hit("1", 125, 1236), hit("2", 3, 1235)]
    assert out["failed_splits"] == [{"error": "fake error", "split_id": "3", "retryable_error": True}]
    assert (out["num_attempted_splits"], out["num_successful_splits"], out["resource_stats"]) == (5, 4, None)
    out = proto.dec_leaf_search_response(service.merge_leaf_responses(req(ASC), [a(stats(100)), b(stats(50))]))
    assert out["partial_hits"] == [hit("1", 123, 1234), hit("2", 3, 1235)]
    rs = out["resource_stats"]
    assert rs["split_resources_sum"]["cpu_search_microsecs"] == 150
    assert rs["split_resources_worst"]["cpu_search_microsecs"] == 100 and rs["localexec_num_splits"] == 2


def test_merge_empty_intermediate_aggregation_result():
    req = proto.enc_search_request("{}", aggregation_request='{"avg_price": {"avg": {"field": "price"}}}')
    out = proto.dec_leaf_search_response(service.merge_leaf_responses(req, []))
    final = json.loads(service.finalize_aggregation('{"avg_price": {"avg": {"field": "price"}}}', out["intermediate_aggregation_result"]))
    assert final == {"avg_price": {"value": None}}


# ---- tests.rs:264-315 test_single_node_several_splits ----------------------------------------------------------
def test_several_splits_default_order():
    mapping = {"field_mappings": [{"name": "title", "type": "text"}, {"name": "body", "type": "text"}, {"name": "url", "type": "text"},
                                  {"name": "owner", "type": "text", "tokenizer": "raw"}]}
    docs = [{"title": "snoopy", "body": "Snoopy is an anthropomorphic beagle[5] in the comic strip...", "url": "http://snoopy"},
            {"title": "beagle", "body": "The beagle is a breed of small scent hound, similar in appearance to the much larger foxhound.", "url": "http://beagle"}]
    imgs = [S.build_split(docs, mapping, f"split-{i:02d}") for i in range(10)]
    res = cpu_root_search(imgs, full_text("body", "beagle"), mapping, max_hits=6)
    assert res["num_hits"] == 20 and len(res["partial_hits"]) == 6
    keys = [(h["split_id"], h["doc_id"]) for h in res["partial_hits"]]
    assert keys == sorted(keys, reverse=True) and len(set(keys)) == 6
    assert keys[0] == ("split-09", 1) and keys[1] == ("split-09", 0)  # "breed" doc first, then "Snoopy"


# ---- tests.rs:318-428 test_single_node_filtering -------------------------------------------------------------------
def test_filtering_term_and_timestamp_range():
    mapping = {"field_mappings": [{"name": "body", "type": "text"},
                                  {"name": "ts", "type": "datetime", "fast": True, "input_formats": ["rfc3339", "unix_timestamp"]},
                                  {"name": "owner", "type": "text", "tokenizer": "raw"}], "timestamp_field": "ts"}
    start = 1_700_000_000
    docs = [{"body": f"info @ t:{i + 1}", "ts": start + i + 1} for i in range(30)]
    img = S.build_split(docs, mapping, "filtering")
    res = cpu_root_search([img], full_text("body", "info"), mapping, max_hits=15, sort_fields=[("ts", DESC)],
                          start_timestamp=start + 10, end_timestamp=start + 20)
    assert res["num_hits"] == 10 and len(res["partial_hits"]) == 10
    assert res["partial_hits"][0]["doc_id"] == 18 and res["partial_hits"][9]["doc_id"] == 9  # t:19 ... t:10
    assert res["partial_hits"][0]["sort_value"] == ("i64", (start + 19) * 10**9)
    res = cpu_root_search([img], full_text("body", "info"), mapping, max_hits=25, sort_fields=[("ts", DESC)], end_timestamp=start + 20)
    assert res["num_hits"] == 19 and res["partial_hits"][0]["doc_id"] == 18 and res["partial_hits"][18]["doc_id"] == 0
    with pytest.raises(ffi.QwGpuError) as e:
        cpu_root_search([img], bool_(must=[full_text("tag", "foo"), full_text("body", "info")]), mapping, max_hits=25)
    assert e.value.code == ffi.EINVALID_QUERY and "invalid query: field does not exist: `tag`" in e.value.msg


# ---- tests.rs:1511-1637 test_single_node_range_queries --------------------------------------------------------------
def test_range_queries():
    mapping = {"field_mappings": [{"name": "datetime", "type": "datetime", "fast": True},
                                  {"name": "log_level", "type": "text", "tokenizer": "raw", "fast": True},
                                  {"name": "status_code", "type": "u64", "fast": True},
                                  {"name": "latency", "type": "f64", "fast": True},
                                  {"name": "error_code", "type": "i64", "fast": True}]}
    docs = [{"datetime": f"2023-01-10T{15 + i}:13:35Z", "log_level": lvl, "status_code": sc, "latency": lat, "error_code": ec}
            for i, (lvl, sc, lat, ec) in enumerate([("DEBUG", 200, 0.5, -10), ("INFO", 201, 1.5, 0), ("WARN", 404, 2.5, 10),
                                                    ("ERROR", 500, 3.5, 20), ("FATAL", 503, 4.5, 30)])]
    img = S.build_split(docs, mapping, "ranges")
    rq = lambda field, lo=None, hi=None: {"type": "range", "field": field, "lower_bound": lo or "Unbounded", "upper_bound": hi or "Unbounded"}
    count = lambda ast: cpu_root_search([img], ast, mapping, max_hits=10)["num_hits"]
    assert count(rq("datetime", {"Included": "2023-01-10T15:13:35Z"}, {"Excluded": "2023-01-10T17:13:35Z"})) == 2
    assert count(rq("status_code", {"Included": 400}, {"Included": 503})) == 3
    assert count(rq("status_code", {"Excluded": 200}, {"Excluded": 503})) == 3
    assert count(rq("latency", {"Included": 1.5}, {"Excluded": 4.5})) == 3
    assert count(rq("error_code", {"Included": -10}, {"Included": 20})) == 4
    assert count(rq("log_level", {"Included": "ERROR"}, {"Included": "INFO"})) == 3  # ERROR, FATAL, INFO
    assert count(rq("status_code", {"Included": "201"})) == 4  # numbers passed as strings (JsonLiteral::String)


# ---- rest-api-tests/scenarii/aggregations ---------------------------------------------------------------------------------
AGG_MAPPING = {"mode": "dynamic", "dynamic_mapping": {"tokenizer": "default", "fast": True},
               "field_mappings": [{"name": "date", "type": "datetime", "input_formats": ["rfc3339"], "fast_precision": "seconds", "fast": True},
                                  {"name": "high_prec_test", "type": "u64", "fast": True}]}
AGG_SPLIT1 = [{"name": "Albert", "response": 100, "id": 1, "date": "2015-01-01T12:10:30Z", "host": "192.168.0.10", "tags": ["nice"]},
              {"name": "Fred", "response": 100, "id": 3, "date": "2015-01-01T12:10:30Z", "host": "192.168.0.1", "tags": ["nice"]},
              {"name": "Manfred", "response": 120, "id": 13, "date": "2015-01-11T12:10:30Z", "host": "192.168.0.11", "tags": ["nice"]},
              {"name": "Horst", "id": 2, "date": "2015-01-01T11:11:30Z", "host": "192.168.0.10", "tags": ["nice", "cool"]},
              {"name": "Fritz", "response": 30, "id": 5, "host": "192.168.0.1", "tags": ["nice", "cool"]}]
AGG_SPLIT2 = [{"name": "Fritz", "high_prec_test": 1769070189829214200, "response": 30, "id": 0},
              {"name": "Fritz", "response": 30, "id": 0},
              {"name": "Holger", "response": 30, "id": 4, "date": "2015-02-06T00:00:00Z", "host": "192.168.0.10"},
              {"name": "Werner", "response": 20, "id": 5, "date": "2015-01-02T00:00:00Z", "host": "192.168.0.10"},
              {"name": "Bernhard", "response": 130, "id": 14, "date": "2015-02-16T00:00:00Z"}]


@pytest.fixture(scope="module")
def agg_splits():
    return [S.build_split(AGG_SPLIT1, AGG_MAPPING, "agg-1"), S.build_split(AGG_SPLIT2, AGG_MAPPING, "agg-2")]


def _aggs(imgs, aggs, query=MATCH_ALL):
    return cpu_root_search(imgs, query, AGG_MAPPING, max_hits=0, aggs=aggs)["aggregations"]


def test_agg_date_histogram(agg_splits):
    dh = {"field": "date", "fixed_interval": "30d", "offset": "-4d"}
    got = _aggs(agg_splits, {"date_histo": {"date_histogram": dh}})
    assert got == {"date_histo": {"buckets": [
        {"doc_count": 5, "key": 1420070400000.0, "key_as_string": "2015-01-01T00:00:00Z"},
        {"doc_count": 2, "key": 1422662400000.0, "key_as_string": "2015-01-31T00:00:00Z"}]}}
    got = _aggs(agg_splits, {"date_histo": {"date_histogram": dict(dh, extended_bounds={"min": 1420070400000, "max": 1425254400000})}})
    assert [(b["doc_count"], b["key"], b["key_as_string"]) for b in got["date_histo"]["buckets"]] == [
        (5, 1420070400000.0, "2015-01-01T00:00:00Z"), (2, 1422662400000.0, "2015-01-31T00:00:00Z"), (0, 1425254400000.0, "2015-03-02T00:00:00Z")]
    nested = {"date_histo": {"date_histogram": dh, "aggs": {"response": {"stats": {"field": "response"}}}}}
    got = _aggs(agg_splits, nested)["date_histo"]["buckets"]
    assert got[0]["response"] == {"avg": 85.0, "count": 4, "max": 120.0, "min": 20.0, "sum": 340.0} and got[0]["doc_count"] == 5
    assert got[1]["response"] == {"avg": 80.0, "count": 2, "max": 130.0, "min": 30.0, "sum": 160.0}
    exists = bool_(must=[{"type": "field_presence", "field": "response"}])
    got = _aggs(agg_splits, nested, query=exists)["date_histo"]["buckets"]
    assert (got[0]["doc_count"], got[1]["doc_count"]) == (4, 2) and got[0]["response"]["sum"] == 340.0


def test_stats_over_a_datetime_column_does_not_overflow():
    """avg / sum / stats over a nanosecond timestamp column: the typed values are ~1.4e18 each, so a
    wrapping 64-bit sum of them overflows after half a dozen docs; the cells carry the sum of the raw
    offsets instead and the host rebuilds the exact sum in 128 bits (QwAggCell, qwgpu_format.h).
    Reference semantics: docs/reference/aggregation.md "stats" (f64 sum / count / min / max / avg)."""
    n = 40
    secs = [1_420_070_400 + 86_400 * i + 7 * (i % 5) for i in range(n)]
    docs = [{"id": i, "date": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime(t))} for i, t in enumerate(secs)]
    img = S.build_split(docs, AGG_MAPPING, "dates")
    got = _aggs([img], {"d": {"stats": {"field": "date"}}, "a": {"avg": {"field": "date"}}})
    ns = [t * 1_000_000_000 for t in secs]
    # tantivy reports datetime metrics in the column's f64 space (nanoseconds)
    assert got["d"]["count"] == n and got["d"]["min"] == float(min(ns)) and got["d"]["max"] == float(max(ns))
    assert got["d"]["sum"] == float(sum(ns))          # exactly rounded, not a wrapped i64
    assert got["d"]["avg"] == float(sum(ns)) / n and got["a"]["value"] == got["d"]["avg"]
    # two splits: the merged sum is the sum of the per-split f64 sums
    a, b = S.build_split(docs[:17], AGG_MAPPING, "dates-a"), S.build_split(docs[17:], AGG_MAPPING, "dates-b")
    two = _aggs([a, b], {"d": {"stats": {"field": "date"}}})
    assert two["d"]["count"] == n and two["d"]["sum"] == float(sum(ns[:17])) + float(sum(ns[17:]))


def test_agg_range_and_histogram(agg_splits):
    rng = {"my_range": {"range": {"field": "response", "ranges": [{"to": 50, "key": "fast"}, {"from": 50, "to": 80, "key": "medium"}, {"from": 80, "key": "slow"}]}}}
    assert _aggs(agg_splits, rng) == {"my_range": {"buckets": [
        {"doc_count": 5, "key": "fast", "to": 50.0}, {"doc_count": 0, "from": 50.0, "key": "medium", "to": 80.0},
        {"doc_count": 4, "from": 80.0, "key": "slow"}]}}
    hist = _aggs(agg_splits, {"metrics": {"histogram": {"field": "response", "interval": 50}}})
    assert hist == {"metrics": {"buckets": [{"doc_count": 5, "key": 0.0}, {"doc_count": 0, "key": 50.0}, {"doc_count": 4, "key": 100.0}]}}
    empty = S.build_split([], AGG_MAPPING, "empty")
    assert _aggs([empty], {"metrics": {"histogram": {"field": "response", "interval": 50}}}) == {"metrics": {"buckets": []}}


def test_agg_terms(agg_splits):
    got = _aggs(agg_splits, {"hosts": {"terms": {"field": "host"}}, "tags": {"terms": {"field": "tags"}}})
    assert got["hosts"] == {"buckets": [{"doc_count": 4, "key": "192.168.0.10"}, {"doc_count": 2, "key": "192.168.0.1"},
                                        {"doc_count": 1, "key": "192.168.0.11"}], "doc_count_error_upper_bound": 0, "sum_other_doc_count": 0}
    assert got["tags"] == {"buckets": [{"doc_count": 5, "key": "nice"}, {"doc_count": 2, "key": "cool"}],
                           "doc_count_error_upper_bound": 0, "sum_other_doc_count": 0}
    for alias in ("split_size", "segment_size", "shard_size"):
        got = _aggs(agg_splits, {"names": {"terms": {"field": "name", "size": 1, alias: 1}}})["names"]
        # one "Fritz" is cut off by split_size=1 on the split where every name occurs once
        assert got == {"buckets": [{"doc_count": 2, "key": "Fritz"}], "sum_other_doc_count": 8, "doc_count_error_upper_bound": 2}
    got = _aggs(agg_splits, {"names": {"terms": {"field": "name", "size": 1, "split_size": 5}}})["names"]
    assert got == {"buckets": [{"doc_count": 3, "key": "Fritz"}], "sum_other_doc_count": 7, "doc_count_error_upper_bound": 0}
    got = _aggs(agg_splits, {"names": {"terms": {"field": "high_prec_test"}}})["names"]
    assert got["buckets"] == [{"doc_count": 1, "key": 1769070189829214200}]


# ---- tests.rs:1315-1387 test_single_node_aggregation (terms ordered by a sub-aggregation) ------------------------------------
def test_terms_order_by_sub_agg():
    mapping = {"field_mappings": [{"name": "color", "type": "text", "tokenizer": "raw", "fast": True}, {"name": "price", "type": "f64", "fast": True}]}
    docs = [{"color": "blue", "price": 10.0}, {"color": "blue", "price": 15.0}, {"color": "green", "price": 10.0},
            {"color": "green", "price": 5.0}, {"color": "green", "price": 20.0}, {"color": "white", "price": 100.0},
            {"color": "white", "price": 1.0}]
    img = S.build_split(docs, mapping, "colors")
    aggs = {"expensive_colors": {"terms": {"field": "color", "order": {"price_stats.max": "desc"}},
                                 "aggs": {"price_stats": {"stats": {"field": "price"}}}}}
    got = cpu_root_search([img], MATCH_ALL, mapping, max_hits=0, aggs=aggs)["aggregations"]["expensive_colors"]["buckets"]
    assert [b["key"] for b in got] == ["white", "green", "blue"]
    assert got[0]["price_stats"] == {"avg": 50.5, "count": 2, "max": 100.0, "min": 1.0, "sum": 101.0}


# ---- rest-api-tests multi_splits/0001-request-optimizations.yaml: split-partition invariance -------------------------------
def test_split_partition_invariance():
    mapping = {"field_mappings": [{"name": "ts", "type": "datetime", "fast": True}, {"name": "body", "type": "text"}], "timestamp_field": "ts"}
    docs = [{"ts": 1_684_993_000 + i, "body": "hello"} for i in range(13)]
    rng = np.random.default_rng(7)
    for _ in range(6):
        perm = rng.permutation(13)
        cuts = sorted(rng.choice(np.arange(1, 13), size=int(rng.integers(0, 6)), replace=False).tolist())
        groups = [g for g in np.split(perm, cuts) if len(g)]
        imgs = [S.build_split([docs[i] for i in g], mapping, f"p{j}") for j, g in enumerate(groups)]
        for order in (ASC, DESC):
            for size in (1, 2, 3, 5):
                for window in (None, (1_684_993_002, 1_684_993_008)):
                    kw = dict(max_hits=size, sort_fields=[("ts", order)])
                    if window:
                        kw.update(start_timestamp=window[0], end_timestamp=window[1])
                    res = cpu_root_search(imgs, MATCH_ALL, mapping, **kw)
                    lo, hi = window or (1_684_993_000, 1_684_993_013)
                    vals = sorted(range(lo, hi), reverse=(order == DESC))[:size]
                    assert [h["sort_value"][1] // 10**9 for h in res["partial_hits"]] == vals
                    assert res["num_hits"] == hi - lo


def test_fieldnorm_table_matches_lucene_smallfloat():
    L = ffi.lib()
    table = [L.qwgpu_id_to_fieldnorm(i) for i in range(256)]
    assert table[:41] == list(range(41)) and table[41:49] == [42, 44, 46, 48, 50, 52, 54, 56]
    assert table[49:52] == [60, 64, 68] and table[255] == 2_013_265_944
    assert all(a < b for a, b in zip(table, table[1:]))
    for n in (0, 1, 40, 41, 57, 1000, 2**31):
        i = L.qwgpu_fieldnorm_to_id(n)
        assert table[i] <= n and (i == 255 or table[i + 1] > n)


# ---- committed fixture: synthetic-corpus BM25 top-20 (tests/golden/bm25_synth_expected.json) -----------------
def _synth_golden():
    import json as _json
    import os as _os
    with open(_os.path.join(_os.path.dirname(__file__), "golden", "bm25_synth_expected.json")) as f:
        want = _json.load(f)
    from quickwit_b200 import ffi as _ffi, plan as _P
    img = S.synth_split(20000, 3, [0.2, 0.1, 0.05, 0.01], split_id="golden-3")
    root = _P.bool_([_P.term(img, "body", f"t{i}", occur=_ffi.OCCUR_SHOULD) for i in range(4)])
    return want, img, _P.make_plan(root, 20, [(_ffi.SORT_SCORE, _ffi.ORDER_DESC, _ffi.ABSENT)])


def test_synthetic_corpus_fixture():
    """The seeded corpus generator + BM25 arithmetic reproduce the committed scores bit for bit."""
    from oracle import oracle as _O
    want, img, pl = _synth_golden()
    r = _O.split_search(img, pl)
    assert r.num_hits == want["num_hits"]
    assert [[int(h[0]), float(np.float32(h[4]))] for h in r.hits] == want["hits"]


def test_cpu_baseline_fast_path_equals_the_oracle():
    """The windowed-union / SIMD-unpack organisation that bench.py times as the CPU baseline returns exactly
    what the doc-at-a-time oracle returns (doc ids, f32 score bits, hit counts, postings visited), also from
    the C thread pool."""
    from quickwit_b200 import plan as P
    img = S.synth_split(60_000, 3, [0.2, 0.1, 0.05, 0.02, 0.01, 0.004, 0.0008], split_id="fast-0")
    plans = []
    for terms, k in [(range(7), 100), ([0], 10), ([6, 5, 4], 1000), ([1, 3, 5], 37)]:
        root = P.bool_([P.term(img, "body", f"t{i}", occur=ffi.OCCUR_SHOULD) for i in terms] + [P.term(img, "body", "absent", occur=ffi.OCCUR_SHOULD)])
        plans.append(P.make_plan(root, k, [(ffi.SORT_SCORE, ffi.ORDER_DESC, ffi.ABSENT)]))
    for pl in plans:
        a, b = O.split_search(img, pl), O.split_search(img, pl, fast=True)
        assert (a.num_hits, a.postings_visited) == (b.num_hits, b.postings_visited)
        assert [(h[0], h[1], h[2]) for h in a.hits] == [(h[0], h[1], h[2]) for h in b.hits]
        assert np.array_equal(np.array([h[4] for h in a.hits], np.float32).view(np.uint32), np.array([h[4] for h in b.hits], np.float32).view(np.uint32))
    many = O.ManySearch([img] * len(plans), plans)
    want = (sum(O.split_search(img, pl).num_hits for pl in plans), sum(O.split_search(img, pl).postings_visited for pl in plans))
    assert many.run(threads=3, fast=True) == want and many.run(threads=2, fast=False) == want
    # a shape the fast path does not cover falls back to the oracle proper
    pl = P.make_plan(P.bool_([P.term(img, "body", "t0"), P.term(img, "body", "t1")]), 10, [(ffi.SORT_SCORE, ffi.ORDER_DESC, ffi.ABSENT)])
    assert O.split_search(img, pl, fast=True).hits == O.split_search(img, pl).hits


# ---- phrase queries (full_text mode `phrase`, slop 0): semantics of rest-api-tests es_compatibility/0013-phrase-query.yaml
# ("zone of explosion" does not match the phrase `zone explosion`), PhraseScorer / Bm25Weight::for_terms scoring ------------
PHRASE_MAPPING = {"field_mappings": [{"name": "body", "type": "text", "record": "position", "fieldnorms": True},
                                     {"name": "tags", "type": "text", "record": "position"},
                                     {"name": "nopos", "type": "text", "record": "freq"}]}
PHRASE_DOCS = [{"body": "the quick brown fox jumps over the lazy dog", "nopos": "quick brown"},
               {"body": "quick fox brown quick brown fox"}, {"body": "brown fox"}, {"body": "a quick brown"},
               {"body": "quick brown quick brown quick brown", "tags": ["alpha beta", "gamma alpha", "beta"]},
               {"body": "there is a zone of explosion and a sign decoration"}, {"tags": ["alpha", "beta gamma"]}]


def phrase_ast(field, text, slop=None):
    mode = {"type": "phrase"}
    if slop is not None:
        mode["slop"] = slop
    return {"type": "full_text", "field": field, "text": text, "params": {"mode": mode}, "lenient": False}


def test_phrase_queries():
    img = S.build_split(PHRASE_DOCS, PHRASE_MAPPING, "phrases")
    run = lambda ast, **kw: cpu_root_search([img], ast, PHRASE_MAPPING, max_hits=10, **kw)
    ids = lambda res: sorted(h["doc_id"] for h in res["partial_hits"])
    assert ids(run(phrase_ast("body", "quick brown"))) == [0, 1, 3, 4]
    assert ids(run(phrase_ast("body", "Quick, BROWN fox!"))) == [0, 1]
    assert ids(run(phrase_ast("body", "fox quick"))) == [] and ids(run(phrase_ast("body", "quick quick"))) == []
    assert ids(run(phrase_ast("body", "zone explosion"))) == [] and ids(run(phrase_ast("body", "zone of explosion"))) == [5]
    assert ids(run(phrase_ast("body", "sign decoration"))) == [5]
    assert ids(run(phrase_ast("body", "quick unknownword"))) == []          # a term missing from the split
    assert ids(run(phrase_ast("body", "brown"))) == [0, 1, 2, 3, 4]         # one token: a plain term query
    # values of a multi-valued field are one position apart: no phrase across two values
    assert ids(run(phrase_ast("tags", "alpha beta"))) == [4] and ids(run(phrase_ast("tags", "beta gamma"))) == [6]
    assert ids(run(phrase_ast("tags", "gamma alpha"))) == [4] and ids(run(phrase_ast("tags", "beta alpha"))) == []
    # inside a bool, with other clauses
    both = bool_(must=[phrase_ast("body", "quick brown")], must_not=[term("body", "lazy")])
    assert ids(run(both)) == [1, 3, 4]
    # scores: (sum of the terms' idf) * 2.2 * count / (count + K1 * (1 - B + B * len / avg_len)), f32
    res = run(phrase_ast("body", "quick brown"), sort_fields=[("_score", DESC)])
    n = len(PHRASE_DOCS)
    f32 = np.float32
    idf = lambda df: f32(np.log(f32(1) + (f32(n - df) + f32(0.5)) / (f32(df) + f32(0.5))))
    lens = [len(d.get("body", "").split()) for d in PHRASE_DOCS]
    avg = f32(sum(lens)) / f32(n)
    weight = (idf(4) + idf(5)) * f32(2.2)
    counts = {0: 1, 1: 1, 3: 1, 4: 3}
    want = {d: weight * (f32(c) / (f32(c) + f32(1.2) * (f32(0.25) + f32(0.75) * f32(lens[d]) / avg))) for d, c in counts.items()}
    got = {h["doc_id"]: h["sort_value"][1] for h in res["partial_hits"]}
    assert set(got) == set(want)
    for d in want:
        assert abs(got[d] - float(want[d])) <= 1e-6 * float(want[d]), (d, got[d], want[d])
    assert [h["doc_id"] for h in res["partial_hits"]] == [4, 3, 1, 0]
    # errors: no positions on the field; slop is not implemented
    with pytest.raises(ffi.QwGpuError) as e:
        run(phrase_ast("nopos", "quick brown"))
    assert e.value.code == ffi.EINVALID_QUERY and "does not have positions indexed" in e.value.msg
    with pytest.raises(ffi.QwGpuError) as e:
        run(phrase_ast("body", "zone explosion", slop=1))
    assert e.value.code == ffi.EUNSUPPORTED
