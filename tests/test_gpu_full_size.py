"""BASELINE-size checks (3.125 M docs per split, the C2 configuration's split size): exact parity
with the oracle where the oracle finishes in seconds, plus size-independent properties —
sortedness, hit-count identities, idempotence and invariance under the partition of splits into
leaf requests (what `merge_leaf_responses` must guarantee, collector.rs:832-974)."""
import json

import numpy as np
import pytest

from quickwit_b200 import proto, service, splitgen as S
from quickwit_b200.proto import ASC, DESC
from pipeline import MATCH_ALL, bool_, cpu_root_search, leafify, search_request, term
from test_gpu_leaf_search import FRACS, SYNTH_MAPPING, gpu_root_search, same

pytestmark = pytest.mark.gpu

DOCS_PER_SPLIT = 3_125_000
N_SPLITS = 8
T0 = 1_700_000_000


@pytest.fixture(scope="module")
def big(gpu_ctx):
    imgs = [S.synth_split(DOCS_PER_SPLIT, i, FRACS * 2, split_id=f"big-{i:02d}", ts_start_secs=T0 + 86_400 * i) for i in range(N_SPLITS)]
    for im in imgs:
        gpu_ctx.register_split(im)
    yield imgs
    for im in imgs:
        gpu_ctx.unregister_split(im.split_id)


OR10 = bool_(should=[term("body", f"t{i}") for i in range(10)])
OR10_B = bool_(should=[term("body", f"t{i}") for i in range(10, 20)])


def _leaf(ctx, imgs, ast, **kw) -> bytes:
    offsets = [proto.enc_split_offsets(im.split_id, im.num_docs) for im in imgs]
    return ctx.leaf_search(proto.enc_leaf_search_request(search_request(ast, **leafify(kw)), offsets, json.dumps(SYNTH_MAPPING)))


def test_c2_exact_against_oracle_at_full_split_size(gpu_ctx, big):
    kw = dict(max_hits=1000, sort_fields=[("_score", DESC)])
    for ast in (OR10, OR10_B):
        got, leaf = gpu_root_search(gpu_ctx, big, ast, SYNTH_MAPPING, **kw)
        same(got, cpu_root_search(big, ast, SYNTH_MAPPING, **kw))
        assert leaf["num_successful_splits"] == N_SPLITS
        scores = [np.float32(h["sort_value"][1]) for h in got["partial_hits"]]
        assert len(scores) == 1000 and all(a >= b for a, b in zip(scores, scores[1:]))
        assert all(s > 0 for s in scores)


def test_c1_c3_c4_exact_against_oracle_at_full_split_size(gpu_ctx, big):
    cases = [
        (term("severity_text", "ERROR"), dict(max_hits=10)),                                                     # C1
        (term("body", "t1"), dict(max_hits=10, sort_fields=[("_score", DESC)])),
        (bool_(must=[term("body", "t2")]), dict(max_hits=1000, sort_fields=[("timestamp", DESC)],
                                                start_timestamp=T0 + 21_600, end_timestamp=T0 + 5 * 86_400 - 21_600)),   # C3
        (bool_(must=[term("body", "t0"), term("body", "t1")], must_not=[term("body", "t4")]),
         dict(max_hits=100, sort_fields=[("tenant_id", ASC), ("timestamp", DESC)])),
        (MATCH_ALL, dict(max_hits=0, aggs={"by_sev": {"terms": {"field": "severity_text"}},
                                           "over_time": {"date_histogram": {"field": "timestamp", "fixed_interval": "1h"}}})),  # C4
    ]
    for ast, kw in cases:
        got, _ = gpu_root_search(gpu_ctx, big, ast, SYNTH_MAPPING, **kw)
        same(got, cpu_root_search(big, ast, SYNTH_MAPPING, **kw))
    # C3 properties: every hit inside [start, end), timestamps non-increasing
    ast, kw = cases[2]
    got, _ = gpu_root_search(gpu_ctx, big, ast, SYNTH_MAPPING, **kw)
    ts = [h["sort_value"][1] for h in got["partial_hits"]]
    assert all(a >= b for a, b in zip(ts, ts[1:]))
    # C4 properties: every doc lands in exactly one bucket of each aggregation
    got, _ = gpu_root_search(gpu_ctx, big, *cases[4][:1], SYNTH_MAPPING, **cases[4][1])
    total = N_SPLITS * DOCS_PER_SPLIT
    assert got["num_hits"] == total
    a = got["aggregations"]
    assert sum(b["doc_count"] for b in a["over_time"]["buckets"]) == total
    assert sum(b["doc_count"] for b in a["by_sev"]["buckets"]) + a["by_sev"]["sum_other_doc_count"] == total
    assert len(a["over_time"]["buckets"]) == 24 * N_SPLITS + 1   # T0 is 800 s past an hour boundary


def test_hit_count_identities(gpu_ctx, big):
    """num_hits of a term query is the dictionary's doc_freq; |A ∪ B| = |A| + |B| − |A ∩ B|;
    |A \\ B| = |A| − |A ∩ B| (count-only requests, max_hits = 0)."""
    def count(ast):
        return proto.dec_leaf_search_response(_leaf(gpu_ctx, big, ast, max_hits=0))["num_hits"]
    df = {t: sum(im.doc_freq(im.term_ord("body", t)) for im in big) for t in ("t0", "t3", "t9")}
    for t, n in df.items():
        assert count(term("body", t)) == n
    a, b = term("body", "t0"), term("body", "t3")
    n_and = count(bool_(must=[a, b]))
    assert count(bool_(should=[a, b])) == df["t0"] + df["t3"] - n_and
    assert count(bool_(must=[a], must_not=[b])) == df["t0"] - n_and
    assert count(MATCH_ALL) == N_SPLITS * DOCS_PER_SPLIT
    n_or10 = count(OR10)
    assert max(df.values()) <= n_or10 <= N_SPLITS * DOCS_PER_SPLIT


def test_idempotent_and_partition_invariant(gpu_ctx, big):
    kw = dict(max_hits=1000, sort_fields=[("_score", DESC)])
    whole = _leaf(gpu_ctx, big, OR10, **kw)
    again = proto.dec_leaf_search_response(_leaf(gpu_ctx, big, OR10, **kw))   # same request, same hits (timings in resource_stats differ)
    first = proto.dec_leaf_search_response(whole)
    assert again["num_hits"] == first["num_hits"] and again["partial_hits"] == first["partial_hits"]
    req_pb = search_request(OR10, **leafify(kw))
    want = proto.dec_leaf_search_response(service.merge_leaf_responses(req_pb, [whole]))
    for cut in (1, 3, 4):
        parts = [_leaf(gpu_ctx, big[:cut], OR10, **kw), _leaf(gpu_ctx, big[cut:], OR10, **kw)]
        got = proto.dec_leaf_search_response(service.merge_leaf_responses(req_pb, parts))
        assert got["num_hits"] == want["num_hits"] and got["partial_hits"] == want["partial_hits"]
    # one request per split, merged: the shape the root produces when every split sits on its own node
    singles = [_leaf(gpu_ctx, [im], OR10, **kw) for im in big]
    got = proto.dec_leaf_search_response(service.merge_leaf_responses(req_pb, singles))
    assert got["num_hits"] == want["num_hits"] and got["partial_hits"] == want["partial_hits"]


def test_search_after_pages_tile_the_result(gpu_ctx, big):
    """Paging with search_after over the full-size splits reproduces the one-shot top-K."""
    kw = dict(max_hits=300, sort_fields=[("timestamp", DESC)])
    ast = term("body", "t5")
    one_shot, _ = gpu_root_search(gpu_ctx, big, ast, SYNTH_MAPPING, **kw)
    pages, after = [], None
    for _ in range(3):
        pkw = dict(max_hits=100, sort_fields=[("timestamp", DESC)])
        if after is not None:
            pkw["search_after"] = after
        page, _ = gpu_root_search(gpu_ctx, big, ast, SYNTH_MAPPING, **pkw)
        assert len(page["partial_hits"]) == 100
        pages += page["partial_hits"]
        after = page["partial_hits"][-1]
    assert pages == one_shot["partial_hits"]
