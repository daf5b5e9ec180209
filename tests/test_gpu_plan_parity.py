"""GPU (libqwgpu.so, seam C) vs CPU oracle on the same seeded split images and plans.

Bit-exact bar: doc ids, sort values, hit counts, bucket counts, f32 BM25 score bits."""
import numpy as np
import pytest

from quickwit_b200 import ffi, plan as P, splitgen as S
from oracle import oracle as O
from helpers import (DOC_ASC, DOC_DESC, SCORE_DESC, assert_same, col_sort, histogram_agg, range_agg, stats_agg,
                     terms_agg)

pytestmark = pytest.mark.gpu

FRACS = [0.2, 0.1, 0.05, 0.05, 0.02, 0.02, 0.01, 0.01, 0.005, 0.001]


@pytest.fixture(scope="module")
def synth(gpu_ctx):
    imgs = [S.synth_split(50_000 + 777 * i, i, FRACS, split_id=f"synth-{i}", ts_start_secs=1_700_000_000 + 86_400 * i)
            for i in range(3)]
    for im in imgs:
        gpu_ctx.register_split(im)
    return imgs


def run_both(sctx, img, plan, **kw):
    got = sctx.split_search([img.split_id], [plan])[0]
    want = O.split_search(img, plan)
    assert_same(got, want, **kw)
    return got, want


def T(img, i, occur=ffi.OCCUR_SHOULD, boost=1.0):
    return P.term(img, "body", f"t{i}", occur=occur, boost=boost)


def test_single_term_scored(gpu_ctx, synth):
    img = synth[0]
    for i in (0, 4, 9):
        for k in (1, 10, 100):
            got, _ = run_both(gpu_ctx, img, P.make_plan(T(img, i, ffi.OCCUR_MUST), k, SCORE_DESC), ctx=f"t{i} k{k}")
            assert got.num_hits == img.doc_freq(img.term_ord("body", f"t{i}"))


def test_bm25_or_topk(gpu_ctx, synth):
    img = synth[1]
    for nterms, k in [(2, 10), (4, 100), (10, 1000), (10, 4096)]:
        root = P.bool_([T(img, i) for i in range(nterms)])
        got, _ = run_both(gpu_ctx, img, P.make_plan(root, k, SCORE_DESC), ctx=f"or{nterms} k{k}")
        assert len(got.hits) == min(k, got.num_hits)
        assert got.num_kernel_launches > 0


def test_bm25_score_asc_and_boost(gpu_ctx, synth):
    img = synth[0]
    root = P.bool_([T(img, 1, boost=2.5), T(img, 3), T(img, 7, boost=0.5)])
    run_both(gpu_ctx, img, P.make_plan(root, 50, [(ffi.SORT_SCORE, ffi.ORDER_ASC, ffi.ABSENT)]), ctx="score asc")


def test_conjunction_and_filters(gpu_ctx, synth):
    img = synth[2]
    ts = img.columns()[img.column_ord("timestamp")]
    lo = ts.min_value + (ts.max_value - ts.min_value) // 4
    hi = ts.min_value + 3 * (ts.max_value - ts.min_value) // 4
    plans = {
        "and2 scored": P.make_plan(P.bool_([T(img, 0, ffi.OCCUR_MUST), T(img, 1, ffi.OCCUR_MUST)]), 100, SCORE_DESC),
        "and3 docdesc": P.make_plan(P.bool_([T(img, 0, ffi.OCCUR_MUST), T(img, 2, ffi.OCCUR_MUST), T(img, 1, ffi.OCCUR_MUST)]), 20, DOC_DESC),
        "term+range ts desc": P.make_plan(P.bool_([T(img, 2, ffi.OCCUR_MUST), P.range_(img, "timestamp", lo, hi)]), 1000,
                                          [col_sort(img, "timestamp", ffi.ORDER_DESC)]),
        "term+range ts asc": P.make_plan(P.bool_([T(img, 2, ffi.OCCUR_MUST), P.range_(img, "timestamp", lo, hi)]), 1000,
                                         [col_sort(img, "timestamp", ffi.ORDER_ASC)]),
        "range only": P.make_plan(P.range_(img, "timestamp", lo, hi, occur=ffi.OCCUR_MUST), 10, DOC_ASC),
        "must + should scored": P.make_plan(P.bool_([T(img, 1, ffi.OCCUR_MUST), T(img, 0), T(img, 5)]), 200, SCORE_DESC),
        "must_not": P.make_plan(P.bool_([T(img, 0, ffi.OCCUR_MUST), T(img, 1, ffi.OCCUR_MUST_NOT)]), 100, DOC_DESC),
        "only must_not": P.make_plan(P.bool_([T(img, 0, ffi.OCCUR_MUST_NOT)]), 100, DOC_DESC),
        "msm2": P.make_plan(P.bool_([T(img, i) for i in range(5)], min_should_match=2), 300, SCORE_DESC),
        "msm3 + must": P.make_plan(P.bool_([T(img, 0, ffi.OCCUR_MUST)] + [T(img, i) for i in range(1, 6)], min_should_match=3), 300, SCORE_DESC),
        "nested": P.make_plan(P.bool_([P.bool_([T(img, 0), T(img, 1)], occur=ffi.OCCUR_MUST),
                                       P.bool_([T(img, 2, ffi.OCCUR_MUST), T(img, 3, ffi.OCCUR_MUST)], occur=ffi.OCCUR_SHOULD),
                                       T(img, 4, ffi.OCCUR_MUST_NOT)]), 150, SCORE_DESC),
        "filter bool": P.make_plan(P.bool_([T(img, 0, ffi.OCCUR_MUST), P.bool_([T(img, 1), T(img, 2)], occur=ffi.OCCUR_FILTER)]), 77, SCORE_DESC),
        "absent term should": P.make_plan(P.bool_([T(img, 0), P.term(img, "body", "nope", occur=ffi.OCCUR_SHOULD)]), 10, SCORE_DESC),
        "absent term must": P.make_plan(P.bool_([T(img, 0, ffi.OCCUR_MUST), P.term(img, "body", "nope", occur=ffi.OCCUR_MUST)]), 10, SCORE_DESC),
        "match_all doc desc": P.make_plan(P.match_all(), 10, DOC_DESC),
        "severity term default sort": P.make_plan(P.term(img, "severity_text", "ERROR"), 10, DOC_DESC),
        "count only": P.make_plan(T(img, 3, ffi.OCCUR_MUST), 0, DOC_DESC),
    }
    for name, pl in plans.items():
        run_both(gpu_ctx, img, pl, ctx=name)


def test_sort_by_columns(gpu_ctx, synth):
    img = synth[0]
    root = P.bool_([T(img, 0), T(img, 3)])
    for o1 in (ffi.ORDER_ASC, ffi.ORDER_DESC):
        for o2 in (ffi.ORDER_ASC, ffi.ORDER_DESC):
            pl = P.make_plan(root, 500, [col_sort(img, "tenant_id", o1), col_sort(img, "timestamp", o2)])
            run_both(gpu_ctx, img, pl, ctx=f"tenant {o1} ts {o2}")
    pl = P.make_plan(root, 500, [col_sort(img, "tenant_id", ffi.ORDER_DESC), (ffi.SORT_SCORE, ffi.ORDER_DESC, ffi.ABSENT)])
    run_both(gpu_ctx, img, pl, ctx="tenant then score")
    pl = P.make_plan(root, 100, [col_sort(img, "no_such_column", ffi.ORDER_DESC)])
    run_both(gpu_ctx, img, pl, ctx="missing sort column")


def test_aggregations(gpu_ctx, synth):
    img = synth[1]
    hour = 3600e9
    aggs = [terms_agg(img, "severity_text"), histogram_agg(img, "timestamp", hour)]
    got, want = run_both(gpu_ctx, img, P.make_plan(P.match_all(), 0, DOC_DESC, aggs=aggs), ctx="C4 aggs")
    assert sum(c[0] for c in got.cells[:4]) == img.num_docs
    nested = [terms_agg(img, "severity_text", children=[histogram_agg(img, "timestamp", hour, children=[stats_agg(img, "tenant_id")])]),
              stats_agg(img, "timestamp"), terms_agg(img, "tenant_id", children=[stats_agg(img, "timestamp")]),
              range_agg(img, "tenant_id", [(0, 1005), (1005, 1050), (1050, 2**64 - 1)])]
    root = P.bool_([T(img, 0), T(img, 2)])
    run_both(gpu_ctx, img, P.make_plan(root, 10, SCORE_DESC, aggs=nested), ctx="nested aggs")


def test_search_after(gpu_ctx, synth):
    img = synth[0]
    root = P.bool_([T(img, 0), T(img, 1)])
    base = P.make_plan(root, 40, [col_sort(img, "tenant_id", ffi.ORDER_DESC), col_sort(img, "timestamp", ffi.ORDER_ASC)])
    want = O.split_search(img, base)
    marker = want.hits[19]
    for coe, pre in [(0, 0), (1, 0), (1, -1), (1, 1)]:
        sa = ffi.QwSearchAfter(1, 1, 1, coe, pre, marker[0], marker[2], marker[3])
        pl = P.make_plan(root, 40, [col_sort(img, "tenant_id", ffi.ORDER_DESC), col_sort(img, "timestamp", ffi.ORDER_ASC)], search_after=sa)
        got, w2 = run_both(gpu_ctx, img, pl, ctx=f"search_after coe={coe} pre={pre}")
        if coe == 1 and pre == 0:
            assert [h[0] for h in got.hits[:20]] == [h[0] for h in want.hits[20:40]]
    sa = ffi.QwSearchAfter(1, 1, 0, 0, 0, 0, S.f64_to_u64(float(np.float32(3.0))), 0)
    run_both(gpu_ctx, img, P.make_plan(root, 25, SCORE_DESC, search_after=sa), ctx="search_after score")


def test_sort_matrix_with_nones(gpu_ctx):
    """The reference's 17-doc None/tie matrix (quickwit-search/src/collector.rs:1391-1413)."""
    data = [(2, 1), (0, 1), (1, 1), (0, 0), (None, 1), (None, 2), (2, 1), (1, 2), (0, None), (None, 0), (2, 0), (2, 2),
            (0, 2), (2, None), (None, None), (1, 0), (1, None)]
    docs = [{k: v for k, v in (("sort1", a), ("sort2", b)) if v is not None} for a, b in data]
    mapping = {"field_mappings": [{"name": "sort1", "type": "u64", "fast": True}, {"name": "sort2", "type": "u64", "fast": True}]}
    img = S.build_split(docs, mapping, split_id="sortmatrix")
    gpu_ctx.register_split(img)
    specs = [[], ["sort1"], ["-sort1"], ["sort1", "sort2"], ["-sort1", "sort2"], ["sort1", "-sort2"], ["-sort1", "-sort2"]]
    for spec in specs:
        sort = [col_sort(img, f.lstrip("-"), ffi.ORDER_ASC if f.startswith("-") else ffi.ORDER_DESC) for f in spec] or DOC_DESC
        for k in range(0, 18):
            run_both(gpu_ctx, img, P.make_plan(P.match_all(), k, sort), ctx=f"{spec} k={k}")


def test_batched_splits(gpu_ctx, synth):
    plans, ids = [], []
    for img in synth:
        plans.append(P.make_plan(P.bool_([T(img, i) for i in range(10)]), 1000, SCORE_DESC))
        ids.append(img.split_id)
    got = gpu_ctx.split_search(ids, plans)
    for img, pl, g in zip(synth, plans, got):
        assert_same(g, O.split_search(img, pl), ctx=img.split_id)


def test_degenerate_ties_force_exact_radix_select(gpu_ctx):
    """All docs share the sort value: the threshold digit cannot separate them, so the engine has
    to walk the radix levels down to the doc id."""
    n = 30_000
    docs_v = np.zeros(n, dtype=np.uint64) + 7
    b = S._Builder(n)
    b.add_column("v", ffi.COL_U64, ffi.CARD_FULL, docs_v, None)
    img = b.finish("ties")
    gpu_ctx.register_split(img)
    for order in (ffi.ORDER_ASC, ffi.ORDER_DESC):
        run_both(gpu_ctx, img, P.make_plan(P.match_all(), 100, [col_sort(img, "v", order)]), ctx=f"ties {order}")


def test_synthetic_corpus_fixture_on_gpu(gpu_ctx):
    """tests/golden/bm25_synth_expected.json through the CUDA path (bit-exact f32 scores)."""
    import numpy as np
    from test_oracle_goldens import _synth_golden
    want, img, pl = _synth_golden()
    gpu_ctx.register_split(img)
    r = gpu_ctx.split_search([img.split_id], [pl])[0]
    assert r.num_hits == want["num_hits"]
    assert [[int(h[0]), float(np.float32(h[4]))] for h in r.hits] == want["hits"]
