"""GPU twins of the reference-golden tests (tests/test_oracle_goldens.py).

Every golden the CPU suite checks through `cpu_root_search` (oracle leaves + host merge / finalise) is run
again here with the leaves executed by `qwgpu_leaf_search` on the device: same assertions, same expected
values (quickwit-search/src/tests.rs, collector.rs tests, rest-api-tests aggregations), so the driver's
`-m gpu` run proves the reference's own vectors on the product path and not only on the oracle.
The pure-host merge goldens (collector.rs:1332-1389, 1793-2059) are re-run as well: they exercise
`qwgpu_merge_leaf_responses` from the library the GPU box loads."""
import pytest

import test_oracle_goldens as G
from test_gpu_leaf_search import gpu_root_search

pytestmark = pytest.mark.gpu


@pytest.fixture()
def on_gpu(gpu_ctx, monkeypatch):
    """Routes the golden module's `cpu_root_search` through the GPU leaf. Splits are registered for the
    duration of one call: several goldens reuse a split id with different contents."""
    calls = {"n": 0}

    def root_search(imgs, query_ast, doc_mapper, **req_kw):
        calls["n"] += 1
        for im in imgs:
            gpu_ctx.register_split(im)
        try:
            return gpu_root_search(gpu_ctx, imgs, query_ast, doc_mapper, **req_kw)[0]
        finally:
            for im in imgs:
                gpu_ctx.unregister_split(im.split_id)

    monkeypatch.setattr(G, "cpu_root_search", root_search)
    yield calls
    assert calls["n"] > 0, "the golden did not go through the GPU leaf"


@pytest.fixture(scope="module")
def agg_splits():
    return [G.S.build_split(G.AGG_SPLIT1, G.AGG_MAPPING, "agg-1"), G.S.build_split(G.AGG_SPLIT2, G.AGG_MAPPING, "agg-2")]


def test_sort_bm25_exact_f32_scores(on_gpu):
    G.test_sort_bm25_exact_f32_scores()


@pytest.mark.parametrize("spec", [[], [("sort1", G.DESC)], [("sort1", G.ASC)], [("sort1", G.DESC), ("sort2", G.DESC)],
                                  [("sort1", G.ASC), ("sort2", G.DESC)], [("sort1", G.DESC), ("sort2", G.ASC)],
                                  [("sort1", G.ASC), ("sort2", G.ASC)]])
def test_single_split_sorting(on_gpu, spec):
    G.test_single_split_sorting(spec)


def test_search_after(on_gpu):
    G.test_search_after()


def test_several_splits_default_order(on_gpu):
    G.test_several_splits_default_order()


def test_filtering_term_and_timestamp_range(on_gpu):
    G.test_filtering_term_and_timestamp_range()


def test_range_queries(on_gpu):
    G.test_range_queries()


def test_agg_date_histogram(on_gpu, agg_splits):
    G.test_agg_date_histogram(agg_splits)


def test_stats_over_a_datetime_column_does_not_overflow(on_gpu):
    G.test_stats_over_a_datetime_column_does_not_overflow()


def test_agg_range_and_histogram(on_gpu, agg_splits):
    G.test_agg_range_and_histogram(agg_splits)


def test_agg_terms(on_gpu, agg_splits):
    G.test_agg_terms(agg_splits)


def test_terms_order_by_sub_agg(on_gpu):
    G.test_terms_order_by_sub_agg()


def test_split_partition_invariance(on_gpu):
    G.test_split_partition_invariance()


# host-side merge goldens, on the GPU box's build of the library
def test_merge_goldens():
    G.test_merge_partial_hits_no_tie()
    G.test_merge_partial_hits_with_tie()
    G.test_merge_collectors()
    G.test_merge_empty_intermediate_aggregation_result()


def test_phrase_queries(on_gpu):
    G.test_phrase_queries()
