"""Composite-key and aggregation paths that the BASELINE-shaped corpus does not reach: keys wider than
one 64-bit word (two 60-bit sort columns), radix-select descents through massive ties, histogram
buckets located through the raw-space boundary table (f64 / negative i64 columns, hard bounds,
fractional intervals), optional and multi-valued columns in the generic aggregation path. Every case
is compared with the CPU pipeline (oracle + the same host code)."""
import random

import pytest

from quickwit_b200 import splitgen as S
from quickwit_b200.proto import ASC, DESC
from pipeline import MATCH_ALL, bool_, cpu_root_search, term
from test_gpu_leaf_search import gpu_root_search, same

pytestmark = pytest.mark.gpu

MAPPING = {"field_mappings": [{"name": "a", "type": "u64", "fast": True}, {"name": "b", "type": "i64", "fast": True},
                              {"name": "c", "type": "f64", "fast": True}, {"name": "t", "type": "u64", "fast": True},
                              {"name": "full_i", "type": "i64", "fast": True}, {"name": "full_f", "type": "f64", "fast": True},
                              {"name": "tags", "type": "u64", "fast": True},   # multi-valued (lists in the docs)
                              {"name": "body", "type": "text", "record": "freq", "fieldnorms": True}]}


def _docs(n, seed):
    rnd = random.Random(seed)
    docs = []
    for i in range(n):
        d = {"t": i % 3, "body": "x y" if i % 2 else "x", "full_i": rnd.randrange(-5000, 5000), "full_f": rnd.uniform(-3.0, 7.5)}
        if i % 7:
            d["a"] = rnd.getrandbits(60)
        if i % 5:
            d["b"] = rnd.getrandbits(61) - (1 << 60)
        if i % 3:
            d["c"] = rnd.uniform(-1e6, 1e6)
        if i % 4:
            d["tags"] = [rnd.randrange(6) for _ in range(1 + i % 3)]
        docs.append(d)
    return docs


@pytest.fixture(scope="module")
def splits(gpu_ctx):
    imgs = [S.build_split(_docs(50_000, 11 + k), MAPPING, f"wide-{k}") for k in range(2)]
    for im in imgs:
        gpu_ctx.register_split(im)
    yield imgs
    for im in imgs:
        gpu_ctx.unregister_split(im.split_id)


def test_keys_wider_than_one_word(gpu_ctx, splits):
    for sort in ([("a", DESC), ("b", ASC)], [("b", DESC), ("a", DESC)], [("c", ASC), ("a", DESC)], [("a", ASC), ("c", DESC)], [("b", ASC)]):
        for k in (1, 7, 300):
            kw = dict(max_hits=k, sort_fields=sort)
            got, _ = gpu_root_search(gpu_ctx, splits, term("body", "y"), MAPPING, **kw)
            same(got, cpu_root_search(splits, term("body", "y"), MAPPING, **kw))


def test_massive_ties_descend_radix_levels(gpu_ctx, splits):
    # 3 distinct values over 50 000 docs: the level-0 digit holds far more than the candidate capacity
    for sort in ([("t", DESC)], [("t", ASC), ("a", DESC)], [("t", DESC), ("t", ASC)]):
        for k in (10, 1000):
            kw = dict(max_hits=k, sort_fields=sort)
            got, _ = gpu_root_search(gpu_ctx, splits, MATCH_ALL, MAPPING, **kw)
            same(got, cpu_root_search(splits, MATCH_ALL, MAPPING, **kw))
    # BM25 with only two distinct scores over 50 000 matches (two doc lengths): the union kernels take the
    # exact path through the score bits down to the doc-id bits
    for sort in ([("_score", DESC)], [("_score", ASC)], [("_score", DESC), ("t", ASC)]):
        for k in (10, 1000):
            kw = dict(max_hits=k, sort_fields=sort)
            got, _ = gpu_root_search(gpu_ctx, splits, term("body", "x"), MAPPING, **kw)
            same(got, cpu_root_search(splits, term("body", "x"), MAPPING, **kw))
    page1, _ = gpu_root_search(gpu_ctx, splits, MATCH_ALL, MAPPING, max_hits=50, sort_fields=[("t", ASC)])
    kw = dict(max_hits=50, sort_fields=[("t", ASC)], search_after=page1["partial_hits"][-1])
    got, _ = gpu_root_search(gpu_ctx, splits, MATCH_ALL, MAPPING, **kw)
    same(got, cpu_root_search(splits, MATCH_ALL, MAPPING, **kw))


def test_histograms_through_raw_space_bounds(gpu_ctx, splits):
    cases = [
        {"h": {"histogram": {"field": "full_f", "interval": 0.25, "offset": 0.1}}},
        {"h": {"histogram": {"field": "full_f", "interval": 1.5, "hard_bounds": {"min": -1.0, "max": 4.0}}}},
        {"h": {"histogram": {"field": "full_i", "interval": 7}}, "tt": {"terms": {"field": "t"}}},
        {"h": {"histogram": {"field": "full_i", "interval": 1000, "offset": -250, "min_doc_count": 1}}},
        {"h": {"histogram": {"field": "full_i", "interval": 333, "hard_bounds": {"min": -1000, "max": 999}}},
         "g": {"histogram": {"field": "full_f", "interval": 3}}},
    ]
    for aggs in cases:
        for ast in (MATCH_ALL, term("body", "y")):
            got, _ = gpu_root_search(gpu_ctx, splits, ast, MAPPING, max_hits=0, aggs=aggs)
            same(got, cpu_root_search(splits, ast, MAPPING, max_hits=0, aggs=aggs))


def test_generic_aggregations_optional_and_multivalued(gpu_ctx, splits):
    cases = [
        {"tags": {"terms": {"field": "tags"}}},
        {"by_t": {"terms": {"field": "t"}, "aggs": {"s": {"stats": {"field": "b"}}, "h": {"histogram": {"field": "full_i", "interval": 2500}}}}},
        {"opt": {"histogram": {"field": "c", "interval": 250000.0}, "aggs": {"m": {"stats": {"field": "full_i"}}}}},
        {"r": {"range": {"field": "full_i", "ranges": [{"to": -100}, {"from": -100, "to": 100}, {"from": 100}]}, "aggs": {"tt": {"terms": {"field": "tags"}}}}},
        {"st": {"stats": {"field": "a"}}, "mx": {"max": {"field": "full_f"}}},
        # privatised stats cells of the fast path (always-present single-valued columns)
        {"by_t": {"terms": {"field": "t"}, "aggs": {"s": {"stats": {"field": "full_i"}}, "f": {"avg": {"field": "full_f"}}}}},
        {"h": {"histogram": {"field": "full_i", "interval": 500}, "aggs": {"f": {"stats": {"field": "full_f"}}}}, "all": {"stats": {"field": "full_i"}},
         "tt": {"terms": {"field": "t"}}},
        {"mn": {"min": {"field": "full_f"}}, "sm": {"sum": {"field": "full_i"}}},
    ]
    for aggs in cases:
        ast = bool_(must=[term("body", "x")], must_not=[term("body", "y")])
        got, _ = gpu_root_search(gpu_ctx, splits, ast, MAPPING, max_hits=5, sort_fields=[("full_i", DESC)], aggs=aggs)
        want = cpu_root_search(splits, ast, MAPPING, max_hits=5, sort_fields=[("full_i", DESC)], aggs=aggs)
        assert got["num_hits"] == want["num_hits"] and got["partial_hits"] == want["partial_hits"]
        _assert_aggs_close(got["aggregations"], want["aggregations"])


def _assert_aggs_close(a, b):
    """Equal up to f64 summation order (sum / avg — `value` for the single-metric forms — of f64 columns
    are accumulated in a different order on the device; counts, keys, min and max are exact)."""
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.keys() == b.keys()
        for k in a:
            if k in ("sum", "avg", "value") and isinstance(a[k], float) and isinstance(b[k], float):
                assert a[k] == pytest.approx(b[k], rel=1e-11, abs=1e-9)
            else:
                _assert_aggs_close(a[k], b[k])
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b)
        for x, y in zip(a, b):
            _assert_aggs_close(x, y)
    else:
        assert a == b
