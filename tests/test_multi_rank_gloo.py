"""N > 1 path on CPU: two ranks (gloo), each owning its own splits, exchange fixed-size partials
with ONE all-gather and every rank runs the root merge — must equal a single-process merge of
all the per-rank leaf responses (merge_leaf_responses, collector.rs:914-974)."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quickwit_b200 import proto, service, splitgen as S
from quickwit_b200.proto import DESC
from pipeline import MATCH_ALL, bool_, cpu_split_response, search_request, term

FRACS = [0.2, 0.1, 0.05, 0.02]
MAPPING = {"field_mappings": [{"name": "body", "type": "text", "record": "freq", "fieldnorms": True},
                              {"name": "severity_text", "type": "text", "tokenizer": "raw", "fast": True},
                              {"name": "timestamp", "type": "datetime", "fast": True}, {"name": "tenant_id", "type": "u64", "fast": True}],
           "timestamp_field": "timestamp"}
CASES = [
    (bool_(should=[term("body", f"t{i}") for i in range(4)]), dict(max_hits=50, sort_fields=[("_score", DESC)])),
    (MATCH_ALL, dict(max_hits=5, sort_fields=[("timestamp", DESC)],
                     aggs={"by_sev": {"terms": {"field": "severity_text"}},
                           "over_time": {"date_histogram": {"field": "timestamp", "fixed_interval": "6h"}}})),
]


def rank_leaf_response(rank, req_pb):
    imgs = [S.synth_split(6000 + 500 * s, rank * 2 + s, FRACS, split_id=f"r{rank}-s{s}", ts_start_secs=1_700_000_000 + 86_400 * (rank * 2 + s))
            for s in range(2)]
    return service.merge_leaf_responses(req_pb, [cpu_split_response(im, req_pb, MAPPING) for im in imgs])


def worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for ci, (ast, kw) in enumerate(CASES):
            req_pb = search_request(ast, **kw)
            mine = rank_leaf_response(rank, req_pb)
            nbytes = service.partial_size(req_pb)
            part = torch.zeros(nbytes, dtype=torch.uint8)
            service.response_to_partial(req_pb, mine, part.data_ptr(), nbytes)
            gathered = torch.zeros(world * nbytes, dtype=torch.uint8)
            dist.all_gather_into_tensor(gathered, part)   # the single collective of the data path
            merged = service.merge_partials(req_pb, world, gathered.data_ptr(), nbytes)
            with open(os.path.join(out_dir, f"case{ci}-rank{rank}.bin"), "wb") as f:
                f.write(merged)
    finally:
        dist.destroy_process_group()


def test_two_ranks_allgather_merge(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for ci, (ast, kw) in enumerate(CASES):
        req_pb = search_request(ast, **kw)
        want = proto.dec_leaf_search_response(service.merge_leaf_responses(req_pb, [rank_leaf_response(r, req_pb) for r in range(2)]))
        for rank in range(2):
            got = proto.dec_leaf_search_response(open(tmp_path / f"case{ci}-rank{rank}.bin", "rb").read())
            assert got["num_hits"] == want["num_hits"] and got["partial_hits"] == want["partial_hits"]
            assert got["num_successful_splits"] == 4
            if kw.get("aggs"):
                fin = lambda r: json.loads(service.finalize_aggregation(json.dumps(kw["aggs"]), r["intermediate_aggregation_result"]))
                assert fin(got) == fin(want)
