import json, os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from quickwit_b200 import proto, service
from quickwit_b200.service import SearcherContext
imgs = bench.build_splits(0, 32, 3_125_000, threads=32)
ctx = SearcherContext(0)
for im in imgs: ctx.register_split(im)
dm = json.dumps({"field_mappings": [{"name": "body", "type": "text", "record": "freq", "fieldnorms": True}, {"name": "timestamp", "type": "datetime", "fast": True}], "timestamp_field": "timestamp"})
offsets = [proto.enc_split_offsets(im.split_id, im.num_docs) for im in imgs]
sr = proto.enc_search_request(json.dumps({"type": "bool", "should": [{"type": "term", "field": "body", "value": f"t{i}"} for i in range(10)]}), max_hits=1000, sort_fields=[("_score", 1)])
lr = proto.enc_leaf_search_request(sr, offsets, dm)
for i in range(12):
    t=time.perf_counter(); ctx.leaf_search(lr); print("wall_us", 1e6*(time.perf_counter()-t), file=sys.stderr)
