// proto.cpp — encode/decode of the quickwit.search messages (see proto.h for field numbers).
#include "proto.h"

namespace qw {
namespace pb {

static SortValue decode_sort_by_value(Reader r) {
  SortValue v;
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (f == 1 && wt == 0) { v.kind = SortValue::U64; v.u = r.varint(); }
    else if (f == 2 && wt == 0) { v.kind = SortValue::I64; v.i = (int64_t)r.varint(); }
    else if (f == 3 && wt == 1) { v.kind = SortValue::F64; uint64_t b = r.fixed64(); memcpy(&v.f, &b, 8); }
    else if (f == 4 && wt == 0) { v.kind = SortValue::Bool; v.b = r.varint() != 0; }
    else r.skip(wt);
  }
  return v;
}
static std::string encode_sort_by_value(const SortValue& v) {
  Writer w;
  switch (v.kind) {
    case SortValue::U64: w.u64_always(1, v.u); break;
    case SortValue::I64: w.u64_always(2, (uint64_t)v.i); break;
    case SortValue::F64: w.f64_always(3, v.f); break;
    case SortValue::Bool: w.u64_always(4, v.b ? 1 : 0); break;
    default: break;
  }
  return w.out;
}

PartialHit decode_partial_hit(Reader r) {
  PartialHit h;
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (f == 10 && wt == 2) { h.has_sv1 = true; h.sv1 = decode_sort_by_value(r.sub()); }
    else if (f == 11 && wt == 2) { h.has_sv2 = true; h.sv2 = decode_sort_by_value(r.sub()); }
    else if (f == 2 && wt == 2) h.split_id = r.str();
    else if (f == 3 && wt == 0) h.segment_ord = (uint32_t)r.varint();
    else if (f == 4 && wt == 0) h.doc_id = (uint32_t)r.varint();
    else r.skip(wt);
  }
  return h;
}
std::string encode_partial_hit(const PartialHit& h) {
  Writer w;
  w.str(2, h.split_id);
  w.u64(3, h.segment_ord);
  w.u64(4, h.doc_id);
  if (h.has_sv1) w.bytes(10, encode_sort_by_value(h.sv1));
  if (h.has_sv2) w.bytes(11, encode_sort_by_value(h.sv2));
  return w.out;
}

// Same bytes as `Writer::bytes(field, encode_partial_hit(..))`, written straight into `out` through a
// per-thread scratch buffer: a leaf response carries up to thousands of hits and the temporaries of
// the generic path dominated its encoding time.
void append_partial_hit(std::string& out, uint32_t field, const std::string& split_id, uint32_t segment_ord, uint32_t doc_id,
                        bool has_sv1, const SortValue& sv1, bool has_sv2, const SortValue& sv2) {
  // body: split_id (2) | segment_ord (3) | doc_id (4) | sort_value (10) | sort_value2 (11); a body is at
  // most 2 + len + 6 + 6 + 2 * 13 bytes, assembled in a stack buffer and appended once
  char stack[192];
  std::string big;
  char* buf = stack;
  if (split_id.size() > sizeof(stack) - 48) { big.resize(split_id.size() + 48); buf = &big[0]; }
  char* p = buf;
  auto varint = [&](uint64_t v) { while (v >= 0x80) { *p++ = (char)(v | 0x80); v >>= 7; } *p++ = (char)v; };
  auto tag = [&](uint32_t f, uint32_t wt) { varint(((uint64_t)f << 3) | wt); };
  if (!split_id.empty()) { tag(2, 2); varint(split_id.size()); memcpy(p, split_id.data(), split_id.size()); p += split_id.size(); }
  if (segment_ord) { tag(3, 0); varint(segment_ord); }
  if (doc_id) { tag(4, 0); varint(doc_id); }
  auto sort_value = [&](uint32_t f, const SortValue& v) {
    tag(f, 2);
    char* len = p++;  // the nested SortByValue is at most 11 bytes: one length byte
    switch (v.kind) {
      case SortValue::U64: tag(1, 0); varint(v.u); break;
      case SortValue::I64: tag(2, 0); varint((uint64_t)v.i); break;
      case SortValue::F64: { tag(3, 1); uint64_t b; memcpy(&b, &v.f, 8); memcpy(p, &b, 8); p += 8; break; }
      case SortValue::Bool: tag(4, 0); varint(v.b ? 1 : 0); break;
      default: break;
    }
    *len = (char)(p - len - 1);
  };
  if (has_sv1) sort_value(10, sv1);
  if (has_sv2) sort_value(11, sv2);
  const size_t n = (size_t)(p - buf);
  char head[16];
  char* q = head;
  { uint64_t v = ((uint64_t)field << 3) | 2; while (v >= 0x80) { *q++ = (char)(v | 0x80); v >>= 7; } *q++ = (char)v; }
  { uint64_t v = n; while (v >= 0x80) { *q++ = (char)(v | 0x80); v >>= 7; } *q++ = (char)v; }
  out.append(head, (size_t)(q - head));
  out.append(buf, n);
}

static SortField decode_sort_field(Reader r) {
  SortField s;
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (f == 1 && wt == 2) s.field_name = r.str();
    else if (f == 2 && wt == 0) s.sort_order = (int32_t)r.varint();
    else r.skip(wt);
  }
  return s;
}

SearchRequest decode_search_request(Reader r) {
  SearchRequest q;
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (f == 1 && wt == 2) q.index_id_patterns.push_back(r.str());
    else if (f == 13 && wt == 2) q.query_ast = r.str();
    else if (f == 4 && wt == 0) q.start_timestamp = (int64_t)r.varint();
    else if (f == 5 && wt == 0) q.end_timestamp = (int64_t)r.varint();
    else if (f == 6 && wt == 0) q.max_hits = r.varint();
    else if (f == 7 && wt == 0) q.start_offset = r.varint();
    else if (f == 11 && wt == 2) q.aggregation_request = r.str();
    else if (f == 12 && wt == 2) q.snippet_fields.push_back(r.str());
    else if (f == 14 && wt == 2) q.sort_fields.push_back(decode_sort_field(r.sub()));
    else if (f == 15 && wt == 0) q.scroll_ttl_secs = (uint32_t)r.varint();
    else if (f == 16 && wt == 2) q.search_after = decode_partial_hit(r.sub());
    else if (f == 17 && wt == 0) q.count_hits = (int32_t)r.varint();
    else if (f == 18 && wt == 0) q.ignore_missing_indexes = r.varint() != 0;
    else if (f == 19 && wt == 0) q.skip_aggregation_finalization = r.varint() != 0;
    else r.skip(wt);
  }
  return q;
}

std::string encode_search_request(const SearchRequest& q) {
  Writer w;
  for (auto& s : q.index_id_patterns) w.bytes(1, s);
  if (q.start_timestamp) w.u64_always(4, (uint64_t)*q.start_timestamp);
  if (q.end_timestamp) w.u64_always(5, (uint64_t)*q.end_timestamp);
  w.u64(6, q.max_hits);
  w.u64(7, q.start_offset);
  if (q.aggregation_request) w.bytes(11, *q.aggregation_request);
  for (auto& s : q.snippet_fields) w.bytes(12, s);
  w.str(13, q.query_ast);
  for (auto& sf : q.sort_fields) {
    Writer s;
    s.str(1, sf.field_name);
    s.u64(2, (uint64_t)sf.sort_order);
    w.bytes(14, s.out);
  }
  if (q.scroll_ttl_secs) w.u64_always(15, *q.scroll_ttl_secs);
  if (q.search_after) w.bytes(16, encode_partial_hit(*q.search_after));
  w.u64(17, (uint64_t)q.count_hits);
  w.boolean(18, q.ignore_missing_indexes);
  w.boolean(19, q.skip_aggregation_finalization);
  return w.out;
}

static SplitIdAndFooterOffsets decode_split_offsets(Reader r) {
  SplitIdAndFooterOffsets s;
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (f == 1 && wt == 2) s.split_id = r.str();
    else if (f == 2 && wt == 0) s.split_footer_start = r.varint();
    else if (f == 3 && wt == 0) s.split_footer_end = r.varint();
    else if (f == 4 && wt == 0) s.timestamp_start = (int64_t)r.varint();
    else if (f == 5 && wt == 0) s.timestamp_end = (int64_t)r.varint();
    else if (f == 6 && wt == 0) s.num_docs = r.varint();
    else r.skip(wt);
  }
  return s;
}

LeafSearchRequest decode_leaf_search_request(const uint8_t* p, size_t n) {
  LeafSearchRequest q;
  Reader r(p, n);
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (f == 1 && wt == 2) q.search_request = decode_search_request(r.sub());
    else if (f == 7 && wt == 2) {
      Reader s = r.sub();
      LeafRequestRef lr;
      while (!s.done()) {
        uint64_t k2 = s.varint();
        uint32_t f2 = (uint32_t)(k2 >> 3), w2 = (uint32_t)(k2 & 7);
        if (f2 == 1 && w2 == 0) lr.doc_mapper_ord = (uint32_t)s.varint();
        else if (f2 == 2 && w2 == 0) lr.index_uri_ord = (uint32_t)s.varint();
        else if (f2 == 3 && w2 == 2) lr.split_offsets.push_back(decode_split_offsets(s.sub()));
        else s.skip(w2);
      }
      q.leaf_requests.push_back(std::move(lr));
    } else if (f == 8 && wt == 2) q.doc_mappers.push_back(r.str());
    else if (f == 9 && wt == 2) q.index_uris.push_back(r.str());
    else r.skip(wt);
  }
  return q;
}

static SplitResourceStats decode_split_stats(Reader r) {
  SplitResourceStats s;
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (wt == 0 && f >= 1 && f <= 9) s.v[f - 1] = r.varint();
    else r.skip(wt);
  }
  return s;
}
static std::string encode_split_stats(const SplitResourceStats& s) {
  Writer w;
  for (uint32_t f = 1; f <= 9; f++) w.u64(f, s.v[f - 1]);
  return w.out;
}
static LeafResourceStats decode_leaf_stats(Reader r) {
  LeafResourceStats s;
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (wt == 0 && f == 1) s.partial_result_cache_num_splits = r.varint();
    else if (wt == 0 && f == 2) s.partial_result_cache_num_docs = r.varint();
    else if (wt == 0 && f == 3) s.localexec_num_splits = r.varint();
    else if (wt == 0 && f == 4) s.localexec_num_docs = r.varint();
    else if (wt == 2 && f == 5) s.split_resources_worst = decode_split_stats(r.sub());
    else if (wt == 2 && f == 6) s.split_resources_sum = decode_split_stats(r.sub());
    else if (wt == 0 && f == 7) s.min_wait_for_search_permit_microsecs = r.varint();
    else if (wt == 0 && f == 8) s.min_wait_for_cpu_pool_microsecs = r.varint();
    else if (wt == 0 && f == 9) s.wall_time_microsecs = r.varint();
    else if (wt == 0 && f >= 10 && f <= 14) s.lambda[f - 10] = r.varint();
    else r.skip(wt);
  }
  return s;
}
static std::string encode_leaf_stats(const LeafResourceStats& s) {
  Writer w;
  w.u64(1, s.partial_result_cache_num_splits);
  w.u64(2, s.partial_result_cache_num_docs);
  w.u64(3, s.localexec_num_splits);
  w.u64(4, s.localexec_num_docs);
  if (s.split_resources_worst) w.bytes(5, encode_split_stats(*s.split_resources_worst));
  if (s.split_resources_sum) w.bytes(6, encode_split_stats(*s.split_resources_sum));
  if (s.min_wait_for_search_permit_microsecs) w.u64_always(7, *s.min_wait_for_search_permit_microsecs);
  if (s.min_wait_for_cpu_pool_microsecs) w.u64_always(8, *s.min_wait_for_cpu_pool_microsecs);
  w.u64(9, s.wall_time_microsecs);
  for (uint32_t f = 10; f <= 14; f++) w.u64(f, s.lambda[f - 10]);
  return w.out;
}

LeafSearchResponse decode_leaf_search_response(const uint8_t* p, size_t n) {
  LeafSearchResponse q;
  Reader r(p, n);
  while (!r.done()) {
    uint64_t key = r.varint();
    uint32_t f = (uint32_t)(key >> 3), wt = (uint32_t)(key & 7);
    if (f == 1 && wt == 0) q.num_hits = r.varint();
    else if (f == 2 && wt == 2) q.partial_hits.push_back(decode_partial_hit(r.sub()));
    else if (f == 3 && wt == 2) {
      Reader s = r.sub();
      SplitSearchError e;
      while (!s.done()) {
        uint64_t k2 = s.varint();
        uint32_t f2 = (uint32_t)(k2 >> 3), w2 = (uint32_t)(k2 & 7);
        if (f2 == 1 && w2 == 2) e.error = s.str();
        else if (f2 == 2 && w2 == 2) e.split_id = s.str();
        else if (f2 == 3 && w2 == 0) e.retryable_error = s.varint() != 0;
        else s.skip(w2);
      }
      q.failed_splits.push_back(std::move(e));
    } else if (f == 4 && wt == 0) q.num_attempted_splits = r.varint();
    else if (f == 6 && wt == 2) q.intermediate_aggregation_result = r.str();
    else if (f == 7 && wt == 0) q.num_successful_splits = r.varint();
    else if (f == 9 && wt == 2) q.resource_stats = decode_leaf_stats(r.sub());
    else r.skip(wt);
  }
  return q;
}

std::string encode_leaf_search_response(const LeafSearchResponse& q) {
  Writer w;
  w.out.reserve(64 + q.encoded_partial_hits.size() + 72 * q.partial_hits.size());
  w.u64(1, q.num_hits);
  for (auto& h : q.partial_hits) append_partial_hit(w.out, 2, h.split_id, h.segment_ord, h.doc_id, h.has_sv1, h.sv1, h.has_sv2, h.sv2);
  w.out += q.encoded_partial_hits;  // hits a merge wrote in wire form already (same field, same order)
  for (auto& e : q.failed_splits) {
    Writer s;
    s.str(1, e.error);
    s.str(2, e.split_id);
    s.boolean(3, e.retryable_error);
    w.bytes(3, s.out);
  }
  w.u64(4, q.num_attempted_splits);
  if (q.intermediate_aggregation_result) w.bytes(6, *q.intermediate_aggregation_result);
  w.u64(7, q.num_successful_splits);
  if (q.resource_stats) w.bytes(9, encode_leaf_stats(*q.resource_stats));
  return w.out;
}

std::string encode_lambda_responses(const std::vector<LambdaSingleSplitResult>& v) {
  Writer w;
  for (auto& r : v) {
    Writer s;
    s.str(1, r.split_id);
    if (r.is_error) s.bytes(3, r.error);
    else s.bytes(2, encode_leaf_search_response(r.response));
    w.bytes(2, s.out);
  }
  return w.out;
}

}  // namespace pb
}  // namespace qw
