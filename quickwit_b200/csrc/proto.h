// proto.h — hand-rolled protobuf wire codec for the quickwit.search messages this path exchanges
// (quickwit-proto/protos/quickwit/search.proto; field numbers cited per struct). There is no
// protoc in the build image, and the ABI payload is bytes-in / bytes-out anyway.
#pragma once
#include <optional>
#include <string>
#include <vector>

#include "common.h"

namespace qw {
namespace pb {

// ---- wire primitives ------------------------------------------------------------------------------
struct Reader {
  const uint8_t *p, *e;
  Reader(const uint8_t* b, size_t n) : p(b), e(b + n) {}
  bool done() const { return p >= e; }
  uint64_t varint() {
    uint64_t v = 0;
    int sh = 0;
    while (p < e) {
      uint8_t c = *p++;
      v |= (uint64_t)(c & 0x7F) << sh;
      if (!(c & 0x80)) return v;
      sh += 7;
      if (sh > 63) break;
    }
    fail(QWGPU_EINVALID_ARG, "malformed protobuf varint");
  }
  uint64_t fixed64() {
    if (e - p < 8) fail(QWGPU_EINVALID_ARG, "truncated protobuf fixed64");
    uint64_t v;
    memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  Reader sub() {
    uint64_t n = varint();
    if ((uint64_t)(e - p) < n) fail(QWGPU_EINVALID_ARG, "truncated protobuf length-delimited field");
    Reader r(p, (size_t)n);
    p += n;
    return r;
  }
  std::string str() {
    Reader r = sub();
    return std::string((const char*)r.p, r.e - r.p);
  }
  void skip(uint32_t wt) {
    switch (wt) {
      case 0: varint(); break;
      case 1: fixed64(); break;
      case 2: sub(); break;
      case 5: if (e - p < 4) fail(QWGPU_EINVALID_ARG, "truncated protobuf fixed32"); p += 4; break;
      default: fail(QWGPU_EINVALID_ARG, "unsupported protobuf wire type %u", wt);
    }
  }
};

struct Writer {
  std::string out;
  void varint(uint64_t v) {
    while (v >= 0x80) { out += (char)(v | 0x80); v >>= 7; }
    out += (char)v;
  }
  void tag(uint32_t field, uint32_t wt) { varint(((uint64_t)field << 3) | wt); }
  void u64(uint32_t field, uint64_t v) { if (v) { tag(field, 0); varint(v); } }
  void u64_always(uint32_t field, uint64_t v) { tag(field, 0); varint(v); }
  void boolean(uint32_t field, bool v) { if (v) { tag(field, 0); varint(1); } }
  void f64_always(uint32_t field, double d) { tag(field, 1); uint64_t b; memcpy(&b, &d, 8); out.append((const char*)&b, 8); }
  void bytes(uint32_t field, const std::string& s) { tag(field, 2); varint(s.size()); out += s; }
  void str(uint32_t field, const std::string& s) { if (!s.empty()) bytes(field, s); }
};

// ---- messages ------------------------------------------------------------------------------------
// SortByValue (search.proto:572-581): oneof u64=1, i64=2, f64=3, boolean=4
struct SortValue {
  enum Kind { None = 0, U64 = 1, I64 = 2, F64 = 3, Bool = 4 } kind = None;
  uint64_t u = 0;
  int64_t i = 0;
  double f = 0;
  bool b = false;
};
// PartialHit (search.proto:543-570): sort_value=10, sort_value2=11, split_id=2, segment_ord=3, doc_id=4.
// has_sv*: the SortByValue message is present (Some(SortByValue{..})) even if its oneof is unset.
struct PartialHit {
  bool has_sv1 = false, has_sv2 = false;
  SortValue sv1, sv2;
  std::string split_id;
  uint32_t segment_ord = 0, doc_id = 0;
};
// SortField (search.proto:259-270): field_name=1, sort_order=2 (ASC=0, DESC=1), sort_datetime_format=3
struct SortField {
  std::string field_name;
  int32_t sort_order = 0;
};
// SearchRequest (search.proto:188-257)
struct SearchRequest {
  std::vector<std::string> index_id_patterns;  // 1
  std::string query_ast;                       // 13
  std::optional<int64_t> start_timestamp, end_timestamp;  // 4, 5
  uint64_t max_hits = 0, start_offset = 0;     // 6, 7
  std::optional<std::string> aggregation_request;  // 11
  std::vector<std::string> snippet_fields;     // 12
  std::vector<SortField> sort_fields;          // 14
  std::optional<uint32_t> scroll_ttl_secs;     // 15
  std::optional<PartialHit> search_after;      // 16
  int32_t count_hits = 0;                      // 17
  bool ignore_missing_indexes = false;         // 18
  bool skip_aggregation_finalization = false;  // 19
};
// SplitIdAndFooterOffsets (search.proto:489-503)
struct SplitIdAndFooterOffsets {
  std::string split_id;  // 1
  uint64_t split_footer_start = 0, split_footer_end = 0;  // 2, 3
  std::optional<int64_t> timestamp_start, timestamp_end;  // 4, 5
  uint64_t num_docs = 0;  // 6
};
// LeafRequestRef (search.proto:477-487)
struct LeafRequestRef {
  uint32_t doc_mapper_ord = 0, index_uri_ord = 0;  // 1, 2
  std::vector<SplitIdAndFooterOffsets> split_offsets;  // 3
};
// LeafSearchRequest (search.proto:343-359)
struct LeafSearchRequest {
  SearchRequest search_request;            // 1
  std::vector<LeafRequestRef> leaf_requests;  // 7
  std::vector<std::string> doc_mappers;    // 8
  std::vector<std::string> index_uris;     // 9
};
// SplitSearchError (search.proto:331-340)
struct SplitSearchError {
  std::string error, split_id;  // 1, 2
  bool retryable_error = false;  // 3
};
// SplitResourceStats (search.proto:364-388), LeafResourceStats (search.proto:390-452)
struct SplitResourceStats {
  uint64_t v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // fields 1..9; [8] = cpu_search_microsecs
};
struct LeafResourceStats {
  uint64_t partial_result_cache_num_splits = 0, partial_result_cache_num_docs = 0;  // 1, 2
  uint64_t localexec_num_splits = 0, localexec_num_docs = 0;                        // 3, 4
  std::optional<SplitResourceStats> split_resources_worst, split_resources_sum;     // 5, 6
  std::optional<uint64_t> min_wait_for_search_permit_microsecs, min_wait_for_cpu_pool_microsecs;  // 7, 8
  uint64_t wall_time_microsecs = 0;  // 9
  uint64_t lambda[5] = {0, 0, 0, 0, 0};  // 10..14
};
// LeafSearchResponse (search.proto:583-613)
struct LeafSearchResponse {
  uint64_t num_hits = 0;                          // 1
  std::vector<PartialHit> partial_hits;           // 2
  std::string encoded_partial_hits;               // 2, already in wire form (follows partial_hits; encode-only)
  std::vector<SplitSearchError> failed_splits;    // 3
  uint64_t num_attempted_splits = 0;              // 4
  std::optional<std::string> intermediate_aggregation_result;  // 6
  uint64_t num_successful_splits = 0;             // 7
  std::optional<LeafResourceStats> resource_stats;  // 9
};
// LambdaSingleSplitResult (search.proto:617-626): split_id=1, oneof {response=2, error=3}
struct LambdaSingleSplitResult {
  std::string split_id;
  bool is_error = false;
  LeafSearchResponse response;
  std::string error;
};

SearchRequest decode_search_request(Reader r);
LeafSearchRequest decode_leaf_search_request(const uint8_t* p, size_t n);
LeafSearchResponse decode_leaf_search_response(const uint8_t* p, size_t n);
PartialHit decode_partial_hit(Reader r);
std::string encode_partial_hit(const PartialHit& h);
void append_partial_hit(std::string& out, uint32_t field, const std::string& split_id, uint32_t segment_ord, uint32_t doc_id,
                        bool has_sv1, const SortValue& sv1, bool has_sv2, const SortValue& sv2);
std::string encode_leaf_search_response(const LeafSearchResponse& r);
std::string encode_search_request(const SearchRequest& r);
// LambdaSearchResponses (search.proto:628-632): repeated LambdaSingleSplitResult split_results = 2
std::string encode_lambda_responses(const std::vector<LambdaSingleSplitResult>& v);

}  // namespace pb
}  // namespace qw
