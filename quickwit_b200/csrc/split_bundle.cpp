// split_bundle.cpp — reader of a `.split` file's footer (host only).
//
// A split is a bundle: the tantivy files back to back, then
//   [BundleStorageFileOffsets: 8-byte versioned header (magic 403881646, version 1) + JSON {"files": {path: {start, end}}}]
//   [u32 LE length of that component][hotcache][u32 LE length of the hotcache]
// (quickwit-storage/src/bundle_storage.rs:92-174, versioned_component.rs:35-110). The hotcache is a HotDirectory
// image: 8-byte versioned header (magic 2557869106, version 1), u32 LE length, postcard(HotDirectoryMeta{file_lengths:
// map path -> u64, slice_offsets: [(path, u64)]}), then the cached slices (quickwit-directories/src/hot_directory.rs:
// 40-80,167-200). `SplitIdAndFooterOffsets.split_footer_{start,end}` (search.proto:489-503) delimit exactly these bytes:
// one ranged read opens a split (leaf.rs:210-251). This is the first step of real-split ingestion:
// it tells an ingester where the .term / .idx / .pos / .fast / .fieldnorm files of the split lie.
#include <algorithm>

#include "common.h"
#include "json.h"

namespace qw {
namespace {
constexpr uint32_t kBundleMagic = 403881646u, kHotMagic = 2557869106u;

uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

struct Postcard {
  const uint8_t* p; const uint8_t* e;
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
      if (p >= e) fail(QWGPU_EINVALID_ARG, "hotcache metadata truncated");
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
    }
    fail(QWGPU_EINVALID_ARG, "hotcache metadata: varint too long");
  }
  std::string str() {
    const uint64_t n = varint();
    if (n > (uint64_t)(e - p)) fail(QWGPU_EINVALID_ARG, "hotcache metadata truncated");
    std::string s((const char*)p, (size_t)n);
    p += n;
    return s;
  }
};

}  // namespace

// `tail` = the last `len` bytes of the split file (at least the footer); `file_len` = size of the whole split.
std::string parse_split_footer(const uint8_t* tail, uint64_t len, uint64_t file_len) {
  if (len < 8 || len > file_len) fail(QWGPU_EINVALID_ARG, "split footer: need the end of the split file");
  const uint64_t tail_off = file_len - len;  // file offset of tail[0]
  const uint32_t hot_len = rd32(tail + len - 4);
  if ((uint64_t)hot_len + 8 > len) fail(QWGPU_EINVALID_ARG, "split footer: hotcache of %u bytes does not fit the %llu bytes given", hot_len, (unsigned long long)len);
  const uint64_t hot_at = len - 4 - hot_len;           // inside tail
  const uint32_t meta_len = rd32(tail + hot_at - 4);
  if ((uint64_t)meta_len + 4 > hot_at || meta_len < 8) fail(QWGPU_EINVALID_ARG, "split footer: bundle metadata of %u bytes does not fit", meta_len);
  const uint64_t meta_at = hot_at - 4 - meta_len;
  if (rd32(tail + meta_at) != kBundleMagic) fail(QWGPU_EINVALID_ARG, "split footer: bundle metadata magic number does not match");
  if (rd32(tail + meta_at + 4) != 1) fail(QWGPU_EUNSUPPORTED, "split footer: bundle metadata version %u", rd32(tail + meta_at + 4));
  Json meta = parse_json(std::string((const char*)tail + meta_at + 8, meta_len - 8), QWGPU_EINVALID_ARG);
  const Json* files = meta.get("files");
  if (!files || files->type != Json::Obj) fail(QWGPU_EINVALID_ARG, "split footer: no `files` map");
  struct F { std::string path; uint64_t start, end; };
  std::vector<F> fl;
  const uint64_t body_end = tail_off + meta_at;  // the files end where the bundle metadata starts
  for (auto& kv : files->obj) {
    const Json* s = kv.second.get("start");
    const Json* e = kv.second.get("end");
    if (!s || !e) fail(QWGPU_EINVALID_ARG, "split footer: file `%s` has no range", kv.first.c_str());
    F f{kv.first, (uint64_t)s->as_f64(), (uint64_t)e->as_f64()};
    if (s->type == Json::U64) f.start = s->u;
    if (e->type == Json::U64) f.end = e->u;
    if (f.start > f.end || f.end > body_end) fail(QWGPU_EINVALID_ARG, "split footer: file `%s` lies outside the split body", kv.first.c_str());
    fl.push_back(f);
  }
  std::sort(fl.begin(), fl.end(), [](const F& a, const F& b) { return a.start < b.start || (a.start == b.start && a.path < b.path); });
  std::string out = "{\"files\":[";
  for (size_t i = 0; i < fl.size(); i++) {
    if (i) out += ",";
    out += "{\"path\":"; json_escape(fl[i].path, out);
    out += ",\"start\":" + std::to_string(fl[i].start) + ",\"end\":" + std::to_string(fl[i].end) + "}";
  }
  out += "],\"bundle_metadata\":{\"offset\":" + std::to_string(tail_off + meta_at) + ",\"len\":" + std::to_string(meta_len) + "}";
  out += ",\"footer_start\":" + std::to_string(tail_off + meta_at) + ",\"footer_end\":" + std::to_string(file_len);
  out += ",\"hotcache\":{\"offset\":" + std::to_string(tail_off + hot_at) + ",\"len\":" + std::to_string(hot_len);
  if (hot_len >= 12) {
    const uint8_t* h = tail + hot_at;
    if (rd32(h) != kHotMagic) fail(QWGPU_EINVALID_ARG, "split footer: hot directory metadata's magic number does not match");
    if (rd32(h + 4) != 1) fail(QWGPU_EUNSUPPORTED, "split footer: hot directory version %u", rd32(h + 4));
    const uint32_t pc_len = rd32(h + 8);
    if ((uint64_t)pc_len + 12 > hot_len) fail(QWGPU_EINVALID_ARG, "split footer: hot directory metadata truncated");
    Postcard pc{h + 12, h + 12 + pc_len};
    const uint64_t slices_at = tail_off + hot_at + 12 + pc_len;  // slice offsets are relative to the bytes after the metadata
    out += ",\"file_lengths\":{";
    const uint64_t nfl = pc.varint();
    for (uint64_t i = 0; i < nfl; i++) {
      std::string path = pc.str();
      const uint64_t l = pc.varint();
      if (i) out += ",";
      json_escape(path, out); out += ":" + std::to_string(l);
    }
    out += "},\"slices\":[";
    const uint64_t nso = pc.varint();
    for (uint64_t i = 0; i < nso; i++) {
      std::string path = pc.str();
      const uint64_t o = pc.varint();
      if (i) out += ",";
      out += "{\"path\":"; json_escape(path, out); out += ",\"offset\":" + std::to_string(slices_at + o) + "}";
    }
    out += "]";
  }
  out += "}}";
  return out;
}
}  // namespace qw

extern "C" int qwgpu_parse_split_footer(const uint8_t* tail, uint64_t tail_len, uint64_t split_file_len, uint8_t** json_out, size_t* json_len) {
  try {
    if (!tail || !json_out || !json_len) qw::fail(QWGPU_EINVALID_ARG, "null argument");
    std::string js = qw::parse_split_footer(tail, tail_len, split_file_len);
    *json_out = (uint8_t*)malloc(js.size() ? js.size() : 1);
    if (!*json_out) qw::fail(QWGPU_EINTERNAL, "out of memory");
    memcpy(*json_out, js.data(), js.size());
    *json_len = js.size();
    return 0;
  } catch (const qw::Error& e) {
    qw::set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    qw::set_last_error(e.what());
    return QWGPU_EINTERNAL;
  }
}
