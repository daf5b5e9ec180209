// json.h — minimal JSON value / parser / writer for the host side (QueryAst, aggregation requests,
// doc mapper JSON; serde_json's role in the reference). Numbers keep their integer-ness the way
// serde_json::Number does (u64 / i64 / f64).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.h"

namespace qw {

struct Json {
  enum Type { Null, Bool, U64, I64, F64, Str, Arr, Obj } type = Null;
  bool b = false;
  uint64_t u = 0;
  int64_t i = 0;
  double f = 0;
  std::string s;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order preserved

  bool is_null() const { return type == Null; }
  bool is_num() const { return type == U64 || type == I64 || type == F64; }
  bool is_str() const { return type == Str; }
  bool is_obj() const { return type == Obj; }
  bool is_arr() const { return type == Arr; }
  double as_f64() const { return type == U64 ? (double)u : (type == I64 ? (double)i : f); }
  const Json* get(const std::string& k) const {
    if (type != Obj) return nullptr;
    for (auto& kv : obj) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  std::string str_or(const std::string& k, const std::string& d) const {
    const Json* v = get(k);
    return v && v->type == Str ? v->s : d;
  }
  bool bool_or(const std::string& k, bool d) const {
    const Json* v = get(k);
    return v && v->type == Bool ? v->b : d;
  }
};

class JsonParser {
 public:
  JsonParser(const char* p, size_t n, int err_code) : p_(p), e_(p + n), code_(err_code) {}
  Json parse() {
    Json v = value(0);
    ws();
    if (p_ != e_) bad("trailing characters");
    return v;
  }

 private:
  const char *p_, *e_;
  int code_;
  [[noreturn]] void bad(const char* m) { fail(code_, "invalid JSON: %s", m); }
  void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) p_++; }
  bool lit(const char* t) {
    size_t n = strlen(t);
    if ((size_t)(e_ - p_) >= n && memcmp(p_, t, n) == 0) { p_ += n; return true; }
    return false;
  }
  Json value(int depth) {
    if (depth > 64) bad("nesting too deep");
    ws();
    if (p_ >= e_) bad("unexpected end");
    Json v;
    char c = *p_;
    if (c == '{') {
      p_++;
      v.type = Json::Obj;
      ws();
      if (p_ < e_ && *p_ == '}') { p_++; return v; }
      for (;;) {
        ws();
        if (p_ >= e_ || *p_ != '"') bad("expected object key");
        std::string k = string();
        ws();
        if (p_ >= e_ || *p_ != ':') bad("expected ':'");
        p_++;
        v.obj.emplace_back(std::move(k), value(depth + 1));
        ws();
        if (p_ < e_ && *p_ == ',') { p_++; continue; }
        if (p_ < e_ && *p_ == '}') { p_++; break; }
        bad("expected ',' or '}'");
      }
    } else if (c == '[') {
      p_++;
      v.type = Json::Arr;
      ws();
      if (p_ < e_ && *p_ == ']') { p_++; return v; }
      for (;;) {
        v.arr.push_back(value(depth + 1));
        ws();
        if (p_ < e_ && *p_ == ',') { p_++; continue; }
        if (p_ < e_ && *p_ == ']') { p_++; break; }
        bad("expected ',' or ']'");
      }
    } else if (c == '"') {
      v.type = Json::Str;
      v.s = string();
    } else if (lit("true")) { v.type = Json::Bool; v.b = true; }
    else if (lit("false")) { v.type = Json::Bool; v.b = false; }
    else if (lit("null")) { v.type = Json::Null; }
    else v = number();
    return v;
  }
  std::string string() {
    std::string out;
    p_++;  // opening quote
    while (p_ < e_ && *p_ != '"') {
      if (*p_ == '\\') {
        p_++;
        if (p_ >= e_) bad("bad escape");
        char c = *p_++;
        switch (c) {
          case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break; case '/': out += '/'; break;
          case '\\': out += '\\'; break; case '"': out += '"'; break;
          case 'u': {
            if (e_ - p_ < 4) bad("bad \\u escape");
            unsigned cp = (unsigned)strtoul(std::string(p_, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp >= 0xD800 && cp < 0xDC00 && e_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
              unsigned lo = (unsigned)strtoul(std::string(p_ + 2, 4).c_str(), nullptr, 16);
              p_ += 6;
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: bad("bad escape");
        }
      } else out += *p_++;
    }
    if (p_ >= e_) bad("unterminated string");
    p_++;
    return out;
  }
  Json number() {
    const char* s = p_;
    bool is_float = false;
    if (p_ < e_ && (*p_ == '-' || *p_ == '+')) p_++;
    while (p_ < e_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) {
      if (*p_ == '.' || *p_ == 'e' || *p_ == 'E') is_float = true;
      p_++;
    }
    if (p_ == s) bad("unexpected character");
    std::string t(s, p_ - s);
    Json v;
    if (!is_float) {
      errno = 0;
      if (t[0] == '-') {
        long long x = strtoll(t.c_str(), nullptr, 10);
        if (errno == 0) { v.type = Json::I64; v.i = x; return v; }
      } else {
        unsigned long long x = strtoull(t.c_str(), nullptr, 10);
        if (errno == 0) { v.type = Json::U64; v.u = x; return v; }
      }
    }
    v.type = Json::F64;
    v.f = strtod(t.c_str(), nullptr);
    return v;
  }
};

inline Json parse_json(const std::string& s, int err_code) { return JsonParser(s.data(), s.size(), err_code).parse(); }

inline void json_escape(const std::string& s, std::string& out) {
  out += '"';
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break; case '\\': out += "\\\\"; break; case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break; case '\t': out += "\\t"; break;
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
        else out += (char)c;
    }
  }
  out += '"';
}
// f64 formatting like serde_json (shortest round-trip, always with a fractional part or exponent)
inline void json_f64(double d, std::string& out) {
  if (!std::isfinite(d)) { out += "null"; return; }
  char b[40];
  for (int prec = 1; prec <= 17; prec++) {
    snprintf(b, sizeof b, "%.*g", prec, d);
    if (strtod(b, nullptr) == d) break;
  }
  std::string t = b;
  if (t.find_first_of(".eEn") == std::string::npos) t += ".0";
  out += t;
}

}  // namespace qw
