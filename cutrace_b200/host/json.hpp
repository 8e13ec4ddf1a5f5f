// json.hpp — minimal JSON reader for the scene front-end (the reference uses picojson, which is not
// vendored and not installable offline).  Numbers are parsed with strtod and kept as double, exactly the
// value picojson hands to the reference's `(T)v.get<double>()` casts (inc/json_helpers.hpp:88-93).
// Duplicate object keys: last one wins (std::map assignment, as in picojson's default parse context).
#ifndef CUTRACE_B200_HOST_JSON_HPP
#define CUTRACE_B200_HOST_JSON_HPP
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace cthost {

struct JsonValue {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<JsonValue> arr;
  std::map<std::string, JsonValue> obj;

  bool is_object() const { return kind == Object; }
  bool is_array() const { return kind == Array; }
  bool is_number() const { return kind == Number; }
  bool is_string() const { return kind == String; }
  const JsonValue *find(const std::string &key) const {
    auto it = obj.find(key);
    return it == obj.end() ? nullptr : &it->second;
  }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string &text) : s_(text), p_(0) {}
  bool parse(JsonValue &out, std::string &err) {
    skip_ws();
    if (!value(out, err, 0)) return false;
    skip_ws();
    if (p_ != s_.size()) { err = at("trailing characters after the JSON document"); return false; }
    return true;
  }

 private:
  const std::string &s_;
  size_t p_;

  std::string at(const std::string &msg) const {
    size_t line = 1;
    for (size_t i = 0; i < p_ && i < s_.size(); i++) if (s_[i] == '\n') line++;
    return msg + " (line " + std::to_string(line) + ")";
  }
  void skip_ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) p_++; }
  bool lit(const char *w) {
    size_t n = strlen(w);
    if (s_.compare(p_, n, w) == 0) { p_ += n; return true; }
    return false;
  }
  static void utf8(std::string &o, unsigned cp) {
    if (cp < 0x80) o += (char)cp;
    else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
  }
  bool hex4(unsigned &v) {
    if (p_ + 4 > s_.size()) return false;
    v = 0;
    for (int i = 0; i < 4; i++) {
      char c = s_[p_++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else return false;
    }
    return true;
  }
  bool string(std::string &out, std::string &err) {
    p_++;  // opening quote
    out.clear();
    while (p_ < s_.size()) {
      char c = s_[p_++];
      if (c == '"') return true;
      if (c == '\\') {
        if (p_ >= s_.size()) break;
        char e = s_[p_++];
        switch (e) {
          case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
          case 'r': out += '\r'; break; case 't': out += '\t'; break;
          case 'u': {
            unsigned cp;
            if (!hex4(cp)) { err = at("bad \\u escape"); return false; }
            if (cp >= 0xD800 && cp < 0xDC00 && s_.compare(p_, 2, "\\u") == 0) {
              p_ += 2;
              unsigned lo;
              if (!hex4(lo)) { err = at("bad \\u escape"); return false; }
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            utf8(out, cp);
            break;
          }
          default: err = at("bad escape in string"); return false;
        }
      } else {
        if ((unsigned char)c < 0x20) { err = at("control character in string"); return false; }   // as picojson and RFC 8259
        out += c;
      }
    }
    err = at("unterminated string");
    return false;
  }
  bool value(JsonValue &v, std::string &err, int depth) {
    if (depth > 256) { err = at("nesting too deep"); return false; }
    if (p_ >= s_.size()) { err = at("unexpected end of input"); return false; }
    char c = s_[p_];
    if (c == '{') {
      v.kind = JsonValue::Object;
      p_++; skip_ws();
      if (p_ < s_.size() && s_[p_] == '}') { p_++; return true; }
      for (;;) {
        skip_ws();
        if (p_ >= s_.size() || s_[p_] != '"') { err = at("expected a string key"); return false; }
        std::string key;
        if (!string(key, err)) return false;
        skip_ws();
        if (p_ >= s_.size() || s_[p_] != ':') { err = at("expected ':'"); return false; }
        p_++; skip_ws();
        JsonValue child;
        if (!value(child, err, depth + 1)) return false;
        v.obj[key] = std::move(child);
        skip_ws();
        if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
        if (p_ < s_.size() && s_[p_] == '}') { p_++; return true; }
        err = at("expected ',' or '}'");
        return false;
      }
    }
    if (c == '[') {
      v.kind = JsonValue::Array;
      p_++; skip_ws();
      if (p_ < s_.size() && s_[p_] == ']') { p_++; return true; }
      for (;;) {
        skip_ws();
        JsonValue child;
        if (!value(child, err, depth + 1)) return false;
        v.arr.push_back(std::move(child));
        skip_ws();
        if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
        if (p_ < s_.size() && s_[p_] == ']') { p_++; return true; }
        err = at("expected ',' or ']'");
        return false;
      }
    }
    if (c == '"') { v.kind = JsonValue::String; return string(v.str, err); }
    if (lit("true")) { v.kind = JsonValue::Bool; v.b = true; return true; }
    if (lit("false")) { v.kind = JsonValue::Bool; v.b = false; return true; }
    if (lit("null")) { v.kind = JsonValue::Null; return true; }
    if (c == '-' || (c >= '0' && c <= '9')) {
      // picojson's number rule (the reference's JSON library, CMakeLists.txt:14-17), restated: take the longest run of
      // [0-9+-.eE], and strtod must consume all of it.  Lenient where strict JSON is not ("041", "0.", "1.e5") and strict
      // where a bare strtod is not ("0x10", "-inf": the run ends at 'x' / 'i').
      size_t q = p_;
      while (q < s_.size() && ((s_[q] >= '0' && s_[q] <= '9') || s_[q] == '+' || s_[q] == '-' || s_[q] == '.' || s_[q] == 'e' || s_[q] == 'E')) q++;
      const std::string run = s_.substr(p_, q - p_);
      char *end = nullptr;
      double d = strtod(run.c_str(), &end);
      if (run.empty() || end != run.c_str() + run.size()) { err = at("bad number"); return false; }
      p_ = q;
      v.kind = JsonValue::Number;
      v.num = d;
      return true;
    }
    err = at(std::string("unexpected character '") + c + "'");
    return false;
  }
};

}  // namespace cthost
#endif
