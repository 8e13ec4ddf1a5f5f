// Stub of picojson for the oracle build ONLY: just enough surface for the reference's loader
// templates (inc/loader.hpp, inc/json_helpers.hpp) to parse. None of it is instantiated by the
// render path; parse() is never called.
#ifndef ORACLE_STUB_PICOJSON_H
#define ORACLE_STUB_PICOJSON_H
#include <iostream>
#include <map>
#include <string>
#include <vector>
namespace picojson {
class value;
typedef std::vector<value> array;
typedef std::map<std::string, value> object;
class value {
public:
  value() {}
  explicit value(double) {}
  template <typename T> bool is() const { return false; }
  template <typename T> const T &get() const { static T t{}; return t; }
  template <typename T> T &get() { static T t{}; return t; }
};
inline std::string parse(value &, std::istream &) { return "stub"; }
inline const std::string &get_last_error() { static std::string s; return s; }
}
#endif
