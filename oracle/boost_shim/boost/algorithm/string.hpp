// Minimal stand-in for <boost/algorithm/string.hpp> (iequals, split, is_any_of, trim, to_lower...).
#pragma once
#include <string>
#include <vector>
#include <cctype>
#include <algorithm>
namespace boost {
namespace algorithm {
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct is_any_of {
  std::string set;
  is_any_of(const std::string& s) : set(s) {}
  is_any_of(const char* s) : set(s) {}
  bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline bool iequals(const std::string& a, const std::string& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return true;
}
template <class Seq, class Pred>
inline Seq& split(Seq& out, const std::string& in, Pred pred, token_compress_mode_type mode = token_compress_off) {
  out.clear();
  std::string cur;
  bool prev_delim = false;
  for (size_t i = 0; i < in.size(); ++i) {
    if (pred(in[i])) {
      if (mode == token_compress_on && prev_delim) continue;
      out.push_back(cur); cur.clear(); prev_delim = true;
    } else { cur += in[i]; prev_delim = false; }
  }
  out.push_back(cur);
  return out;
}
inline void trim_left(std::string& s) { size_t i = 0; while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; s.erase(0, i); }
inline void trim_right(std::string& s) { size_t n = s.size(); while (n > 0 && std::isspace((unsigned char)s[n - 1])) --n; s.erase(n); }
inline void trim(std::string& s) { trim_right(s); trim_left(s); }
inline std::string trim_copy(std::string s) { trim(s); return s; }
inline void to_lower(std::string& s) { for (size_t i = 0; i < s.size(); ++i) s[i] = std::tolower((unsigned char)s[i]); }
inline void to_upper(std::string& s) { for (size_t i = 0; i < s.size(); ++i) s[i] = std::toupper((unsigned char)s[i]); }
inline std::string to_lower_copy(std::string s) { to_lower(s); return s; }
inline std::string to_upper_copy(std::string s) { to_upper(s); return s; }
inline bool starts_with(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }
inline bool contains(const std::string& s, const std::string& p) { return s.find(p) != std::string::npos; }
}  // namespace algorithm
using algorithm::token_compress_on;
using algorithm::token_compress_off;
using algorithm::is_any_of;
using algorithm::iequals;
using algorithm::split;
using algorithm::trim;
using algorithm::trim_left;
using algorithm::trim_right;
using algorithm::trim_copy;
using algorithm::to_lower;
using algorithm::to_upper;
using algorithm::to_lower_copy;
using algorithm::to_upper_copy;
using algorithm::starts_with;
using algorithm::contains;
}
