// Minimal stand-in for <boost/format.hpp>: printf-style directives fed with operator%.
// Supports the subset the reference uses: %s %d %i %u %f %e %g with flags/width/precision.
#pragma once
#include <boost/shared_ptr.hpp>
#include <string>
#include <sstream>
#include <vector>
#include <iomanip>
#include <ostream>
namespace boost {
class format {
  struct piece { std::string lit; std::string spec; };  // literal text followed by one directive
  std::vector<piece> pieces_;
  std::string tail_;
  std::vector<std::string> args_;
  static void apply_spec(std::ostream& os, const std::string& spec) {
    // spec is like "-20.14e" (without the leading %)
    size_t i = 0;
    bool left = false, zero = false, plus = false;
    for (; i < spec.size(); ++i) {
      char c = spec[i];
      if (c == '-') left = true; else if (c == '0') zero = true; else if (c == '+') plus = true;
      else if (c == ' ' || c == '#') {} else break;
    }
    int width = 0; while (i < spec.size() && isdigit((unsigned char)spec[i])) width = width * 10 + (spec[i++] - '0');
    int prec = -1;
    if (i < spec.size() && spec[i] == '.') { ++i; prec = 0; while (i < spec.size() && isdigit((unsigned char)spec[i])) prec = prec * 10 + (spec[i++] - '0'); }
    while (i < spec.size() && (spec[i] == 'l' || spec[i] == 'h' || spec[i] == 'z')) ++i;
    char conv = i < spec.size() ? spec[i] : 's';
    if (left) os << std::left; else os << std::right;
    if (zero && !left) os << std::setfill('0');
    if (plus) os << std::showpos;
    if (width) os << std::setw(width);
    if (prec >= 0) os << std::setprecision(prec);
    switch (conv) {
      case 'f': case 'F': os << std::fixed; break;
      case 'e': os << std::scientific; break;
      case 'E': os << std::scientific << std::uppercase; break;
      default: break;
    }
  }
 public:
  explicit format(const std::string& f) {
    std::string lit;
    for (size_t i = 0; i < f.size(); ++i) {
      if (f[i] != '%') { lit += f[i]; continue; }
      if (i + 1 < f.size() && f[i + 1] == '%') { lit += '%'; ++i; continue; }
      size_t j = i + 1;
      while (j < f.size() && !isalpha((unsigned char)f[j])) ++j;
      while (j < f.size() && (f[j] == 'l' || f[j] == 'h' || f[j] == 'z')) ++j;
      piece p; p.lit = lit; p.spec = f.substr(i + 1, j - i);
      pieces_.push_back(p); lit.clear(); i = j;
    }
    tail_ = lit;
  }
  template <class T> format& operator%(const T& v) {
    std::ostringstream os;
    if (args_.size() < pieces_.size()) apply_spec(os, pieces_[args_.size()].spec);
    os << v;
    args_.push_back(os.str());
    return *this;
  }
  std::string str() const {
    std::string out;
    for (size_t i = 0; i < pieces_.size(); ++i) { out += pieces_[i].lit; if (i < args_.size()) out += args_[i]; }
    out += tail_;
    return out;
  }
};
inline std::string str(const format& f) { return f.str(); }
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
}
