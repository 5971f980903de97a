#pragma once
#include <sstream>
#include <string>
#include <stdexcept>
namespace boost {
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <class Target, class Source> inline Target lexical_cast(const Source& s) {
  std::stringstream ss;
  ss.precision(17);
  Target t;
  if (!(ss << s) || !(ss >> t) || !(ss >> std::ws).eof()) throw bad_lexical_cast();
  return t;
}
}
