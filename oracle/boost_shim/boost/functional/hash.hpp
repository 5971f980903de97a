#pragma once
#include <functional>
#include <cstddef>
namespace boost {
template <class T> struct hash : std::hash<T> {};
template <class T> inline void hash_combine(std::size_t& seed, const T& v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
template <class It> inline std::size_t hash_range(It a, It b) {
  std::size_t seed = 0;
  for (; a != b; ++a) hash_combine(seed, *a);
  return seed;
}
}
