#pragma once
#include <boost/shared_ptr.hpp>
#include <memory>
#include <cstddef>
namespace boost {
template <class T> class shared_array {
  std::shared_ptr<T> p_;
 public:
  shared_array() {}
  explicit shared_array(T* p) : p_(p, std::default_delete<T[]>()) {}
  void reset(T* p = 0) { if (p) p_.reset(p, std::default_delete<T[]>()); else p_.reset(); }
  T& operator[](std::ptrdiff_t i) const { return p_.get()[i]; }
  T* get() const { return p_.get(); }
  explicit operator bool() const { return bool(p_); }
};
}
