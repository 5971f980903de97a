// In-repo stand-in for Boost.Serialization, written from scratch for ONE purpose:
// compiling the CPU reference (sanshar/Block, -DSERIAL) as a test oracle in an image without Boost.
// The archive format only needs to round-trip with itself; it is NOT Boost's format.
// Supported: arithmetic/enum, std::string, vector/list/set/map/multimap/pair, C arrays, raw-pointer-free
// class types through member or free serialize(), split save/load, base_object, shared_ptr with
// object tracking and polymorphic types registered through Archive::register_type.
#pragma once
#include <boost/shared_ptr.hpp>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <typeindex>
#include <typeinfo>
#include <utility>
#include <vector>

namespace boost {
namespace archive {
class shim_oarchive;
class shim_iarchive;
}
namespace serialization {

struct version_type {
  unsigned v;
  explicit version_type(unsigned x = 0) : v(x) {}
  operator unsigned int() const { return v; }
};

class access {
 public:
  template <class Ar, class T> static void serialize(Ar& ar, T& t, const unsigned int v) { t.serialize(ar, v); }
  template <class Ar, class T> static void member_save(Ar& ar, const T& t, const unsigned int v) { t.save(ar, v); }
  template <class Ar, class T> static void member_load(Ar& ar, T& t, const unsigned int v) { t.load(ar, v); }
  template <class T> static T* construct() { return new T(); }
  template <class T> static void destroy(const T* t) { delete const_cast<T*>(t); }
};

// default free serialize -> member serialize; user headers add more specialised overloads in this namespace
template <class Ar, class T> inline void serialize(Ar& ar, T& t, const unsigned int v) { access::serialize(ar, t, v); }

template <class Ar, class T> inline void split_member(Ar& ar, T& t, const unsigned int v) {
  if (Ar::is_saving::value) access::member_save(ar, t, v); else access::member_load(ar, t, v);
}
template <class Ar, class T> inline void save(Ar&, const T&, const unsigned int);
template <class Ar, class T> inline void load(Ar&, T&, const unsigned int);
template <class Ar, class T> inline void split_free(Ar& ar, T& t, const unsigned int v);

// ---- up-cast registry: filled in lazily by base_object<Base>(derived) ----
struct cast_registry {
  typedef void* (*cast_fn)(void*);
  static std::map<std::pair<std::type_index, std::type_index>, cast_fn>& table() {
    static std::map<std::pair<std::type_index, std::type_index>, cast_fn> t;
    return t;
  }
  static void* upcast(void* p, std::type_index from, std::type_index to, int depth = 0) {
    if (from == to) return p;
    auto& t = table();
    auto it = t.find(std::make_pair(from, to));
    if (it != t.end()) return it->second(p);
    if (depth < 4)
      for (auto& kv : t)
        if (kv.first.first == from) {
          void* r = upcast(kv.second(p), kv.first.second, to, depth + 1);
          if (r) return r;
        }
    return nullptr;
  }
};
template <class Derived, class Base> struct caster {
  static void* up(void* p) { return static_cast<Base*>(static_cast<Derived*>(p)); }
  static void ensure() {
    static bool done = (cast_registry::table()[std::make_pair(std::type_index(typeid(Derived)), std::type_index(typeid(Base)))] = &up, true);
    (void)done;
  }
};
template <class Base, class Derived> inline Base& base_object(Derived& d) {
  caster<typename std::remove_const<Derived>::type, typename std::remove_const<Base>::type>::ensure();
  return static_cast<Base&>(d);
}

template <class T> struct nvp_wrap { T& t; };
template <class T> inline T& make_nvp(const char*, T& t) { return t; }

// ---- polymorphic type registry (Archive::register_type) ----
struct poly_entry {
  std::string key;
  void (*save)(archive::shim_oarchive&, const void*);
  std::shared_ptr<void> (*create)();                 // default-constructed most-derived object
  void (*load)(archive::shim_iarchive&, void*);
  std::type_index ti;
  poly_entry() : save(nullptr), create(nullptr), load(nullptr), ti(typeid(void)) {}
};
struct poly_registry {
  static std::map<std::type_index, poly_entry>& by_type() { static std::map<std::type_index, poly_entry> m; return m; }
  static std::map<std::string, poly_entry>& by_key() { static std::map<std::string, poly_entry> m; return m; }
};

}  // namespace serialization

namespace archive {

class archive_exception : public std::runtime_error {
 public:
  explicit archive_exception(const std::string& s) : std::runtime_error(s) {}
};

enum archive_flags { no_header = 1, no_codecvt = 2, no_xml_tag_checking = 4, no_tracking = 8 };

class shim_oarchive {
  std::ostream& os_;
  std::map<const void*, uint32_t> tracked_;
 public:
  typedef std::true_type is_saving;
  typedef std::false_type is_loading;
  explicit shim_oarchive(std::ostream& os, unsigned = 0) : os_(os) {}
  void raw(const void* p, size_t n) { os_.write(static_cast<const char*>(p), (std::streamsize)n); }
  void save_binary(const void* p, size_t n) { raw(p, n); }
  template <class T> void register_type(const T* = nullptr);
  template <class T> void register_type();

  template <class T> shim_oarchive& operator<<(const T& t) { save(t); return *this; }
  template <class T> shim_oarchive& operator&(const T& t) { save(t); return *this; }

  // --- dispatch ---
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value || std::is_enum<T>::value>::type
  save(const T& t) { raw(&t, sizeof(T)); }
  void save(const std::string& s) { uint64_t n = s.size(); raw(&n, 8); raw(s.data(), n); }
  template <class T, size_t N> void save(const T (&a)[N]) { for (size_t i = 0; i < N; ++i) save(a[i]); }
  template <class T, class A> void save(const std::vector<T, A>& v) {
    uint64_t n = v.size(); raw(&n, 8);
    save_vec(v, std::integral_constant<bool, std::is_arithmetic<T>::value && !std::is_same<T, bool>::value>());
  }
  template <class T, class A> void save_vec(const std::vector<T, A>& v, std::true_type) { if (!v.empty()) raw(v.data(), v.size() * sizeof(T)); }
  template <class T, class A> void save_vec(const std::vector<T, A>& v, std::false_type) { for (size_t i = 0; i < v.size(); ++i) { T tmp = v[i]; (void)tmp; save_elem(v, i); } }
  template <class A> void save_elem(const std::vector<bool, A>& v, size_t i) { bool b = v[i]; save(b); }
  template <class T, class A> void save_elem(const std::vector<T, A>& v, size_t i) { save(v[i]); }
  template <class T, class A> void save(const std::list<T, A>& v) { uint64_t n = v.size(); raw(&n, 8); for (auto& x : v) save(x); }
  template <class T, class C, class A> void save(const std::set<T, C, A>& v) { uint64_t n = v.size(); raw(&n, 8); for (auto& x : v) save(x); }
  template <class K, class V, class C, class A> void save(const std::map<K, V, C, A>& m) { uint64_t n = m.size(); raw(&n, 8); for (auto& kv : m) { save(kv.first); save(kv.second); } }
  template <class K, class V, class C, class A> void save(const std::multimap<K, V, C, A>& m) { uint64_t n = m.size(); raw(&n, 8); for (auto& kv : m) { save(kv.first); save(kv.second); } }
  template <class A, class B> void save(const std::pair<A, B>& p) { save(p.first); save(p.second); }
  template <class T> void save(const std::shared_ptr<T>& p);
  template <class T> typename std::enable_if<std::is_class<T>::value>::type
  save(const T& t) {
    using boost::serialization::serialize;
    serialize(*this, const_cast<T&>(t), boost::serialization::version_type(0));
  }
};

class shim_iarchive {
  std::istream& is_;
  struct tracked { std::shared_ptr<void> sp; std::type_index ti; tracked() : ti(typeid(void)) {} };
  std::vector<tracked> tracked_;
 public:
  typedef std::false_type is_saving;
  typedef std::true_type is_loading;
  explicit shim_iarchive(std::istream& is, unsigned = 0) : is_(is) {}
  void raw(void* p, size_t n) {
    is_.read(static_cast<char*>(p), (std::streamsize)n);
    if ((size_t)is_.gcount() != n) throw archive_exception("shim_iarchive: input stream error");
  }
  void load_binary(void* p, size_t n) { raw(p, n); }
  template <class T> void register_type(const T* = nullptr);
  template <class T> void register_type();

  template <class T> shim_iarchive& operator>>(T& t) { load(t); return *this; }
  template <class T> shim_iarchive& operator&(T& t) { load(t); return *this; }

  template <class T> typename std::enable_if<std::is_arithmetic<T>::value || std::is_enum<T>::value>::type
  load(T& t) { raw(&t, sizeof(T)); }
  void load(std::string& s) { uint64_t n; raw(&n, 8); s.resize(n); if (n) raw(&s[0], n); }
  template <class T, size_t N> void load(T (&a)[N]) { for (size_t i = 0; i < N; ++i) load(a[i]); }
  template <class T, class A> void load(std::vector<T, A>& v) {
    uint64_t n; raw(&n, 8); v.clear(); v.resize(n);
    load_vec(v, std::integral_constant<bool, std::is_arithmetic<T>::value && !std::is_same<T, bool>::value>());
  }
  template <class T, class A> void load_vec(std::vector<T, A>& v, std::true_type) { if (!v.empty()) raw(v.data(), v.size() * sizeof(T)); }
  template <class T, class A> void load_vec(std::vector<T, A>& v, std::false_type) { for (size_t i = 0; i < v.size(); ++i) load_elem(v, i); }
  template <class A> void load_elem(std::vector<bool, A>& v, size_t i) { bool b; load(b); v[i] = b; }
  template <class T, class A> void load_elem(std::vector<T, A>& v, size_t i) { load(v[i]); }
  template <class T, class A> void load(std::list<T, A>& v) { uint64_t n; raw(&n, 8); v.clear(); for (uint64_t i = 0; i < n; ++i) { v.emplace_back(); load(v.back()); } }
  template <class T, class C, class A> void load(std::set<T, C, A>& v) { uint64_t n; raw(&n, 8); v.clear(); for (uint64_t i = 0; i < n; ++i) { T x; load(x); v.insert(v.end(), x); } }
  template <class K, class V, class C, class A> void load(std::map<K, V, C, A>& m) { uint64_t n; raw(&n, 8); m.clear(); for (uint64_t i = 0; i < n; ++i) { K k; load(k); V v; load(v); m.insert(m.end(), std::make_pair(k, v)); } }
  template <class K, class V, class C, class A> void load(std::multimap<K, V, C, A>& m) { uint64_t n; raw(&n, 8); m.clear(); for (uint64_t i = 0; i < n; ++i) { K k; load(k); V v; load(v); m.insert(m.end(), std::make_pair(k, v)); } }
  template <class A, class B> void load(std::pair<A, B>& p) { load(const_cast<typename std::remove_const<A>::type&>(p.first)); load(p.second); }
  template <class T> void load(std::shared_ptr<T>& p);
  template <class T> typename std::enable_if<std::is_class<T>::value>::type
  load(T& t) {
    using boost::serialization::serialize;
    serialize(*this, t, boost::serialization::version_type(0));
  }
};

namespace shim_detail {
template <class T> struct poly_fns {
  static void save(shim_oarchive& ar, const void* p) { ar.save(*static_cast<const T*>(p)); }
  static std::shared_ptr<void> create() { return std::shared_ptr<void>(std::shared_ptr<T>(boost::serialization::access::construct<T>())); }
  static void load(shim_iarchive& ar, void* p) { ar.load(*static_cast<T*>(p)); }
  static void ensure() {
    auto& bt = boost::serialization::poly_registry::by_type();
    std::type_index ti(typeid(T));
    if (bt.count(ti)) return;
    boost::serialization::poly_entry e;
    e.key = typeid(T).name(); e.save = &save; e.create = &create; e.load = &load; e.ti = ti;
    bt[ti] = e;
    boost::serialization::poly_registry::by_key()[e.key] = e;
  }
};
template <class T, bool Poly = std::is_polymorphic<T>::value> struct most_derived {
  static const void* ptr(const T* p) { return dynamic_cast<const void*>(p); }
  static std::type_index type(const T* p) { return std::type_index(typeid(*p)); }
};
template <class T> struct most_derived<T, false> {
  static const void* ptr(const T* p) { return p; }
  static std::type_index type(const T*) { return std::type_index(typeid(T)); }
};
// default-construct T when it is concrete; abstract bases can only arrive through the registry
template <class T, bool Abstract = std::is_abstract<T>::value> struct maker {
  static std::shared_ptr<T> make() { return std::shared_ptr<T>(boost::serialization::access::construct<T>()); }
};
template <class T> struct maker<T, true> {
  static std::shared_ptr<T> make() { throw archive_exception("shim: cannot construct abstract type"); }
};
template <class T, bool Abstract = std::is_abstract<T>::value> struct direct_io {
  static void save(shim_oarchive& ar, const T& t) { ar.save(t); }
};
template <class T> struct direct_io<T, true> {
  static void save(shim_oarchive&, const T&) { throw archive_exception("shim: unregistered derived type behind abstract base"); }
};
}  // namespace shim_detail

template <class T> inline void shim_oarchive::register_type(const T*) { shim_detail::poly_fns<T>::ensure(); }
template <class T> inline void shim_oarchive::register_type() { shim_detail::poly_fns<T>::ensure(); }
template <class T> inline void shim_iarchive::register_type(const T*) { shim_detail::poly_fns<T>::ensure(); }
template <class T> inline void shim_iarchive::register_type() { shim_detail::poly_fns<T>::ensure(); }

template <class T> inline void shim_oarchive::save(const std::shared_ptr<T>& p) {
  typedef typename std::remove_const<T>::type U;
  uint8_t tag;
  if (!p) { tag = 0; raw(&tag, 1); return; }
  const void* key = shim_detail::most_derived<U>::ptr(p.get());
  auto it = tracked_.find(key);
  if (it != tracked_.end()) { tag = 2; raw(&tag, 1); raw(&it->second, 4); return; }
  uint32_t id = (uint32_t)tracked_.size();
  tracked_[key] = id;
  tag = 1; raw(&tag, 1);
  std::type_index dyn = shim_detail::most_derived<U>::type(p.get());
  if (dyn == std::type_index(typeid(U))) {
    save(std::string());
    shim_detail::direct_io<U>::save(*this, *p);
  } else {
    auto& bt = boost::serialization::poly_registry::by_type();
    auto e = bt.find(dyn);
    if (e == bt.end()) throw archive_exception(std::string("shim_oarchive: unregistered class ") + dyn.name());
    save(e->second.key);
    e->second.save(*this, key);
  }
}

template <class T> inline void shim_iarchive::load(std::shared_ptr<T>& p) {
  typedef typename std::remove_const<T>::type U;
  uint8_t tag; raw(&tag, 1);
  if (tag == 0) { p.reset(); return; }
  if (tag == 2) {
    uint32_t id; raw(&id, 4);
    if (id >= tracked_.size()) throw archive_exception("shim_iarchive: bad tracking id");
    void* b = boost::serialization::cast_registry::upcast(tracked_[id].sp.get(), tracked_[id].ti, std::type_index(typeid(U)));
    if (!b) throw archive_exception("shim_iarchive: no upcast for tracked pointer");
    p = std::shared_ptr<T>(tracked_[id].sp, static_cast<U*>(b));
    return;
  }
  std::string key; load(key);
  size_t slot = tracked_.size();
  tracked_.emplace_back();
  if (key.empty()) {
    std::shared_ptr<U> sp = shim_detail::maker<U>::make();
    tracked_[slot].sp = sp; tracked_[slot].ti = std::type_index(typeid(U));
    load(*sp);
    p = sp;
  } else {
    auto& bk = boost::serialization::poly_registry::by_key();
    auto e = bk.find(key);
    if (e == bk.end()) throw archive_exception("shim_iarchive: unregistered class key " + key);
    std::shared_ptr<void> sp = e->second.create();
    tracked_[slot].sp = sp; tracked_[slot].ti = e->second.ti;
    e->second.load(*this, sp.get());
    void* b = boost::serialization::cast_registry::upcast(sp.get(), e->second.ti, std::type_index(typeid(U)));
    if (!b) throw archive_exception("shim_iarchive: no upcast " + key);
    p = std::shared_ptr<T>(sp, static_cast<U*>(b));
  }
}

class binary_oarchive : public shim_oarchive { public: explicit binary_oarchive(std::ostream& os, unsigned f = 0) : shim_oarchive(os, f) {} };
class binary_iarchive : public shim_iarchive { public: explicit binary_iarchive(std::istream& is, unsigned f = 0) : shim_iarchive(is, f) {} };
class text_oarchive : public shim_oarchive { public: explicit text_oarchive(std::ostream& os, unsigned f = 0) : shim_oarchive(os, f) {} };
class text_iarchive : public shim_iarchive { public: explicit text_iarchive(std::istream& is, unsigned f = 0) : shim_iarchive(is, f) {} };

}  // namespace archive

namespace serialization {
template <class Ar, class T> inline void split_free(Ar& ar, T& t, const unsigned int v) {
  if (Ar::is_saving::value) save(ar, t, v); else load(ar, t, v);
}
}
}  // namespace boost

#define BOOST_SERIALIZATION_SPLIT_MEMBER()                                        \
  template <class Archive> void serialize(Archive& ar, const unsigned int version) { \
    boost::serialization::split_member(ar, *this, version);                       \
  }
#define BOOST_CLASS_EXPORT(T)
#define BOOST_CLASS_EXPORT_GUID(T, K)
#define BOOST_CLASS_EXPORT_KEY(T)
#define BOOST_CLASS_EXPORT_IMPLEMENT(T)
#define BOOST_CLASS_VERSION(T, N)
#define BOOST_SERIALIZATION_ASSUME_ABSTRACT(T)
#define BOOST_CLASS_TRACKING(T, E)
#define BOOST_SERIALIZATION_NVP(x) x
