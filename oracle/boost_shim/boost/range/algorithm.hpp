#pragma once
#include <algorithm>
namespace boost {
template <class R> inline R& sort(R& r) { std::sort(r.begin(), r.end()); return r; }
template <class R, class C> inline R& sort(R& r, C c) { std::sort(r.begin(), r.end(), c); return r; }
namespace range { using boost::sort; }
}
