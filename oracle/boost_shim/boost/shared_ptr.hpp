// Minimal stand-in for <boost/shared_ptr.hpp>: maps the names onto <memory>.
// Part of the in-repo Boost shim used ONLY to compile the CPU reference as a test oracle.
#pragma once
#include <memory>
#include <climits>
#include <cstdlib>
#include <cstdio>
namespace boost {
using std::shared_ptr;
using std::weak_ptr;
using std::make_shared;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
using std::const_pointer_cast;
using std::enable_shared_from_this;
}
