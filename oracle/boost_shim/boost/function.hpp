#pragma once
#include <boost/shared_ptr.hpp>
#include <functional>
namespace boost {
using std::function;
using std::ref;
using std::cref;
using std::reference_wrapper;
}
