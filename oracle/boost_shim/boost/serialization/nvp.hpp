#pragma once
#include <boost/serialization/shim_core.hpp>
