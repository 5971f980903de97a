#pragma once
#include <boost/function.hpp>
#include <boost/functional/hash.hpp>
