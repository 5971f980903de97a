#pragma once
#include <unordered_map>
namespace std { namespace tr1 { using std::unordered_map; using std::hash; } }
