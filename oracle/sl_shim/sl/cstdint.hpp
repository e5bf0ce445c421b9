// Minimal stand-in for SpaceLand (sl 7.8.2) <sl/cstdint.hpp>.
// TEST INFRASTRUCTURE ONLY: lets oracle/Makefile compile the untouched reference
// sources in /root/reference (the real SL library is not vendored there:
// CMakeLists.txt:32, .gitignore:5-8).  Written from scratch from the call sites.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <iostream>
#include <algorithm>
#include <functional>
#include <limits>
#include <tuple>
namespace sl {
typedef std::uint8_t  uint8_t;
typedef std::uint16_t uint16_t;
typedef std::uint32_t uint32_t;
typedef std::uint64_t uint64_t;
typedef std::int32_t  int32_t;
typedef std::int64_t  int64_t;
template <class T> inline const T& max(const T& a, const T& b) { return (a < b) ? b : a; }
template <class T> inline const T& min(const T& a, const T& b) { return (b < a) ? b : a; }
using std::tie;
}
