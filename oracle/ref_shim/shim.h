// Force-included (-include) when compiling the UNMODIFIED reference host sources under
// /root/reference with g++ (see oracle/Makefile). The reference is MSVC-only code; these are
// the missing standard headers and the global min/max it relies on.
#pragma once
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
using std::max;
using std::min;
