// oracle/shim/prelude.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Force-included before every reference translation unit when building oracle/_ref.
// The reference relies on Boost headers to pull these standard headers in transitively
// (strcpy in Estimation/Cell.cpp:18, std::sort/std::iota in CellsDataContainer.cpp:129,265).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
