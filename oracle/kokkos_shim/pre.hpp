// TEST INFRASTRUCTURE — force-included before every reference translation unit.
// src/device_solver.cpp:351 calls unqualified min(); the reference relies on nvcc's global
// min/max.  Nothing else is injected.
#pragma once
#include <algorithm>
#include <cmath>
using std::max;
using std::min;
