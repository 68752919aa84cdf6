// Minimal stand-in for the reference's Types.hpp, only to compile ITS OWN
// src/preqx/unit_tests/remap.cpp (the plain-C++ PPM twin) in place. Test infrastructure.
#ifndef REF_SHIM_TYPES_HPP
#define REF_SHIM_TYPES_HPP
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "Dimensions.hpp"
namespace Homme { using Real = double; }
#endif
