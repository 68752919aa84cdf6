#ifndef REF_SHIM_DIMENSIONS_HPP
#define REF_SHIM_DIMENSIONS_HPP
#ifndef PLEV
#define PLEV 72
#endif
namespace Homme {
static constexpr int NP = 4;
static constexpr int NUM_PHYSICAL_LEV = PLEV;
}
#endif
