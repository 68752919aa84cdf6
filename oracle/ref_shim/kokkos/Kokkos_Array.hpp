// forwards to the serial stand-in (oracle/ref_shim/kokkos/Kokkos_Core.hpp). TEST INFRASTRUCTURE.
#include "Kokkos_Core.hpp"
