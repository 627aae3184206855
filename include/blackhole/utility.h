// Forwarding header: the reference's include path blackhole/utility.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_UTILITY_H_
#define BH8_FWD_UTILITY_H_
#include "blackhole/core/numeric.h"
#endif  // BH8_FWD_UTILITY_H_
