// Forwarding header: the reference's include path blackhole/constants.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_CONSTANTS_H_
#define BH8_FWD_CONSTANTS_H_
#include "blackhole/core/numeric.h"
#endif  // BH8_FWD_CONSTANTS_H_
