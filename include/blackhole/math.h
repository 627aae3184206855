// Forwarding header: the reference's include path blackhole/math.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_MATH_H_
#define BH8_FWD_MATH_H_
#include "blackhole/core/linear.h"
#endif  // BH8_FWD_MATH_H_
