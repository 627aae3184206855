// Forwarding header: the reference's include path blackhole/blackhole_solution.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_BLACKHOLE_SOLUTION_H_
#define BH8_FWD_BLACKHOLE_SOLUTION_H_
#include "blackhole/core/schwarzschild.h"
#endif  // BH8_FWD_BLACKHOLE_SOLUTION_H_
