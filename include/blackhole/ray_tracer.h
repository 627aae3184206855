// Forwarding header: the reference's include path blackhole/ray_tracer.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_RAY_TRACER_H_
#define BH8_FWD_RAY_TRACER_H_
#include "blackhole/core/optics.h"
#endif  // BH8_FWD_RAY_TRACER_H_
