// Forwarding header: the reference's include path blackhole/object/pattern.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_OBJECT_PATTERN_H_
#define BH8_FWD_OBJECT_PATTERN_H_
#include "blackhole/core/shapes.h"
#endif  // BH8_FWD_OBJECT_PATTERN_H_
