// Forwarding header: the reference's include path blackhole/object/object.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_OBJECT_OBJECT_H_
#define BH8_FWD_OBJECT_OBJECT_H_
#include "blackhole/core/scene_object.h"
#endif  // BH8_FWD_OBJECT_OBJECT_H_
