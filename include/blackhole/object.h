// Forwarding header: the reference's include path blackhole/object.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_OBJECT_H_
#define BH8_FWD_OBJECT_H_
#include "blackhole/core/scene_object.h"
#include "blackhole/core/scene.h"
#include "blackhole/core/shapes.h"
#endif  // BH8_FWD_OBJECT_H_
