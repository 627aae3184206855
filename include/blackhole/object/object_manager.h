// Forwarding header: the reference's include path blackhole/object/object_manager.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_OBJECT_OBJECT_MANAGER_H_
#define BH8_FWD_OBJECT_OBJECT_MANAGER_H_
#include "blackhole/core/scene.h"
#endif  // BH8_FWD_OBJECT_OBJECT_MANAGER_H_
