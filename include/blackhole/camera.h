// Forwarding header: the reference's include path blackhole/camera.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_CAMERA_H_
#define BH8_FWD_CAMERA_H_
#include "blackhole/core/optics.h"
#endif  // BH8_FWD_CAMERA_H_
