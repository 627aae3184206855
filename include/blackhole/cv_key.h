// Forwarding header: the reference's include path blackhole/cv_key.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_CV_KEY_H_
#define BH8_FWD_CV_KEY_H_
#include "blackhole/core/environment.h"
#endif  // BH8_FWD_CV_KEY_H_
