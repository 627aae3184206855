// Forwarding header: the reference's include path blackhole/config.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_CONFIG_H_
#define BH8_FWD_CONFIG_H_
#include "blackhole/core/environment.h"
#endif  // BH8_FWD_CONFIG_H_
