// Forwarding header: the reference's include path blackhole/matrix.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_MATRIX_H_
#define BH8_FWD_MATRIX_H_
#include "blackhole/core/linear.h"
#endif  // BH8_FWD_MATRIX_H_
