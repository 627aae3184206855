// Forwarding header: the reference's include path blackhole/object/vector_object.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_OBJECT_VECTOR_OBJECT_H_
#define BH8_FWD_OBJECT_VECTOR_OBJECT_H_
#include "blackhole/core/shapes.h"
#endif  // BH8_FWD_OBJECT_VECTOR_OBJECT_H_
