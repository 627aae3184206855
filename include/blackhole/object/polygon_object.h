// Forwarding header: the reference's include path blackhole/object/polygon_object.h maps onto this repository's
// implementation in blackhole/core/.
#ifndef BH8_FWD_OBJECT_POLYGON_OBJECT_H_
#define BH8_FWD_OBJECT_POLYGON_OBJECT_H_
#endif  // BH8_FWD_OBJECT_POLYGON_OBJECT_H_
