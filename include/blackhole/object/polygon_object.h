// blackhole/object/polygon_object.h -- placeholder, empty in the reference as well.
#ifndef BLACKHOLE_POLYGON_OBJECT_H_
#define BLACKHOLE_POLYGON_OBJECT_H_
#endif  // BLACKHOLE_POLYGON_OBJECT_H_
