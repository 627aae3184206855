// blackhole/object.h -- umbrella include for the scene-object headers (reference: object.h:8-13).
#ifndef BLACKHOLE_OBJECT_H_
#define BLACKHOLE_OBJECT_H_

#include "blackhole/object/material.h"
#include "blackhole/object/object.h"
#include "blackhole/object/object_manager.h"
#include "blackhole/object/pattern.h"
#include "blackhole/object/polygon_object.h"
#include "blackhole/object/vector_object.h"

#endif  // BLACKHOLE_OBJECT_H_
