// blackhole/ray_tracer.h -- point-pair ray marching for flat space.
//
// Source-compatible with the reference's ray_tracer.h:17-99: RayTracer(old, present) keeps the last
// two points of a ray; Prograde() tests the current segment against the scene and, on a miss,
// extends the ray with a recurrence (default: a constant step vector of length >= 10).
#ifndef BLACKHOLE_RAY_TRACER_H_
#define BLACKHOLE_RAY_TRACER_H_

#include <cmath>
#include <functional>
#include <utility>

#include "blackhole/object.h"

#include "opencv2/opencv.hpp"

namespace blackhole {

// present + (present0 - old0), the initial step stretched to squared length 100 when shorter
// (by the factor 100/|v|^2, as the reference does).
template <typename Point>
class BasicLinearRayRecurrence {
 public:
  using point_type = Point;

  BasicLinearRayRecurrence(const point_type& old, const point_type& present) : step_(present - old) {
    const auto len2 = step_.dot(step_);
    if (len2 < 100) step_ *= 100.0 / len2;
  }

  point_type operator()(const point_type& /* old */, const point_type& present) const { return present + step_; }

 private:
  point_type step_;
};

// Placeholder in the reference as well: it only stores its arguments.
template <typename Point>
class FixedSingleBlackholeRayRecurrence {
 public:
  using point_type = Point;
  FixedSingleBlackholeRayRecurrence(const point_type& bh_pos, double mass, const point_type& old,
                                    const point_type& present)
      : points_(old, present), blackhole_(bh_pos), blackhole_mass_(mass) {}

 private:
  std::pair<point_type, point_type> points_;
  point_type blackhole_;
  double blackhole_mass_;
};

template <typename Point>
class RayTracer {
 public:
  using point_type = Point;
  using value_type = typename point_type::value_type;
  using recurrence_type = std::function<point_type(const point_type& old, const point_type& present)>;

  RayTracer(const point_type& old, const point_type& present)
      : point_(old, present), recurrence_(BasicLinearRayRecurrence<point_type>{old, present}) {}

  RayTracer(const point_type& old, const point_type& present, recurrence_type recurrence)
      : point_(old, present), recurrence_(std::move(recurrence)) {}

  // Up to step_size segments; writes the hit object's colour (3 bytes, BGR) and returns true on a hit.
  template <typename ObjManager>
  bool Prograde(ObjManager& obj_manager, unsigned char* color_dst, int step_size) {
    point_type hit_point;
    for (int s = 0; s < step_size; ++s) {
      if (const auto* obj = obj_manager.FindCollision(old(), present(), &hit_point); obj != nullptr) {
        const auto c = obj->color(hit_point);
        color_dst[0] = c[0];
        color_dst[1] = c[1];
        color_dst[2] = c[2];
        return true;
      }
      auto next = recurrence_(old(), present());
      point_.first = present();
      point_.second = std::move(next);
    }
    return false;
  }

  const point_type& old() const { return point_.first; }
  const point_type& present() const { return point_.second; }

 private:
  std::pair<point_type, point_type> point_;  // old, present
  recurrence_type recurrence_;
};

template <typename Point>
RayTracer(const Point&, ...) -> RayTracer<Point>;

}  // namespace blackhole

#endif  // BLACKHOLE_RAY_TRACER_H_
