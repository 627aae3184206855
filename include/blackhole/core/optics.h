// blackhole/core/optics.h -- Camera and the flat-space RayTracer.
//
// Implementation header of this repository's blackhole:: API.  The file names the reference uses
// (blackhole/camera.h, blackhole/object/vector_object.h, ...) are thin forwarding headers onto the
// blackhole/core/ set, so code written against the reference's include paths compiles unchanged.
//
// blackhole/camera.h -- pin-hole camera: an Object (position + basis) plus the pixel -> direction map.
//
// Source-compatible with the reference's camera.h:21-71.  The camera looks along vector_x(); image
// x grows along -vector_y(), image y along -vector_z() (default basis (1,0,0), (0,-1,0), (0,0,-1)).
// PixelVector() is NOT normalised: its length is of the order of focus_len = width / (2 tan(fov/2)),
// and the geodesic drivers depend on that (their impact parameter involves |F - d|).
// Additive: focus_len().
// blackhole/ray_tracer.h -- point-pair ray marching for flat space.
//
// Source-compatible with the reference's ray_tracer.h:17-99: RayTracer(old, present) keeps the last
// two points of a ray; Prograde() tests the current segment against the scene and, on a miss,
// extends the ray with a recurrence (default: a constant step vector of length >= 10).
#ifndef BLACKHOLE_CORE_OPTICS_H_
#define BLACKHOLE_CORE_OPTICS_H_

#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <type_traits>
#include <utility>

#include "opencv2/opencv.hpp"
#include "blackhole/core/numeric.h"
#include "blackhole/core/linear.h"
#include "blackhole/core/scene_object.h"
#include "blackhole/core/shapes.h"
#include "blackhole/core/scene.h"

namespace blackhole {

template <typename T>
class Camera : public Object<T> {
 public:
  using base = Object<T>;
  using value_type = typename base::value_type;
  using point_type = typename base::point_type;
  using vector_type = typename base::vector_type;
  using matrix_type = typename base::matrix_type;

  Camera(int width, int height, value_type fov_x)
      : base({0, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 0, -1}), width_(width), height_(height) {
    fov(fov_x);
  }

  // Horizontal field of view, clamped to [0, pi).
  void fov(value_type fov_x) {
    const value_type below_pi = std::nextafter(blackhole::kPi<value_type>, (value_type)0);
    fov_ = std::min(std::max((value_type)0, fov_x), below_pi);
    focus_len_ = width() / (2 * std::tan(fov_ / 2.0));
  }
  [[nodiscard]] value_type fov() const { return fov_; }
  [[nodiscard]] value_type focus_len() const { return focus_len_; }  // additive

  [[nodiscard]] int width() const { return width_; }
  [[nodiscard]] int height() const { return height_; }

  [[nodiscard]] const vector_type& focus() const { return this->position(); }

  [[nodiscard]] vector_type focus_vector() const { return this->vector_x() * focus_len_; }

  [[nodiscard]] vector_type PixelVector(int x, int y) const { return PixelVector(x, y, focus_vector()); }

  [[nodiscard]] vector_type PixelVector(int x, int y, const vector_type& focus_v) const {
    const auto right = static_cast<value_type>(width() / 2.0 - x);
    const auto down = static_cast<value_type>(height() / 2.0 - y);
    return focus_v - this->vector_y() * right - this->vector_z() * down;
  }

 private:
  int width_, height_;
  value_type focus_len_ = 1;
  value_type fov_ = pi / 2;
};

}  // namespace blackhole

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

#endif  // BLACKHOLE_CORE_OPTICS_H_
