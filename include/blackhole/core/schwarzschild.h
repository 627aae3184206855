// blackhole/core/schwarzschild.h -- StaticBlackhole.
//
// Implementation header of this repository's blackhole:: API.  The file names the reference uses
// (blackhole/camera.h, blackhole/object/vector_object.h, ...) are thin forwarding headers onto the
// blackhole/core/ set, so code written against the reference's include paths compiles unchanged.
//
// blackhole/blackhole_solution.h -- the Schwarzschild black hole the geodesics bend around.
//
// Source-compatible with the reference's blackhole_solution.h:16-96.  In units G = c = 1 a photon
// with impact parameter b obeys (du/dphi)^2 = G(u, b) = 2M u^3 - u^2 + 1/b^2, u = 1/r.
//   G(), InvSqrtG()  the orbit function and dphi/du = 1/sqrt(G)
//   SolveG()         the turning point for b >= b_c = 3 sqrt(3) M: exactly 20 bisections on
//                    [cbrt(eps), 1/(3M)], returning the LEFT end (so G(result) > 0)
//   Collide()        the horizon as a sphere of radius 2M: entry point of the segment, 0 < t < 1
// The evaluation order of every expression follows the reference so results are bit-identical.
#ifndef BLACKHOLE_CORE_SCHWARZSCHILD_H_
#define BLACKHOLE_CORE_SCHWARZSCHILD_H_

#include <cmath>

#include "blackhole/core/numeric.h"
#include "blackhole/core/linear.h"
#include "blackhole/core/scene_object.h"
#include "blackhole/core/shapes.h"
#include "blackhole/core/scene.h"

namespace blackhole {

template <typename T>
class StaticBlackhole : public DrawableObject<T> {
 public:
  using object = DrawableObject<T>;
  using value_type = T;
  using point_type = typename object::point_type;
  using vector_type = typename object::vector_type;

  StaticBlackhole(const vector_type& position, value_type mass)
      : object(position), mass_(mass), b_c_(3.0 * std::sqrt(3) * mass) {}

  value_type mass() const { return mass_; }
  value_type b_c() const { return b_c_; }              // critical impact parameter 3 sqrt(3) M
  value_type radius() const { return 2 * mass(); }     // Schwarzschild radius
  const point_type& center() const { return this->position(); }

  ShapeKind kind() const override { return ShapeKind::kBlackhole; }

  value_type G(value_type u, value_type b) const { return (u * u * (2.0 * mass_ * u - 1.0)) + (1.0 / (b * b)); }

  value_type InvSqrtG(value_type u, value_type b) const { return 1.0 / std::sqrt(G(u, b)); }

  value_type SolveG(value_type b) const {
    static const int kIterations = 20;
    value_type lo = 0.0 + epsilon<value_type>();
    value_type hi = 1.0 / (3.0 * mass());
    for (int i = 0; i < kIterations; ++i) {
      const value_type mid = (lo + hi) / 2.0;
      if (G(mid, b) > 0.0)
        lo = mid;
      else
        hi = mid;
    }
    return lo;
  }

  cv::Vec3b color(value_type, value_type, value_type) const override { return {0, 0, 0}; }

  bool Collide(const point_type& q1, const point_type& q2, point_type* intersection) const override {
    const auto from = q1 - center();
    const auto to = q2 - center();
    const auto r = radius();
    if (std::sqrt(from.dot(from)) < r && std::sqrt(to.dot(to)) < r) return false;  // wholly inside

    const auto dir = to - from;
    const auto disc = std::pow(dir.dot(from), 2) - dir.dot(dir) * (from.dot(from) - r * r);
    if (disc <= 0) return false;
    const value_type t = (((value_type)1) / dir.dot(dir)) * (-dir.dot(from) - std::sqrt(disc));
    if (t <= 0 || t >= 1) return false;
    *intersection = from + dir * t + center();
    return true;
  }

 private:
  value_type mass_;
  value_type b_c_;
};

template <typename T>
using SchwarzschildBlackhole = StaticBlackhole<T>;

}  // namespace blackhole

#endif  // BLACKHOLE_CORE_SCHWARZSCHILD_H_
