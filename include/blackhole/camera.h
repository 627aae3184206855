// blackhole/camera.h -- pin-hole camera: an Object (position + basis) plus the pixel -> direction map.
//
// Source-compatible with the reference's camera.h:21-71.  The camera looks along vector_x(); image
// x grows along -vector_y(), image y along -vector_z() (default basis (1,0,0), (0,-1,0), (0,0,-1)).
// PixelVector() is NOT normalised: its length is of the order of focus_len = width / (2 tan(fov/2)),
// and the geodesic drivers depend on that (their impact parameter involves |F - d|).
// Additive: focus_len().
#ifndef BLACKHOLE_CAMERA_H_
#define BLACKHOLE_CAMERA_H_

#include <algorithm>
#include <cmath>
#include <limits>
#include <type_traits>

#include "opencv2/opencv.hpp"

#include "blackhole/constants.h"
#include "blackhole/matrix.h"
#include "blackhole/object.h"

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

#endif  // BLACKHOLE_CAMERA_H_
