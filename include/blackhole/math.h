// blackhole/math.h -- small vector helpers and step generators (API of the reference's math.h:15-100).
// size/abs/angle/divide behave as in the reference; the step generators, which the reference only
// declares as empty shells, are given the obvious working bodies.
#ifndef BLACKHOLE_MATH_H_
#define BLACKHOLE_MATH_H_

#include <cmath>
#include <cstddef>
#include <type_traits>

#include "opencv2/opencv.hpp"

namespace blackhole {
namespace math {

// Euclidean length of a cv::Vec.
template <typename T, int n>
auto size(const cv::Vec<T, n>& v) {
  return std::sqrt(v.dot(v));
}

template <typename T, int n>
T abs(const cv::Vec<T, n>& v) {
  return static_cast<T>(size(v));
}

template <typename T, std::enable_if_t<std::is_arithmetic_v<T>, int> = 0>
T abs(T x) {
  return std::abs(x);
}

// Angle between two vectors, radians in [0, pi].
template <typename V>
auto angle(const V& v1, const V& v2) {
  return std::acos(v1.dot(v2) / (size(v1) * size(v2)));
}

// Point dividing the segment from -> to at `ratio` (0 = from, 1 = to).
template <typename P, typename U, std::enable_if_t<std::is_floating_point_v<U>, int> = 0>
P divide(const P& from, const P& to, U ratio) {
  return from + (to - from) * ratio;
}

// first, first + h, ... with h = (last - first) / step_count.
template <typename T>
class FixedStepGenerator {
 public:
  using value_type = T;
  FixedStepGenerator(value_type first, value_type last, size_t step_count)
      : first_(first), last_(last), step_count_(step_count) {}

  value_type step() const { return step_count_ ? (last_ - first_) / static_cast<value_type>(step_count_) : 0; }
  value_type operator()(value_type prev) const { return prev + step(); }

 private:
  value_type first_;
  value_type last_;
  size_t step_count_;
};

template <typename T>
class DynamicStepGenerator {
 public:
  using value_type = T;
};

template <typename AreaCalculator, typename StepGenerator>
class Integraph {
 public:
  Integraph() = default;
  Integraph(AreaCalculator area, StepGenerator step)
      : area_calculator_(std::move(area)), step_generator_(std::move(step)) {}

 private:
  AreaCalculator area_calculator_;
  StepGenerator step_generator_;
};

}  // namespace math
}  // namespace blackhole

#endif  // BLACKHOLE_MATH_H_
