// blackhole/object/pattern.h -- procedural patterns for InfinitePlane.
//
// ChessPattern2D reproduces the reference's object/pattern.h:13-50 cell rule exactly, including its
// integer-division period (`(int)|a| / (int)(2 size)`) and the colour swap across the axes.
#ifndef BLACKHOLE_OBJECT_PATTERN_H_
#define BLACKHOLE_OBJECT_PATTERN_H_

#include <cmath>

#include "opencv2/opencv.hpp"

namespace blackhole {

template <typename T>
class ChessPattern2D {
 public:
  using value_type = T;

  ChessPattern2D() = default;
  explicit ChessPattern2D(int pattern_size) : pattern_size_(pattern_size) {}

  value_type pattern_size() const { return pattern_size_; }  // additive: lets a scene be snapshotted

  cv::Vec3b operator()(value_type x, value_type y, value_type /* z */) {
    const auto fx = Fold(x);
    const auto fy = Fold(y);
    // Both coordinates in the same half of the 2*size period?
    const bool same_half = (fx <= pattern_size_ && fy <= pattern_size_) || (pattern_size_ <= fx && pattern_size_ <= fy);
    const bool white = (x * y > 0) ? same_half : !same_half;
    return white ? cv::Vec3b{255, 255, 255} : cv::Vec3b{0, 0, 0};
  }

 private:
  // |a| reduced by whole periods of 2*size, the period count taken with integer division.
  value_type Fold(value_type a) {
    const auto magnitude = std::abs(a);
    return magnitude - ((int)(magnitude) / (int)(pattern_size_ * 2)) * (pattern_size_ * 2);
  }

  value_type pattern_size_ = 1;
};

}  // namespace blackhole

#endif  // BLACKHOLE_OBJECT_PATTERN_H_
