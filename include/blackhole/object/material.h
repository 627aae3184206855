// blackhole/object/material.h -- surface appearance of a drawable object.
//
// API of the reference's object/material.h:16-61: a public cv::Mat texture_ (CV_8UC3, BGR -- the
// GPU snapshot uploads exactly these bytes), SetTexture() that refuses an empty image the hard way,
// SetColor(), and a virtual color(x, y, z) the shapes override.
#ifndef BLACKHOLE_MATERIAL_H_
#define BLACKHOLE_MATERIAL_H_

#include <exception>
#include <functional>
#include <iostream>
#include <utility>

#include "opencv2/opencv.hpp"

namespace blackhole {

template <typename T>
class Material {
 public:
  using value_type = T;
  using function_type = std::function<cv::Scalar(value_type x, value_type y, value_type z)>;
  using texture_type = cv::Mat;
  using color_type = cv::Scalar;

  Material() = default;
  virtual ~Material() = default;

  // An empty image (failed imread) is a programming error in the drivers: report and terminate,
  // as the reference does (material.h:29-35).
  void SetTexture(texture_type texture) {
    if (texture.empty()) {
      std::cerr << "Empty texture!\n";
      std::terminate();
    }
    texture_ = std::move(texture);
  }

  // A uniform colour is a 2x2 texture of that colour.
  void SetColor(const cv::Vec3b& color) { texture_ = cv::Mat(2, 2, CV_8UC3, color); }

  virtual cv::Vec3b color(value_type /*x*/, value_type /*y*/, value_type /*z*/) const { return {0, 0, 0}; }

  template <typename P>
  cv::Vec3b color(const P& p) const {
    return color(p[0], p[1], p[2]);
  }

  texture_type texture_;
};

}  // namespace blackhole

#endif  // BLACKHOLE_MATERIAL_H_
