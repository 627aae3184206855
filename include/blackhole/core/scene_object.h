// blackhole/core/scene_object.h -- Material, Object, DrawableObject.
//
// Implementation header of this repository's blackhole:: API.  The file names the reference uses
// (blackhole/camera.h, blackhole/object/vector_object.h, ...) are thin forwarding headers onto the
// blackhole/core/ set, so code written against the reference's include paths compiles unchanged.
//
// blackhole/object/material.h -- surface appearance of a drawable object.
//
// API of the reference's object/material.h:16-61: a public cv::Mat texture_ (CV_8UC3, BGR -- the
// GPU snapshot uploads exactly these bytes), SetTexture() that refuses an empty image the hard way,
// SetColor(), and a virtual color(x, y, z) the shapes override.
// blackhole/object/object.h -- positioned, oriented scene objects.
//
// Source-compatible with the reference's object/object.h:20-134.  Conventions the hot path relies on:
//   * vertex()[0] is the object's position / centre; shape constructors APPEND their corners, so a
//     Rectangle's corners are vertex()[1..4] (reference object.h:38-40,109);
//   * the basis (vector_x/y/z) starts as the identity and is re-normalised after every rotation;
//   * Rotate* turns the basis and every vertex about position() with the Rodrigues matrix of
//     matrix.h, Move* translates every vertex along a basis vector.
// Additive to the reference: ShapeKind / DrawableObject::kind(), so a scene can be snapshotted for
// the GPU without RTTI.
#ifndef BLACKHOLE_CORE_SCENE_OBJECT_H_
#define BLACKHOLE_CORE_SCENE_OBJECT_H_

#include <exception>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "opencv2/opencv.hpp"
#include "blackhole/core/linear.h"

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

namespace blackhole {

template <typename Pos, typename Vec>
struct PointVector {
  Pos origin;
  Vec vector;
};

// What kind of Collide()/color() pair an object carries (additive API).
enum class ShapeKind { kUnknown, kBlackhole, kAnnulus, kRectangle, kInfinitePlane, kTriangle, kSphere };

template <typename T>
class Object {
 public:
  using value_type = T;
  static_assert(std::is_floating_point_v<value_type>);

  using point_type = cv::Vec<value_type, 3>;
  using vector_type = cv::Vec<value_type, 3>;
  using matrix_type = cv::Matx<value_type, 3, 3>;

 private:
  template <typename X>
  using if_not_arithmetic = std::enable_if_t<!std::is_arithmetic_v<X>, int>;

 public:
  Object() = default;

  template <typename P = point_type, if_not_arithmetic<P> = 0>
  explicit Object(const P& p) : vertex_({p}) {}

  // Corners only: the position stays at the origin and the corners follow it in vertex().
  template <typename P = point_type, if_not_arithmetic<P> = 0>
  Object(std::initializer_list<P> corners) {
    vertex_.insert(vertex_.end(), corners);
  }

  template <typename Vec = vector_type, if_not_arithmetic<Vec> = 0>
  Object(const Vec& vx, const Vec& vy, const Vec& vz)
      : vx_(cv::normalize(vx)), vy_(cv::normalize(vy)), vz_(cv::normalize(vz)) {}

  template <typename P = point_type, typename Vec = vector_type, if_not_arithmetic<P> = 0, if_not_arithmetic<Vec> = 0>
  Object(const P& p, const Vec& vx, const Vec& vy, const Vec& vz)
      : vertex_({p}), vx_(cv::normalize(vx)), vy_(cv::normalize(vy)), vz_(cv::normalize(vz)) {}

  template <typename P = point_type, typename Vec = vector_type, if_not_arithmetic<P> = 0, if_not_arithmetic<Vec> = 0>
  Object(const P& p, const Vec& vx, const Vec& vy, const Vec& vz, std::initializer_list<point_type> corners)
      : vertex_({p}), vx_(cv::normalize(vx)), vy_(cv::normalize(vy)), vz_(cv::normalize(vz)) {
    vertex_.insert(vertex_.end(), corners);
  }

  void RotateX(value_type rad) { Turn(vx_, rad, &vy_, &vz_); }
  void RotateY(value_type rad) { Turn(vy_, rad, &vx_, &vz_); }
  void RotateZ(value_type rad) { Turn(vz_, rad, &vx_, &vy_); }

  void MoveTo(const point_type& p) { vertex_[0] = p; }
  void MoveTo(value_type x, value_type y, value_type z) { vertex_[0] = point_type(x, y, z); }

  void Move(value_type dx, value_type dy, value_type dz) {
    MoveX(dx);
    MoveY(dy);
    MoveZ(dz);
  }
  void MoveX(value_type distance) { Shift(vx_, distance); }
  void MoveY(value_type distance) { Shift(vy_, distance); }
  void MoveZ(value_type distance) { Shift(vz_, distance); }

  [[nodiscard]] const point_type& position() const { return vertex_[0]; }
  [[nodiscard]] const vector_type& vector_x() const { return vx_; }
  [[nodiscard]] const vector_type& vector_y() const { return vy_; }
  [[nodiscard]] const vector_type& vector_z() const { return vz_; }

  const std::vector<point_type>& vertex() const { return vertex_; }

#ifndef NDEBUG
  void name(std::string name) { name_ = std::move(name); }
  [[nodiscard]] const std::string& name() const { return name_; }
#else
  // Release builds do not store names (the reference's object.h:100-105 does the same).
  void name(const std::string&) {}
  [[nodiscard]] const std::string& name() const {
    static const auto* release_name = new std::string("Unnamed: Release build");
    return *release_name;
  }
#endif

 protected:
  std::vector<point_type> vertex_ = {point_type(0, 0, 0)};

 private:
  // Rotation by `rad` about `axis` (one of the basis vectors): the other two basis vectors are
  // rotated and re-normalised, every vertex is rotated about position().
  void Turn(const vector_type& axis, value_type rad, vector_type* a, vector_type* b) {
    const auto rot = RotationMatrixForAxis<matrix_type>(axis, rad);
    *a = cv::normalize(rot * *a);
    *b = cv::normalize(rot * *b);
    for (auto& v : vertex_) v = rot * (v - position()) + position();
  }
  void Shift(const vector_type& direction, value_type distance) {
    for (auto& v : vertex_) v += direction * distance;
  }

  vector_type vx_ = {1, 0, 0};
  vector_type vy_ = {0, 1, 0};
  vector_type vz_ = {0, 0, 1};
#ifndef NDEBUG
  std::string name_ = "Unnamed";
#endif
};

template <typename T>
class DrawableObject : public Object<T>, public Material<T> {
 public:
  using object = Object<T>;
  using value_type = typename object::value_type;
  using point_type = typename object::point_type;
  using vector_type = typename object::vector_type;
  using matrix_type = typename object::matrix_type;

  using object::object;
  using Material<T>::color;

  // Does the segment p1 -> p2 meet the object?  On a hit *intersection is the meeting point.
  virtual bool Collide(const point_type& /*p1*/, const point_type& /*p2*/, point_type* /*intersection*/) const {
    return false;
  }

  virtual ShapeKind kind() const { return ShapeKind::kUnknown; }
};

}  // namespace blackhole

#endif  // BLACKHOLE_CORE_SCENE_OBJECT_H_
