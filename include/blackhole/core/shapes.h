// blackhole/core/shapes.h -- procedural patterns and the drawable shapes.
//
// Implementation header of this repository's blackhole:: API.  The file names the reference uses
// (blackhole/camera.h, blackhole/object/vector_object.h, ...) are thin forwarding headers onto the
// blackhole/core/ set, so code written against the reference's include paths compiles unchanged.
//
// blackhole/object/pattern.h -- procedural patterns for InfinitePlane.
//
// ChessPattern2D reproduces the reference's object/pattern.h:13-50 cell rule exactly, including its
// integer-division period (`(int)|a| / (int)(2 size)`) and the colour swap across the axes.
// blackhole/object/vector_object.h -- the drawable shapes: Triangle, Rectangle, InfinitePlane,
// Sphere, Annulus (+ the Cylinder placeholders).
//
// Source-compatible with the reference's object/vector_object.h; the three shapes on the geodesic
// hot path keep its exact arithmetic, so a CPU render through these headers is bit-identical:
//   Rectangle::Collide  reference :107-127   strict inside test, plane touch = miss
//   Rectangle::color    reference :159-179   nearest texel by truncation, no clamp, BGR
//   InfinitePlane       reference :210-232   plane touch = hit; colour from absolute x, y
//   Annulus::Collide    reference :328-347   normal fixed at construction, radii tested on the centre distance
// All three planar shapes find the meeting point of a segment with their plane by the same linear
// blend of the end points weighted with the opposite end's distance -- detail::BlendByDistance.
// Differences in code no driver instantiates: Sphere::Collide here is a working segment/sphere
// entry test (the reference's version falls off the end of the function, :281-300).
// Additive API: kind(), Annulus::r_inner()/r_outer(), InfinitePlane::pattern_kind()/pattern_size().
#ifndef BLACKHOLE_CORE_SHAPES_H_
#define BLACKHOLE_CORE_SHAPES_H_

#include <cmath>
#include <functional>
#include <limits>

#include "opencv2/opencv.hpp"
#include "blackhole/core/numeric.h"
#include "blackhole/core/linear.h"
#include "blackhole/core/scene_object.h"

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

namespace blackhole {

namespace detail {

// Point of the segment a -> b at signed plane distances da, db (opposite signs): each end weighted
// by the other end's distance.
template <typename V, typename T>
V BlendByDistance(const V& a, T da, const V& b, T db) {
  return (std::abs(db) * a + std::abs(da) * b) / (std::abs(da) + std::abs(db));
}

// Nearest texel of `texture` for the point at offset v from the texture origin, with s1 / s2 the
// edges that span the image's width / height.  The offset is taken in polar form about the origin:
// the column from its component along s1, the row from the MAGNITUDE of the remainder (sin of an
// acos is never negative).  Truncation, no filtering, no bounds check.
template <typename V>
cv::Vec3b NearestTexel(const cv::Mat& texture, const V& s1, const V& s2, const V& v) {
  const auto r = std::sqrt(v.dot(v));
  auto theta = std::acos(s1.dot(v) / (std::sqrt(s1.dot(s1)) * r));
  theta = std::isnan(theta) ? 0 : theta;
  const int column = (int)((r * std::cos(theta) / std::sqrt(s1.dot(s1))) * texture.cols);
  const int row = (int)((r * std::sin(theta) / std::sqrt(s2.dot(s2))) * texture.rows);
  const unsigned char* texel = texture.data + ((row * texture.cols + column) * 3);
  return {texel[0], texel[1], texel[2]};
}

}  // namespace detail

// Triangle with vertices vertex()[1..3]; Moller-Trumbore segment test (t > 0 along p1 -> p2).
template <typename T>
class Triangle : public DrawableObject<T> {
 public:
  using object = Object<T>;
  using value_type = typename object::value_type;
  using point_type = typename object::point_type;
  using vector_type = typename object::vector_type;
  using matrix_type = typename object::matrix_type;

  Triangle() : DrawableObject<T>({{1, 0, 0}, {-1, 0, 0}, {0, 0, 1.73}}) {}

  Triangle(value_type v11, value_type v12, value_type v13,  //
           value_type v21, value_type v22, value_type v23,  //
           value_type v31, value_type v32, value_type v33)
      : DrawableObject<T>({{v11, v12, v13}, {v21, v22, v23}, {v31, v32, v33}}) {}

  ShapeKind kind() const override { return ShapeKind::kTriangle; }

  bool Collide(const point_type& p1, const point_type& p2, point_type* intersection) const override {
    const value_type tiny = std::numeric_limits<value_type>::epsilon() * 10;
    const auto& a = this->vertex()[1];
    const auto ab = this->vertex()[2] - a;
    const auto ac = this->vertex()[3] - a;
    const auto dir = p2 - p1;
    const auto h = dir.cross(ac);
    const auto det = ab.dot(h);
    if (-tiny < det && det < tiny) return false;  // segment parallel to the triangle
    const auto inv = static_cast<value_type>(1.0) / det;
    const auto s = p1 - a;
    const auto u = inv * s.dot(h);
    if (u < 0 || u > 1) return false;
    const auto q = s.cross(ab);
    const auto v = inv * dir.dot(q);
    if (v < 0 || u + v > 1) return false;
    const auto t = inv * ac.dot(q);
    if (t <= tiny) return false;
    *intersection = p1 + dir * t;
    return true;
  }

  static double SignedTetraVolume(const point_type& a, const point_type& b, const point_type& c,
                                  const point_type& d) {
    return (b - a).cross(c - a).dot(d - a);
  }
};

// Textured parallelogram with corners vertex()[1..4]; [1] is the texture origin, [2]-[1] spans the
// image width and [4]-[1] its height.
template <typename T>
class Rectangle : public DrawableObject<T> {
 public:
  using object = DrawableObject<T>;
  using value_type = typename object::value_type;
  using point_type = typename object::point_type;
  using vector_type = typename object::vector_type;
  using matrix_type = typename object::matrix_type;

  Rectangle() : object({{1, 0, 1}, {-1, 0, 1}, {-1, 0, -1}, {1, 0, -1}}) {}

  Rectangle(const point_type& p1, const point_type& p2, const point_type& p3, const point_type& p4)
      : object({p1, p2, p3, p4}) {}

  Rectangle(value_type v11, value_type v12, value_type v13,  //
            value_type v21, value_type v22, value_type v23,  //
            value_type v31, value_type v32, value_type v33,  //
            value_type v41, value_type v42, value_type v43)
      : object({{v11, v12, v13}, {v21, v22, v23}, {v31, v32, v33}, {v41, v42, v43}}) {}

  ShapeKind kind() const override { return ShapeKind::kRectangle; }

  bool Collide(const point_type& aa, const point_type& bb, point_type* intersection) const override {
    const auto origin = this->vertex()[1];
    const auto edge_w = this->vertex()[2] - origin;
    const auto edge_h = this->vertex()[4] - origin;
    const auto normal = cv::normalize(edge_w.cross(edge_h));
    const auto rel_a = aa - origin;
    const auto rel_b = bb - origin;
    const auto side_a = normal.dot(rel_a);
    const auto side_b = normal.dot(rel_b);
    if (side_a * side_b >= 0) return false;  // same side, or an end point exactly in the plane

    const auto in_plane = detail::BlendByDistance(rel_a, side_a, rel_b, side_b);
    const bool inside = edge_w.dot(in_plane) > 0 && edge_w.dot(edge_w) > edge_w.dot(in_plane) &&
                        edge_h.dot(in_plane) > 0 && edge_h.dot(edge_h) > edge_h.dot(in_plane);
    if (!inside) return false;
    *intersection = in_plane + origin;
    return true;
  }

  cv::Vec3b color(value_type x, value_type y, value_type z) const override { return color({x, y, z}); }

  cv::Vec3b color(const point_type& p) const {
    const auto& origin = this->vertex()[1];
    return detail::NearestTexel(this->texture_, this->vertex()[2] - origin, this->vertex()[4] - origin,
                                p - origin);
  }

  void foo() {}
};

// Plane through position() with normal vector_z(); coloured by a pattern evaluated at the plane
// coordinates (a, b) solved from the hit point's absolute x, y.
template <typename T>
class InfinitePlane : public DrawableObject<T> {
 public:
  using object = DrawableObject<T>;
  using value_type = typename object::value_type;
  using point_type = typename object::point_type;
  using vector_type = typename object::vector_type;
  using matrix_type = typename object::matrix_type;
  using pattern_type = std::function<cv::Vec3b(value_type x, value_type y, value_type z)>;

  enum class PatternKind { kBlack, kChess, kOpaque };

  using object::position;
  using object::vector_x;
  using object::vector_y;
  using object::vector_z;

  InfinitePlane(const point_type& p) : object(p) {}

  InfinitePlane(const point_type& p, pattern_type pattern)
      : object(p), pattern_(std::move(pattern)), pattern_kind_(PatternKind::kOpaque) {}

  // Additive overload: a chess pattern is remembered as such so the scene can be sent to the GPU.
  InfinitePlane(const point_type& p, ChessPattern2D<value_type> chess)
      : object(p), pattern_(chess), pattern_kind_(PatternKind::kChess), pattern_size_(chess.pattern_size()) {}

  InfinitePlane(const point_type& p, const cv::Vec3b& color)
      : object(p), pattern_([color](auto...) { return color; }), pattern_kind_(PatternKind::kOpaque) {}

  ShapeKind kind() const override { return ShapeKind::kInfinitePlane; }
  PatternKind pattern_kind() const { return pattern_kind_; }
  value_type pattern_size() const { return pattern_size_; }

  bool Collide(const point_type& p1, const point_type& p2, point_type* intersection) const override {
    const auto t1 = Height(p1);
    const auto t2 = Height(p2);
    if (t1 * t2 > 0) return false;  // strictly the same side; touching the plane counts as a hit
    *intersection = detail::BlendByDistance(p1, t1, p2, t2);
    return true;
  }

  [[nodiscard]] cv::Vec3b color(value_type x, value_type y, value_type /*z*/) const override {
    const auto b = (vector_x()[0] * y - vector_x()[1] * x) /
                   (vector_x()[0] * vector_y()[1] - vector_x()[1] * vector_y()[0]);
    const auto a = (x - b * vector_y()[0]) / vector_x()[0];
    return pattern_(a, b, 0);
  }

 private:
  using object::object;
  using Material<value_type>::SetTexture;

  // Signed distance of q from the plane (unit normal vector_z()).
  value_type Height(const point_type& q) const {
    return vector_z()[0] * (q[0] - position()[0]) + vector_z()[1] * (q[1] - position()[1]) +
           vector_z()[2] * (q[2] - position()[2]);
  }

  pattern_type pattern_ = [](auto...) -> cv::Vec3b { return {0, 0, 0}; };
  PatternKind pattern_kind_ = PatternKind::kBlack;
  value_type pattern_size_ = 0;
};

template <typename T>
class Cylinder : public DrawableObject<T> {};

template <typename T>
class InfiniteCylinder : public DrawableObject<T> {};

template <typename T>
class Sphere : public DrawableObject<T> {
 public:
  using object = DrawableObject<T>;
  using value_type = typename object::value_type;
  using point_type = typename object::point_type;
  using vector_type = typename object::vector_type;
  using matrix_type = typename object::matrix_type;

  using object::position;

  Sphere(const point_type& position, value_type radius) : object(position), radius_(radius) {}

  void radius(value_type r) { radius_ = r; }
  value_type radius() const { return radius_; }
  const point_type& center() const { return position(); }

  ShapeKind kind() const override { return ShapeKind::kSphere; }

  cv::Vec3b color(value_type x, value_type y, value_type z) const override {
    return Material<value_type>::color(x, y, z);
  }

  // First point where the segment enters the sphere, if any.
  bool Collide(const point_type& p1, const point_type& p2, point_type* intersection) const override {
    const auto from = p1 - center();
    const auto dir = p2 - p1;
    const auto dd = dir.dot(dir);
    if (dd == 0) return false;
    const auto half_b = dir.dot(from);
    const auto disc = half_b * half_b - dd * (from.dot(from) - radius_ * radius_);
    if (disc <= 0) return false;
    const auto t = (-half_b - std::sqrt(disc)) / dd;
    if (t <= 0 || t >= 1) return false;
    *intersection = p1 + dir * t;
    return true;
  }

 private:
  value_type radius_ = 10;
};

// Thin accretion disc: the part of a plane between r_inner and r_outer around center()
// (= vertex()[0]); it is-a Rectangle for its texture mapping.  The normal is computed once from the
// constructor's corners and is NOT updated by Rotate* (the vertices and the centre are).
template <typename T>
class Annulus : public Rectangle<T> {
 public:
  using object = Rectangle<T>;
  using value_type = typename object::value_type;
  using point_type = typename object::point_type;
  using vector_type = typename object::vector_type;
  using matrix_type = typename object::matrix_type;

  using object::position;
  using object::vector_x;
  using object::vector_y;
  using object::vector_z;

  Annulus(const point_type& p1, const point_type& p2, const point_type& p3, const point_type& p4,
          value_type r_outer, value_type r_inner)
      : object(p1, p2, p3, p4),
        norm_(cv::normalize((p3 - p1).cross(p2 - p1))),
        r_outer_(r_outer),
        r_inner_(r_inner) {}

  ShapeKind kind() const override { return ShapeKind::kAnnulus; }

  bool Collide(const point_type& q1, const point_type& q2, point_type* intersection) const override {
    const auto rel_1 = q1 - center();
    const auto rel_2 = q2 - center();
    if (norm().dot(rel_1) * norm().dot(rel_2) >= 0) return false;  // no strict plane crossing

    const auto height_1 = norm().dot(rel_1);
    const auto height_2 = norm().dot(rel_2);
    const auto in_plane = detail::BlendByDistance(rel_1, height_1, rel_2, height_2);
    if (std::sqrt(in_plane.dot(in_plane)) > r_outer_) return false;
    if (std::sqrt(in_plane.dot(in_plane)) < r_inner_) return false;
    *intersection = in_plane + center();
    return true;
  }

  const point_type& center() const { return this->vertex()[0]; }
  const vector_type& norm() const { return norm_; }
  value_type r_outer() const { return r_outer_; }  // additive
  value_type r_inner() const { return r_inner_; }  // additive

 private:
  vector_type norm_;
  value_type r_outer_;
  value_type r_inner_;
};

}  // namespace blackhole

#endif  // BLACKHOLE_CORE_SHAPES_H_
