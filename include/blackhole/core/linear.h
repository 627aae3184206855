// blackhole/core/linear.h -- small fixed-size linear algebra, axis-angle rotation, vector helpers, step generators.
//
// Implementation header of this repository's blackhole:: API.  The file names the reference uses
// (blackhole/camera.h, blackhole/object/vector_object.h, ...) are thin forwarding headers onto the
// blackhole/core/ set, so code written against the reference's include paths compiles unchanged.
//
// blackhole/matrix.h -- small fixed-size Matrix / Vector templates and the axis-angle rotation used
// by Object::Rotate*.
//
// API-compatible with the part of the reference's matrix.h that anything uses (matrix_test.cc:16-23
// and object/object.h:58-80): aggregate-initialisable Matrix<T,m,n>, Vector<T,m>, Point<T>, the
// element-wise operators, dot(), and RotationMatrixForAxis<M>(axis, theta) (matrix.h:342-356), whose
// entries are evaluated in the reference's association so rotated scenes are bit-identical.
// Deliberate differences in code no render path touches: normalize(Vector) divides by the length
// (the reference divides by the squared length, matrix.h:337-340), and the reference's det() /
// resize() stubs (sqrt of the element sum; missing return) are not reproduced.
// blackhole/math.h -- small vector helpers and step generators (API of the reference's math.h:15-100).
// size/abs/angle/divide behave as in the reference; the step generators, which the reference only
// declares as empty shells, are given the obvious working bodies.
#ifndef BLACKHOLE_CORE_LINEAR_H_
#define BLACKHOLE_CORE_LINEAR_H_

#include <cmath>
#include <cstddef>
#include <type_traits>
#include <utility>
#include <vector>

#include "opencv2/opencv.hpp"

namespace blackhole {

// Plain element storage; keeping it a public aggregate base lets `Matrix<int,2,2> m = {1,2,3,4};`
// brace-initialise straight into the array.
template <typename T, size_t N>
struct MatrixElement {
  T elem[N];
};

template <typename T, size_t m, size_t n>
class Matrix : public MatrixElement<T, m * n> {
 public:
  enum { rows = m, cols = n, size = m * n };
  using base = MatrixElement<T, m * n>;
  using base::elem;
  using value_type = T;

  constexpr value_type& operator[](size_t i) { return elem[i]; }
  constexpr const value_type& operator[](size_t i) const { return elem[i]; }
  constexpr value_type& operator()(size_t i, size_t j) { return elem[i * cols + j]; }
  constexpr const value_type& operator()(size_t i, size_t j) const { return elem[i * cols + j]; }

  double dot(const Matrix& rhs) const {
    value_type acc = 0;
    for (size_t i = 0; i < size; ++i) acc += elem[i] * rhs[i];
    return acc;
  }
};

template <typename T, size_t m>
class Vector : public Matrix<T, m, 1> {
 public:
  using base = Matrix<T, m, 1>;
};

template <typename T>
using Point = Vector<T, 3>;

namespace detail {
template <typename R, typename A, typename F>
R MapElements(const A& a, F f) {
  R out{};
  for (size_t i = 0; i < static_cast<size_t>(A::size); ++i) out[i] = f(a[i], i);
  return out;
}
}  // namespace detail

// ---- Matrix (op) scalar, Matrix * Matrix ----------------------------------------------------
template <typename T, size_t m, size_t n, typename U>
Matrix<T, m, n> operator*(const Matrix<T, m, n>& lhs, const U& rhs) {
  return detail::MapElements<Matrix<T, m, n>>(lhs, [&](const T& a, size_t) { return a * rhs; });
}
template <typename T, size_t m, size_t n, typename U>
Matrix<T, m, n>& operator*=(Matrix<T, m, n>& lhs, const U& rhs) {
  for (size_t i = 0; i < m * n; ++i) lhs[i] *= rhs;
  return lhs;
}
template <typename T, size_t m, size_t n, size_t l>
Matrix<T, m, n> operator*(const Matrix<T, m, l>& lhs, const Matrix<T, l, n>& rhs) {
  Matrix<T, m, n> out{};
  for (size_t i = 0; i < m; ++i)
    for (size_t j = 0; j < n; ++j)
      for (size_t k = 0; k < l; ++k) out(i, j) += lhs(i, k) * rhs(k, j);
  return out;
}

// ---- Vector (op) scalar ------------------------------------------------------------------------
#define BH_VECTOR_SCALAR_OP(op)                                                             \
  template <typename T, size_t m, typename U>                                               \
  Vector<T, m> operator op(const Vector<T, m>& v, const U& x) {                             \
    return detail::MapElements<Vector<T, m>>(v, [&](const T& a, size_t) { return a op x; }); \
  }                                                                                         \
  template <typename T, size_t m, typename U>                                               \
  Vector<T, m>& operator op##=(Vector<T, m>& v, const U& x) {                               \
    for (size_t i = 0; i < m; ++i) v[i] op## = x;                                           \
    return v;                                                                               \
  }
BH_VECTOR_SCALAR_OP(*)
BH_VECTOR_SCALAR_OP(/)
BH_VECTOR_SCALAR_OP(+)
BH_VECTOR_SCALAR_OP(-)
#undef BH_VECTOR_SCALAR_OP

template <typename T, size_t m>
Vector<T, m> operator*(const T& x, const Vector<T, m>& v) {
  return v * x;
}
template <typename T, size_t m>
Vector<T, m> operator-(const Vector<T, m>& v) {
  return detail::MapElements<Vector<T, m>>(v, [](const T& a, size_t) { return -a; });
}

// ---- Vector (op) Vector ------------------------------------------------------------------------
template <typename T, size_t m>
Vector<T, m> operator+(const Vector<T, m>& lhs, const Vector<T, m>& rhs) {
  return detail::MapElements<Vector<T, m>>(lhs, [&](const T& a, size_t i) { return a + rhs[i]; });
}
template <typename T, size_t m>
Vector<T, m> operator-(const Vector<T, m>& lhs, const Vector<T, m>& rhs) {
  return detail::MapElements<Vector<T, m>>(lhs, [&](const T& a, size_t i) { return a - rhs[i]; });
}
template <typename T, size_t m>
Vector<T, m>& operator+=(Vector<T, m>& lhs, const Vector<T, m>& rhs) {
  for (size_t i = 0; i < m; ++i) lhs[i] += rhs[i];
  return lhs;
}
template <typename T, size_t m>
Vector<T, m>& operator-=(Vector<T, m>& lhs, const Vector<T, m>& rhs) {
  for (size_t i = 0; i < m; ++i) lhs[i] -= rhs[i];
  return lhs;
}

template <typename T, size_t m>
Vector<T, m> normalize(const Vector<T, m>& v) {
  const double len = std::sqrt(v.dot(v));
  return len > 0 ? v * (1.0 / len) : v;
}

// Rodrigues rotation about `axis` by `theta` radians, returned as a 3x3 Matrix type constructible
// from nine scalars in row-major order (cv::Matx33d in this code base).  With k(a,b) = a*b*(1-cos):
//   R = cos*I + sin*[u]_x + (1-cos)*u u^T
template <typename Matrix3, typename Vec>
static Matrix3 RotationMatrixForAxis(const Vec& axis, double theta) {
  const auto u = cv::normalize(axis);
  const auto x = u[0], y = u[1], z = u[2];
  const auto c = std::cos(theta);
  const auto c2 = 1 - c;
  const auto s = std::sin(theta);
  const auto k = [c2](auto a, auto b) { return a * b * c2; };
  return Matrix3(c + k(x, x), k(x, y) - z * s, k(x, z) + y * s,  //
                 k(y, x) + z * s, c + k(y, y), k(y, z) - x * s,  //
                 k(z, x) - y * s, k(z, y) + x * s, c + k(z, z));
}

}  // namespace blackhole

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

#endif  // BLACKHOLE_CORE_LINEAR_H_
