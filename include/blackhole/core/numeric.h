// blackhole/core/numeric.h -- numeric constants and tolerance helpers.
//
// Implementation header of this repository's blackhole:: API.  The file names the reference uses
// (blackhole/camera.h, blackhole/object/vector_object.h, ...) are thin forwarding headers onto the
// blackhole/core/ set, so code written against the reference's include paths compiles unchanged.
//
// blackhole/constants.h -- numeric constants of the blackhole:: API.
// Source-compatible with lackhole/blackhole_8 include/blackhole/constants.h:10-17
// (kPi<T>, kE<T>, pi, e with the same 31-digit literals, so `blackhole::pi / 2` is bit-identical).
// blackhole/utility.h -- tolerance helpers (API of the reference's utility.h:12-27).
//
// epsilon<T>() is the cube root of the machine epsilon (about 6.06e-6 for double).  It is also the
// left end of StaticBlackhole::SolveG's bisection interval, so its exact value is part of the hot
// path's results.
#ifndef BLACKHOLE_CORE_NUMERIC_H_
#define BLACKHOLE_CORE_NUMERIC_H_

#include <cmath>
#include <limits>
#include <type_traits>


namespace blackhole {

template <typename T = double>
inline constexpr T kPi = static_cast<T>(3.141592653589793238462643383279);
template <typename T = double>
inline constexpr T kE = static_cast<T>(2.718281828459045235360287471352);

inline constexpr double pi = kPi<double>;
inline constexpr double e = kE<double>;

}  // namespace blackhole

namespace blackhole {

template <typename T>
struct type_identity {
  using type = T;
};
template <typename T>
using type_identity_t = typename type_identity<T>::type;

template <typename T>
inline auto epsilon() {
  static const auto value = std::cbrt(std::numeric_limits<T>::epsilon());
  return value;
}

// |x - y| <= epsilon of the common type; enabled when at least one side is floating point.
template <typename T, typename U,
          std::enable_if_t<std::is_floating_point_v<T> || std::is_floating_point_v<U>, int> = 0>
bool float_equal(T x, U y) {
  using common = std::common_type_t<T, U>;
  return std::abs(x - y) <= epsilon<common>();
}

}  // namespace blackhole

#endif  // BLACKHOLE_CORE_NUMERIC_H_
