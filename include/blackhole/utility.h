// blackhole/utility.h -- tolerance helpers (API of the reference's utility.h:12-27).
//
// epsilon<T>() is the cube root of the machine epsilon (about 6.06e-6 for double).  It is also the
// left end of StaticBlackhole::SolveG's bisection interval, so its exact value is part of the hot
// path's results.
#ifndef BLACKHOLE_UTILITY_H_
#define BLACKHOLE_UTILITY_H_

#include <cmath>
#include <limits>
#include <type_traits>

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

#endif  // BLACKHOLE_UTILITY_H_
