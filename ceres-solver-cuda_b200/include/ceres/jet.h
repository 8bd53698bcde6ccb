// Forward-mode dual numbers for the B200 evaluation kernels.
//
// Same user-visible contract as the reference's Jet<T, N> (include/ceres/jet.h:222-305):
// a scalar part `a`, an infinitesimal part `v[0..N)`, arithmetic, comparisons on
// the scalar part and the <cmath> overloads cost functors use (jet.h:309-404,
// 530-900, 1096-1250).  Functors written against the reference's Jet compile
// unchanged.
//
// What is different, and why.  On an SM the FP64 pipe retires 16 lanes per clock
// per scheduler, so a dense Jet<double, 12> product (25 FP64 instructions) is the
// unit the whole kernel is paid in.  Most of those are multiplications by
// structural zeros: autodiff seeds each input with a unit vector
// (autodiff.h:186-204), so e.g. in a reprojection error the rotated point only
// ever depends on the 3 rotation and 3 point coordinates.  Every Jet here carries
// a bit mask `m` of the derivative lanes that can be non-zero.  The seeds'
// masks are literals, all Jet code is force-inlined and fully unrolled, so after
// constant propagation the masks exist only at compile time: lanes known to be
// zero generate no instructions and occupy no registers.  (IEEE semantics stop
// the compiler from doing this on the values themselves: 0 * x is not 0 for
// x = inf.)  For SnavelyReprojectionError<2,9,3> this takes the kernel from 736
// to 501 FP64 instructions and from 164 to 94 registers per thread, which is
// what moves it from the FP64 roof to the HBM roof ("Jets").
//
// Invariant: bit i of m clear  =>  v[i] == 0.0 exactly.
//
// The only observable difference from the reference: where the reference would
// compute 0 * inf = NaN in a derivative lane that is structurally zero, this
// Jet keeps the exact zero.
#ifndef CERES_B200_JET_H_
#define CERES_B200_JET_H_

#include <cmath>
#include <cstdint>
#include <limits>
#include <type_traits>

#include "ceres/internal/cuda_defs.h"

#define CERES_B200_JET_FN HOST_DEVICE CERES_B200_INLINE

namespace ceres {

namespace jet_internal {
// Three-argument hypot.  The device has no std::hypot(x, y, z); like the
// reference's device path (include/ceres/internal/cudamath/cuda_math.h:49-65)
// scale by the largest magnitude so the squares cannot overflow/underflow.
CERES_B200_JET_FN double Hypot3(double x, double y, double z) {
  const double px = ::fabs(x), py = ::fabs(y), pz = ::fabs(z);
  const double big = ::fmax(px, ::fmax(py, pz));
  if (big == 0.0) return 0.0;
  const double inv = 1.0 / big;
  const double sx = px * inv, sy = py * inv, sz = pz * inv;
  return big * ::sqrt(sx * sx + sy * sy + sz * sz);
}
CERES_B200_JET_FN bool IsFinite(double x) { return ::fabs(x) <= 1.7976931348623157e308; }
CERES_B200_JET_FN bool IsInf(double x) { return ::fabs(x) > 1.7976931348623157e308; }
CERES_B200_JET_FN int FpClassify(double x) {
#ifdef DEVICE_CODE
  // No std::fpclassify in device code (cuda_math.h:67-96 does the same by bits).
  const unsigned long long u = __double_as_longlong(x) & 0x7fffffffffffffffULL;
  const unsigned e = static_cast<unsigned>(u >> 52);
  const unsigned long long frac = u & 0x000fffffffffffffULL;
  if (e == 0) return frac ? FP_SUBNORMAL : FP_ZERO;
  if (e == 0x7ff) return frac ? FP_NAN : FP_INFINITE;
  return FP_NORMAL;
#else
  return std::fpclassify(x);
#endif
}
}  // namespace jet_internal

template <typename T, int N>
struct Jet {
  static_assert(N >= 1 && N <= 64, "Jet: 1 <= N <= 64 derivative lanes are supported");
  enum { DIMENSION = N };
  using Scalar = T;
  using Mask = unsigned long long;

  T a;
  T v[N];
  Mask m;  // lanes of v that may be non-zero

  static CERES_B200_JET_FN Mask FullMask() {
    return N == 64 ? ~0ULL : ((1ULL << N) - 1ULL);
  }

  CERES_B200_JET_FN Jet() : a(), m(0) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  CERES_B200_JET_FN explicit Jet(const T& value) : a(value), m(0) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  // The k-th unit perturbation (jet.h:246-249).
  CERES_B200_JET_FN Jet(const T& value, int k) : a(value), m(1ULL << k) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (i == k) ? T(1.0) : T();
  }
  // Every lane live, filled with `fill` (used to invalidate outputs).
  static CERES_B200_JET_FN Jet Filled(T value, T fill) {
    Jet h;
    h.a = value;
    h.m = FullMask();
#pragma unroll
    for (int i = 0; i < N; ++i) h.v[i] = fill;
    return h;
  }

  CERES_B200_JET_FN bool lane(int i) const { return (m >> i) & 1ULL; }

  CERES_B200_JET_FN Jet& operator+=(const Jet& y) { *this = *this + y; return *this; }
  CERES_B200_JET_FN Jet& operator-=(const Jet& y) { *this = *this - y; return *this; }
  CERES_B200_JET_FN Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
  CERES_B200_JET_FN Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
  CERES_B200_JET_FN Jet& operator+=(const T& s) { a += s; return *this; }
  CERES_B200_JET_FN Jet& operator-=(const T& s) { a -= s; return *this; }
  CERES_B200_JET_FN Jet& operator*=(const T& s) { *this = *this * s; return *this; }
  CERES_B200_JET_FN Jet& operator/=(const T& s) { *this = *this / s; return *this; }
};

namespace jet_internal {
// h = value + deriv * f.v  (the chain rule for a unary function).
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> Chain(const Jet<T, N>& f, const T& value, const T& deriv) {
  Jet<T, N> h;
  h.a = value;
  h.m = f.m;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.lane(i) ? deriv * f.v[i] : T();
  return h;
}
// h = value + cf * f.v + cg * g.v
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> Chain2(const Jet<T, N>& f, const Jet<T, N>& g, const T& value,
                                  const T& cf, const T& cg) {
  Jet<T, N> h;
  h.a = value;
  h.m = f.m | g.m;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const bool fi = f.lane(i), gi = g.lane(i);
    h.v[i] = (fi && gi) ? cf * f.v[i] + cg * g.v[i]
             : fi       ? cf * f.v[i]
             : gi       ? cg * g.v[i]
                        : T();
  }
  return h;
}
}  // namespace jet_internal

// ------------------------------------------------------------ + - * /
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> const& operator+(const Jet<T, N>& f) { return f; }

template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = -f.a;
  h.m = f.m;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.lane(i) ? -f.v[i] : T();
  return h;
}

template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a + g.a;
  h.m = f.m | g.m;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const bool fi = f.lane(i), gi = g.lane(i);
    h.v[i] = (fi && gi) ? f.v[i] + g.v[i] : fi ? f.v[i] : gi ? g.v[i] : T();
  }
  return h;
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator+(const Jet<T, N>& f, T s) {
  Jet<T, N> h = f;
  h.a = f.a + s;
  return h;
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator+(T s, const Jet<T, N>& f) {
  Jet<T, N> h = f;
  h.a = f.a + s;
  return h;
}

template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a - g.a;
  h.m = f.m | g.m;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const bool fi = f.lane(i), gi = g.lane(i);
    h.v[i] = (fi && gi) ? f.v[i] - g.v[i] : fi ? f.v[i] : gi ? -g.v[i] : T();
  }
  return h;
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator-(const Jet<T, N>& f, T s) {
  Jet<T, N> h = f;
  h.a = f.a - s;
  return h;
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator-(T s, const Jet<T, N>& f) {
  Jet<T, N> h = -f;
  h.a = s - f.a;
  return h;
}

// d(fg) = f dg + g df  (jet.h:359-362)
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  return jet_internal::Chain2(g, f, f.a * g.a, f.a, g.a);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator*(const Jet<T, N>& f, T s) {
  return jet_internal::Chain(f, f.a * s, s);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator*(T s, const Jet<T, N>& f) {
  return jet_internal::Chain(f, f.a * s, s);
}

// (f.v - f.a/g.a * g.v) / g.a through one reciprocal (jet.h:377-389).
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  const T g_inv = T(1.0) / g.a;
  const T q = f.a * g_inv;
  Jet<T, N> h;
  h.a = q;
  h.m = f.m | g.m;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const bool fi = f.lane(i), gi = g.lane(i);
    h.v[i] = (fi && gi) ? (f.v[i] - q * g.v[i]) * g_inv
             : fi       ? f.v[i] * g_inv
             : gi       ? (-(q * g.v[i])) * g_inv
                        : T();
  }
  return h;
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator/(T s, const Jet<T, N>& g) {
  // s / g.a and -s / g.a^2 from one reciprocal (the reference divides twice; the
  // results agree to an ulp or two).
  const T g_inv = T(1.0) / g.a;
  const T q = s * g_inv;
  return jet_internal::Chain(g, q, -q * g_inv);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> operator/(const Jet<T, N>& f, T s) {
  const T s_inv = T(1.0) / s;
  return jet_internal::Chain(f, f.a * s_inv, s_inv);
}

// ------------------------------------------------- comparisons (scalar part)
#define CERES_B200_JET_COMPARE(op)                                                          \
  template <typename T, int N>                                                              \
  CERES_B200_JET_FN bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
  template <typename T, int N, typename S,                                                  \
            typename std::enable_if<std::is_arithmetic<S>::value, int>::type = 0>           \
  CERES_B200_JET_FN bool operator op(const Jet<T, N>& f, S s) { return f.a op s; }          \
  template <typename T, int N, typename S,                                                  \
            typename std::enable_if<std::is_arithmetic<S>::value, int>::type = 0>           \
  CERES_B200_JET_FN bool operator op(S s, const Jet<T, N>& g) { return s op g.a; }
CERES_B200_JET_COMPARE(<)
CERES_B200_JET_COMPARE(<=)
CERES_B200_JET_COMPARE(>)
CERES_B200_JET_COMPARE(>=)
CERES_B200_JET_COMPARE(==)
CERES_B200_JET_COMPARE(!=)
#undef CERES_B200_JET_COMPARE

// ----------------------------------------------------------- <cmath> family
// Plain-double versions so functors templated on T compile for T = double too.
using std::abs;
using std::acos;
using std::asin;
using std::atan;
using std::atan2;
using std::cbrt;
using std::ceil;
using std::copysign;
using std::cos;
using std::cosh;
using std::erf;
using std::erfc;
using std::exp;
using std::exp2;
using std::expm1;
using std::floor;
using std::fma;
using std::fmax;
using std::fmin;
using std::fdim;
using std::isfinite;
using std::isinf;
using std::isnan;
using std::isnormal;
using std::log;
using std::log10;
using std::log1p;
using std::log2;
using std::pow;
using std::sin;
using std::sinh;
using std::sqrt;
using std::tan;
using std::tanh;

CERES_B200_JET_FN double hypot(double x, double y) { return ::hypot(x, y); }
CERES_B200_JET_FN double hypot(double x, double y, double z) {
  return jet_internal::Hypot3(x, y, z);
}
CERES_B200_JET_FN int fpclassify(double x) { return jet_internal::FpClassify(x); }

template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> abs(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::fabs(f.a)), T(::copysign(1.0, f.a)));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> log(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::log(f.a)), T(1.0) / f.a);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> log10(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::log10(f.a)), T(1.0) / (f.a * 2.302585092994045684));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> log1p(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::log1p(f.a)), T(1.0) / (T(1.0) + f.a));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> log2(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::log2(f.a)), T(1.0) / (f.a * 0.693147180559945309));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> exp(const Jet<T, N>& f) {
  const T e = ::exp(f.a);
  return jet_internal::Chain(f, e, e);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> expm1(const Jet<T, N>& f) {
  const T e = ::expm1(f.a);
  return jet_internal::Chain(f, e, e + T(1.0));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> exp2(const Jet<T, N>& f) {
  const T e = ::exp2(f.a);
  return jet_internal::Chain(f, e, e * 0.693147180559945309);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> sqrt(const Jet<T, N>& f) {
  const T r = ::sqrt(f.a);
  return jet_internal::Chain(f, r, T(1.0) / (T(2.0) * r));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> cbrt(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::cbrt(f.a)), T(1.0) / (T(3.0) * ::cbrt(f.a * f.a)));
}
// On the device sin and cos come from one sincos(): the argument reduction is shared, and
// when a functor takes both of the same angle (AngleAxisRotatePoint) the two calls fold
// into one.
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> cos(const Jet<T, N>& f) {
#ifdef __CUDA_ARCH__
  double s, c;
  ::sincos(f.a, &s, &c);
  return jet_internal::Chain(f, T(c), -T(s));
#else
  return jet_internal::Chain(f, T(::cos(f.a)), -T(::sin(f.a)));
#endif
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> sin(const Jet<T, N>& f) {
#ifdef __CUDA_ARCH__
  double s, c;
  ::sincos(f.a, &s, &c);
  return jet_internal::Chain(f, T(s), T(c));
#else
  return jet_internal::Chain(f, T(::sin(f.a)), T(::cos(f.a)));
#endif
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> tan(const Jet<T, N>& f) {
  const T t = ::tan(f.a);
  return jet_internal::Chain(f, t, T(1.0) + t * t);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> acos(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::acos(f.a)), -T(1.0) / ::sqrt(T(1.0) - f.a * f.a));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> asin(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::asin(f.a)), T(1.0) / ::sqrt(T(1.0) - f.a * f.a));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> atan(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::atan(f.a)), T(1.0) / (T(1.0) + f.a * f.a));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> sinh(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::sinh(f.a)), T(::cosh(f.a)));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> cosh(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::cosh(f.a)), T(::sinh(f.a)));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> tanh(const Jet<T, N>& f) {
  const T t = ::tanh(f.a);
  return jet_internal::Chain(f, t, T(1.0) - t * t);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> floor(const Jet<T, N>& f) { return Jet<T, N>(T(::floor(f.a))); }
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> ceil(const Jet<T, N>& f) { return Jet<T, N>(T(::ceil(f.a))); }
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> erf(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::erf(f.a)), T(::exp(-f.a * f.a) * 1.128379167095512574));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> erfc(const Jet<T, N>& f) {
  return jet_internal::Chain(f, T(::erfc(f.a)), T(-::exp(-f.a * f.a) * 1.128379167095512574));
}

// d hypot = (x dx + y dy) / hypot  (jet.h:731-761)
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> hypot(const Jet<T, N>& x, const Jet<T, N>& y) {
  const T h = ::hypot(x.a, y.a);
  const T inv = T(1.0) / h;
  return jet_internal::Chain2(x, y, h, x.a * inv, y.a * inv);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> hypot(const Jet<T, N>& x, const Jet<T, N>& y, const Jet<T, N>& z) {
  const T h = jet_internal::Hypot3(x.a, y.a, z.a);
  const T inv = T(1.0) / h;
  const Jet<T, N> xy = jet_internal::Chain2(x, y, h, x.a * inv, y.a * inv);
  // xy.v + (z.a/h) z.v
  Jet<T, N> r;
  r.a = h;
  r.m = xy.m | z.m;
  const T cz = z.a * inv;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const bool pi = xy.lane(i), zi = z.lane(i);
    r.v[i] = (pi && zi) ? xy.v[i] + cz * z.v[i] : pi ? xy.v[i] : zi ? cz * z.v[i] : T();
  }
  return r;
}

// fma(x, y, z): y dx + x dy + dz  (jet.h:763-772)
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> fma(const Jet<T, N>& x, const Jet<T, N>& y, const Jet<T, N>& z) {
  Jet<T, N> p = jet_internal::Chain2(x, y, T(::fma(x.a, y.a, z.a)), y.a, x.a);
  Jet<T, N> h = p + z;
  h.a = p.a;
  return h;
}

// atan2(g, f): (f dg - g df) / (f^2 + g^2)  (jet.h:1189-1202)
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
  const T tmp = T(1.0) / (f.a * f.a + g.a * g.a);
  return jet_internal::Chain2(f, g, T(::atan2(g.a, f.a)), -g.a * tmp, f.a * tmp);
}

// pow: the three overloads of jet.h:1204-1290.
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> pow(const Jet<T, N>& f, double g) {
  return jet_internal::Chain(f, T(::pow(f.a, g)), T(g * ::pow(f.a, g - 1.0)));
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> pow(T f, const Jet<T, N>& g) {
  Jet<T, N> result;
  if (jet_internal::FpClassify(f) == FP_ZERO && g > 0) {
    result = Jet<T, N>(T(0.0));
  } else if (f < 0 && g == ::floor(g.a)) {
    result = Jet<T, N>(T(::pow(f, g.a)));
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (g.lane(i) && jet_internal::FpClassify(g.v[i]) != FP_ZERO) {
        result.v[i] = std::numeric_limits<T>::quiet_NaN();
        result.m |= 1ULL << i;
      }
    }
  } else {
    const T tmp = ::pow(f, g.a);
    result = jet_internal::Chain(g, tmp, T(::log(f) * tmp));
  }
  return result;
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> pow(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> result;
  if (jet_internal::FpClassify(f.a) == FP_ZERO && g >= 1) {
    if (g > 1) {
      result = Jet<T, N>(T(0.0));
    } else {
      result = f;
    }
  } else {
    if (f < 0 && g == ::floor(g.a)) {
      const T tmp = g.a * ::pow(f.a, g.a - T(1.0));
      result = jet_internal::Chain(f, T(::pow(f.a, g.a)), tmp);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (g.lane(i) && jet_internal::FpClassify(g.v[i]) != FP_ZERO) {
          result.v[i] = T(std::numeric_limits<double>::quiet_NaN());
          result.m |= 1ULL << i;
        }
      }
    } else {
      const T tmp1 = ::pow(f.a, g.a);
      const T tmp2 = g.a * ::pow(f.a, g.a - T(1.0));
      const T tmp3 = tmp1 * ::log(f.a);
      result = jet_internal::Chain2(f, g, tmp1, tmp2, tmp3);
    }
  }
  return result;
}

// copysign(f, g) = sgn(g)|f|  (jet.h:556-576)
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> copysign(const Jet<T, N>& f, const Jet<T, N> g) {
  const T d = jet_internal::FpClassify(g.a) == FP_ZERO
                  ? std::numeric_limits<T>::infinity()
                  : T(0);
  const T sa = ::copysign(T(1), f.a);
  const T sb = ::copysign(T(1), g.a);
  // The g lanes carry |f| * delta(g) which is 0, inf or NaN; they are live.
  Jet<T, N> h;
  h.a = ::copysign(f.a, g.a);
  h.m = f.m | g.m;
  const T cg = ::fabs(f.a) * d;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const T fv = f.lane(i) ? f.v[i] : T();
    const T gv = g.lane(i) ? g.v[i] : T();
    h.v[i] = ((h.m >> i) & 1ULL) ? sa * sb * fv + cg * gv : T();
  }
  return h;
}

// fmax / fmin with Jet averaging on equality (jet.h:806-846).
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> fmax(const Jet<T, N>& x, const Jet<T, N>& y) {
  if (x.a != x.a) return y;
  if (y.a != y.a) return x;
  if (x.a < y.a) return y;
  if (x.a > y.a) return x;
  return (x + y) * T(0.5);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> fmax(const Jet<T, N>& x, T y) { return fmax(x, Jet<T, N>(y)); }
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> fmax(T x, const Jet<T, N>& y) { return fmax(Jet<T, N>(x), y); }
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> fmin(const Jet<T, N>& x, const Jet<T, N>& y) {
  if (x.a != x.a) return y;
  if (y.a != y.a) return x;
  if (x.a > y.a) return y;
  if (x.a < y.a) return x;
  return (x + y) * T(0.5);
}
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> fmin(const Jet<T, N>& x, T y) { return fmin(x, Jet<T, N>(y)); }
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> fmin(T x, const Jet<T, N>& y) { return fmin(Jet<T, N>(x), y); }
// fdim (jet.h:848-861)
template <typename T, int N>
CERES_B200_JET_FN Jet<T, N> fdim(const Jet<T, N>& f, const Jet<T, N>& g) {
  if (f.a != f.a || g.a != g.a)
    return Jet<T, N>::Filled(std::numeric_limits<T>::quiet_NaN(),
                             std::numeric_limits<T>::quiet_NaN());
  return f.a > g.a ? f - g : Jet<T, N>();
}

// Classification looks at the scalar part for fpclassify (jet.h:1096-1130); the
// is* predicates look at every live lane like the reference's.
template <typename T, int N>
CERES_B200_JET_FN int fpclassify(const Jet<T, N>& f) { return jet_internal::FpClassify(f.a); }
template <typename T, int N>
CERES_B200_JET_FN bool isfinite(const Jet<T, N>& f) {
  bool ok = jet_internal::IsFinite(f.a);
#pragma unroll
  for (int i = 0; i < N; ++i) ok = ok && (!f.lane(i) || jet_internal::IsFinite(f.v[i]));
  return ok;
}
template <typename T, int N>
CERES_B200_JET_FN bool isnan(const Jet<T, N>& f) {
  bool any = f.a != f.a;
#pragma unroll
  for (int i = 0; i < N; ++i) any = any || (f.lane(i) && f.v[i] != f.v[i]);
  return any;
}
template <typename T, int N>
CERES_B200_JET_FN bool isinf(const Jet<T, N>& f) {
  bool any = jet_internal::IsInf(f.a);
#pragma unroll
  for (int i = 0; i < N; ++i) any = any || (f.lane(i) && jet_internal::IsInf(f.v[i]));
  return any;
}
template <typename T, int N>
CERES_B200_JET_FN bool isnormal(const Jet<T, N>& f) {
  bool ok = jet_internal::FpClassify(f.a) == FP_NORMAL;
#pragma unroll
  for (int i = 0; i < N; ++i)
    ok = ok && (f.lane(i) ? jet_internal::FpClassify(f.v[i]) == FP_NORMAL : false);
  return ok;
}

}  // namespace ceres

#endif  // CERES_B200_JET_H_
