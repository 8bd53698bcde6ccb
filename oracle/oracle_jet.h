// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, link or call anything under oracle/.
//
// Dense forward-mode dual numbers, restating the arithmetic of the reference's
// Jet<T,N> with plain arrays instead of Eigen::Matrix<T,N,1>.
//
// Follows (reference, read-only): include/ceres/jet.h
//   struct / ctors ............ :222-305
//   unary/binary + - * / ...... :309-404
//   abs, log, exp, sqrt, cos, sin, ... :530-660
//   hypot (2- and 3-arg) ...... :731-770
//   atan2 ..................... :1189-1202
//   fpclassify/isfinite ....... :1096-1130
// Eigen is only a fixed-size container there (element-wise FP64 + - *); the
// formulas are all in Ceres source, so this restatement reproduces results up
// to FMA contraction.  This file is compiled with -ffp-contract=off so that the
// oracle itself has no contraction.
#ifndef ORACLE_JET_H_
#define ORACLE_JET_H_

#include <cmath>
#include <limits>

namespace oracle {

// internal/ceres/array_utils.h: kImpossibleValue = 1e302.
constexpr double kImpossibleValue = 1e302;

template <int N>
struct Jet {
  double a;
  double v[N];

  // jet.h:232-237  default: a = 0, v = 0.
  Jet() : a(0.0) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  // jet.h:240-243  constant.
  Jet(const double& value) : a(value) {  // NOLINT (implicit like reference)
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  // jet.h:246-249  k-th unit perturbation.
  Jet(const double& value, int k) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
    v[k] = 1.0;
  }

  Jet& operator+=(const Jet& y) { *this = *this + y; return *this; }
  Jet& operator-=(const Jet& y) { *this = *this - y; return *this; }
  Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
  Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
  Jet& operator+=(const double& s) { a += s; return *this; }
  Jet& operator-=(const double& s) { a -= s; return *this; }
  Jet& operator*=(const double& s) { *this = *this * s; return *this; }
  Jet& operator/=(const double& s) { *this = *this / s; return *this; }
};

// jet.h:309-311
template <int N> inline Jet<N> const& operator+(const Jet<N>& f) { return f; }
// jet.h:317-320
template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a;
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h;
}
// jet.h:323-326
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i];
  return h;
}
// jet.h:329-338
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) {
  Jet<N> h = f; h.a = f.a + s; return h;
}
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) {
  Jet<N> h = f; h.a = f.a + s; return h;
}
// jet.h:341-344
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i];
  return h;
}
// jet.h:347-356
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) {
  Jet<N> h = f; h.a = f.a - s; return h;
}
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) {
  Jet<N> h; h.a = s - f.a;
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h;
}
// jet.h:359-362   h.v = f.a * g.v + f.v * g.a
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h;
}
// jet.h:365-374
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) {
  Jet<N> h; h.a = f.a * s;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s;
  return h;
}
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) {
  Jet<N> h; h.a = f.a * s;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s;
  return h;
}
// jet.h:377-389   (f.v - f.a/g.a * g.v) / g.a, via the reciprocal.
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  const double g_a_inverse = 1.0 / g.a;
  const double f_a_by_g_a = f.a * g_a_inverse;
  Jet<N> h; h.a = f_a_by_g_a;
  for (int i = 0; i < N; ++i)
    h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}
// jet.h:392-396
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) {
  const double minus_s_g_a_inverse2 = -s / (g.a * g.a);
  Jet<N> h; h.a = s / g.a;
  for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2;
  return h;
}
// jet.h:399-403
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) {
  const double s_inverse = 1.0 / s;
  Jet<N> h; h.a = f.a * s_inverse;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s_inverse;
  return h;
}

// jet.h:405-470 comparisons act on the scalar part only.
#define ORACLE_JET_CMP(op)                                                   \
  template <int N> inline bool operator op(const Jet<N>& f, const Jet<N>& g) { return f.a op g.a; } \
  template <int N> inline bool operator op(const Jet<N>& f, double s) { return f.a op s; }          \
  template <int N> inline bool operator op(double s, const Jet<N>& g) { return s op g.a; }
ORACLE_JET_CMP(<)
ORACLE_JET_CMP(<=)
ORACLE_JET_CMP(>)
ORACLE_JET_CMP(>=)
ORACLE_JET_CMP(==)
ORACLE_JET_CMP(!=)
#undef ORACLE_JET_CMP

template <int N> inline Jet<N> scaled_(const Jet<N>& f, double value, double c) {
  Jet<N> h; h.a = value;
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i];
  return h;
}

// jet.h:533-536  abs(x + h) ~= abs(x) + sgn(x) h
template <int N> inline Jet<N> abs(const Jet<N>& f) {
  return scaled_(f, std::abs(f.a), std::copysign(1.0, f.a));
}
// jet.h:579-583
template <int N> inline Jet<N> log(const Jet<N>& f) {
  const double a_inverse = 1.0 / f.a;
  Jet<N> h; h.a = std::log(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * a_inverse;
  return h;
}
// jet.h:601-605
template <int N> inline Jet<N> exp(const Jet<N>& f) {
  const double tmp = std::exp(f.a);
  return scaled_(f, tmp, tmp);
}
// jet.h:616-621
template <int N> inline Jet<N> sqrt(const Jet<N>& f) {
  const double tmp = std::sqrt(f.a);
  const double two_a_inverse = 1.0 / (2.0 * tmp);
  Jet<N> h; h.a = tmp;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
  return h;
}
// jet.h:624-627
template <int N> inline Jet<N> cos(const Jet<N>& f) {
  return scaled_(f, std::cos(f.a), -std::sin(f.a));
}
// jet.h:630-634
template <int N> inline Jet<N> acos(const Jet<N>& f) {
  const double tmp = -1.0 / std::sqrt(1.0 - f.a * f.a);
  return scaled_(f, std::acos(f.a), tmp);
}
// jet.h:637-640
template <int N> inline Jet<N> sin(const Jet<N>& f) {
  return scaled_(f, std::sin(f.a), std::cos(f.a));
}
// jet.h:643-647
template <int N> inline Jet<N> asin(const Jet<N>& f) {
  const double tmp = 1.0 / std::sqrt(1.0 - f.a * f.a);
  return scaled_(f, std::asin(f.a), tmp);
}
// jet.h:650-655
template <int N> inline Jet<N> tan(const Jet<N>& f) {
  const double tan_a = std::tan(f.a);
  const double tmp = 1.0 + tan_a * tan_a;
  return scaled_(f, tan_a, tmp);
}
// jet.h:658-662
template <int N> inline Jet<N> atan(const Jet<N>& f) {
  const double tmp = 1.0 / (1.0 + f.a * f.a);
  return scaled_(f, std::atan(f.a), tmp);
}
// jet.h:731-740  2-arg hypot: x.a/tmp * x.v + y.a/tmp * y.v
template <int N> inline Jet<N> hypot(const Jet<N>& x, const Jet<N>& y) {
  const double tmp = std::hypot(x.a, y.a);
  Jet<N> h; h.a = tmp;
  const double cx = x.a / tmp, cy = y.a / tmp;
  for (int i = 0; i < N; ++i) h.v[i] = cx * x.v[i] + cy * y.v[i];
  return h;
}
// jet.h:742-761  3-arg hypot (host path uses std::hypot(x,y,z)).
template <int N> inline Jet<N> hypot(const Jet<N>& x, const Jet<N>& y, const Jet<N>& z) {
  const double tmp = std::hypot(x.a, y.a, z.a);
  Jet<N> h; h.a = tmp;
  const double cx = x.a / tmp, cy = y.a / tmp, cz = z.a / tmp;
  for (int i = 0; i < N; ++i) h.v[i] = cx * x.v[i] + cy * y.v[i] + cz * z.v[i];
  return h;
}
// jet.h:1189-1202  atan2(g, f): tmp * (-g.a * f.v + f.a * g.v)
template <int N> inline Jet<N> atan2(const Jet<N>& g, const Jet<N>& f) {
  const double tmp = 1.0 / (f.a * f.a + g.a * g.a);
  Jet<N> h; h.a = std::atan2(g.a, f.a);
  for (int i = 0; i < N; ++i) h.v[i] = tmp * (-g.a * f.v[i] + f.a * g.v[i]);
  return h;
}
// jet.h:1204-1222 pow(Jet, double)
template <int N> inline Jet<N> pow(const Jet<N>& f, double g) {
  const double tmp = g * std::pow(f.a, g - 1.0);
  return scaled_(f, std::pow(f.a, g), tmp);
}

// jet.h:556-576 copysign
template <int N> inline Jet<N> copysign(const Jet<N>& f, const Jet<N> g) {
  const double d = std::fpclassify(g.a) == FP_ZERO ? std::numeric_limits<double>::infinity() : 0.0;
  const double sa = std::copysign(1.0, f.a);
  const double sb = std::copysign(1.0, g.a);
  Jet<N> h; h.a = std::copysign(f.a, g.a);
  for (int i = 0; i < N; ++i) h.v[i] = sa * sb * f.v[i] + std::abs(f.a) * d * g.v[i];
  return h;
}
// jet.h:586-598 log10 / log1p, :608-613 expm1
template <int N> inline Jet<N> log10(const Jet<N>& f) {
  const double a_inverse = 1.0 / (f.a * std::log(10.0));
  Jet<N> h; h.a = std::log10(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * a_inverse;
  return h;
}
template <int N> inline Jet<N> log1p(const Jet<N>& f) {
  const double a_inverse = 1.0 / (1.0 + f.a);
  Jet<N> h; h.a = std::log1p(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * a_inverse;
  return h;
}
template <int N> inline Jet<N> expm1(const Jet<N>& f) {
  const double tmp = std::expm1(f.a);
  return scaled_(f, tmp, tmp + 1.0);
}
// jet.h:665-682 sinh / cosh / tanh
template <int N> inline Jet<N> sinh(const Jet<N>& f) { return scaled_(f, std::sinh(f.a), std::cosh(f.a)); }
template <int N> inline Jet<N> cosh(const Jet<N>& f) { return scaled_(f, std::cosh(f.a), std::sinh(f.a)); }
template <int N> inline Jet<N> tanh(const Jet<N>& f) {
  const double t = std::tanh(f.a);
  return scaled_(f, t, 1.0 - t * t);
}
// jet.h:705-724 cbrt / exp2 / log2
template <int N> inline Jet<N> cbrt(const Jet<N>& f) {
  const double derivative = 1.0 / (3.0 * std::cbrt(f.a * f.a));
  Jet<N> h; h.a = std::cbrt(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * derivative;
  return h;
}
template <int N> inline Jet<N> exp2(const Jet<N>& f) {
  const double tmp = std::exp2(f.a);
  const double derivative = tmp * std::log(2.0);
  Jet<N> h; h.a = tmp;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * derivative;
  return h;
}
template <int N> inline Jet<N> log2(const Jet<N>& f) {
  const double derivative = 1.0 / (f.a * std::log(2.0));
  Jet<N> h; h.a = std::log2(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * derivative;
  return h;
}
// jet.h:763-772 fma
template <int N> inline Jet<N> fma(const Jet<N>& x, const Jet<N>& y, const Jet<N>& z) {
  Jet<N> h; h.a = std::fma(x.a, y.a, z.a);
  for (int i = 0; i < N; ++i) h.v[i] = y.a * x.v[i] + x.a * y.v[i] + z.v[i];
  return h;
}
// jet.h:806-846 fmax / fmin (Jet average on equality)
template <int N> inline Jet<N> fmax(const Jet<N>& x, const Jet<N>& y) {
  if (std::isnan(x.a) || std::isnan(y.a) || std::islessgreater(x.a, y.a))
    return (std::isnan(x.a) || std::isless(x.a, y.a)) ? y : x;
  return (x + y) * 0.5;
}
template <int N> inline Jet<N> fmin(const Jet<N>& x, const Jet<N>& y) {
  if (std::isnan(x.a) || std::isnan(y.a) || std::islessgreater(x.a, y.a))
    return (std::isnan(x.a) || std::isgreater(x.a, y.a)) ? y : x;
  return (x + y) * 0.5;
}
// jet.h:863-885 erf / erfc
template <int N> inline Jet<N> erf(const Jet<N>& x) {
  const double c = std::exp(-x.a * x.a) * (1.0 / std::sqrt(std::atan(1.0)));
  Jet<N> h; h.a = std::erf(x.a);
  for (int i = 0; i < N; ++i) h.v[i] = x.v[i] * std::exp(-x.a * x.a) * (1.0 / std::sqrt(std::atan(1.0)));
  (void)c;
  return h;
}
template <int N> inline Jet<N> erfc(const Jet<N>& x) {
  Jet<N> h; h.a = std::erfc(x.a);
  for (int i = 0; i < N; ++i) h.v[i] = -x.v[i] * std::exp(-x.a * x.a) * (1.0 / std::sqrt(std::atan(1.0)));
  return h;
}
// jet.h:1262-1290 pow(Jet, Jet), generic branch (f > 0)
template <int N> inline Jet<N> pow(const Jet<N>& f, const Jet<N>& g) {
  const double tmp1 = std::pow(f.a, g.a);
  const double tmp2 = g.a * std::pow(f.a, g.a - 1.0);
  const double tmp3 = tmp1 * std::log(f.a);
  Jet<N> h; h.a = tmp1;
  for (int i = 0; i < N; ++i) h.v[i] = tmp2 * f.v[i] + tmp3 * g.v[i];
  return h;
}
// jet.h:1096-1130: classification acts on the scalar part.
template <int N> inline int fpclassify(const Jet<N>& f) { return std::fpclassify(f.a); }
inline int fpclassify(double x) { return std::fpclassify(x); }

// Scalar overloads so functors templated on T compile with T = double.
using std::abs; using std::sqrt; using std::sin; using std::cos; using std::atan2;
using std::log; using std::exp; using std::acos; using std::asin; using std::tan;
using std::atan; using std::pow; using std::sinh; using std::cosh; using std::tanh;
using std::cbrt; using std::exp2; using std::log2; using std::log10; using std::log1p;
using std::expm1; using std::fmax; using std::fmin; using std::erf; using std::erfc;
using std::copysign; using std::fma;
inline double hypot(double x, double y) { return std::hypot(x, y); }
inline double hypot(double x, double y, double z) { return std::hypot(x, y, z); }

}  // namespace oracle

#endif  // ORACLE_JET_H_
