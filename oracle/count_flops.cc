// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Counts the ALGORITHMIC FP64 operations of one residual block: the reference
// arithmetic with a DENSE Jet<double, N> (every derivative lane computed, as
// include/ceres/jet.h does), plus loss, Corrector and J^T r.  This is the
// denominator-side figure for the FP64 roofline in DESIGN.md / bench.py
// ("replace this estimate by an exact count obtained by
// instantiating the functor with a counting scalar/Jet type in the oracle").
// Every +, -, *, / on a double counts 1; sqrt/sin/cos/atan2 count 1 each and are
// also reported separately.
//   g++ -std=c++17 -O1 -o /tmp/count_flops oracle/count_flops.cc && /tmp/count_flops
#include <cmath>
#include <cstdio>

namespace oracle {
struct Counter {
  long add = 0, mul = 0, div = 0, special = 0;
  long total() const { return add + mul + div + special; }
} g_count;

template <int N>
struct CJet {
  double a;
  double v[N];
  CJet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  CJet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT
  CJet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1; }
  CJet& operator+=(const CJet& y) { *this = *this + y; return *this; }
};
#define J CJet<N>
template <int N> J operator+(const J& f, const J& g) { J h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; g_count.add += 1 + N; return h; }
template <int N> J operator-(const J& f, const J& g) { J h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; g_count.add += 1 + N; return h; }
template <int N> J operator-(const J& f) { J h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> J operator+(const J& f, double s) { J h = f; h.a += s; g_count.add += 1; return h; }
template <int N> J operator+(double s, const J& f) { J h = f; h.a += s; g_count.add += 1; return h; }
template <int N> J operator-(const J& f, double s) { J h = f; h.a -= s; g_count.add += 1; return h; }
template <int N> J operator-(double s, const J& f) { J h = -f; h.a = s - f.a; g_count.add += 1; return h; }
template <int N> J operator*(const J& f, const J& g) { J h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; g_count.mul += 1 + 2 * N; g_count.add += N; return h; }
template <int N> J operator*(const J& f, double s) { J h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; g_count.mul += 1 + N; return h; }
template <int N> J operator*(double s, const J& f) { return f * s; }
template <int N> J operator/(const J& f, const J& g) { const double gi = 1.0 / g.a, q = f.a * gi; J h; h.a = q; for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; g_count.div += 1; g_count.mul += 1 + 2 * N; g_count.add += N; return h; }
template <int N> J operator/(double s, const J& g) { const double c = -s / (g.a * g.a); J h; h.a = s / g.a; for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * c; g_count.div += 2; g_count.mul += 1 + N; return h; }
template <int N> J sqrt(const J& f) { const double t = std::sqrt(f.a), c = 1.0 / (2.0 * t); J h; h.a = t; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * c; g_count.special += 1; g_count.div += 1; g_count.mul += 1 + N; return h; }
template <int N> J cos(const J& f) { const double c = -std::sin(f.a); J h; h.a = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; g_count.special += 2; g_count.mul += N; return h; }
template <int N> J sin(const J& f) { const double c = std::cos(f.a); J h; h.a = std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; g_count.special += 2; g_count.mul += N; return h; }
template <int N> J hypot(const J& x, const J& y, const J& z) {
  const double t = std::hypot(x.a, y.a, z.a); J h; h.a = t; const double cx = x.a / t, cy = y.a / t, cz = z.a / t;
  for (int i = 0; i < N; ++i) h.v[i] = cx * x.v[i] + cy * y.v[i] + cz * z.v[i];
  g_count.special += 1; g_count.mul += 3 /*squares*/ + 3 * N; g_count.add += 2 + 2 * N; g_count.div += 3; return h; }
template <int N> int fpclassify(const J& f) { return std::fpclassify(f.a); }
#undef J
}  // namespace oracle

#define ORACLE_JET_H_  // the functors below only need the operations defined above
namespace oracle { constexpr double kImpossibleValue = 1e302; }
#include "oracle_functors.h"

int main() {
  using namespace oracle;
  constexpr int N = 12;
  double cam[9] = {0.01, -0.02, 0.03, 0.1, -0.2, -8.0, 800.0, 1e-7, 1e-13}, pt[3] = {0.3, -0.4, 0.5};
  CJet<N> x[12], r[2];
  for (int i = 0; i < 9; ++i) x[i] = CJet<N>(cam[i], i);
  for (int i = 0; i < 3; ++i) x[9 + i] = CJet<N>(pt[i], 9 + i);
  const double obs[2] = {1.0, 2.0};
  g_count = Counter();
  SnavelyReprojectionError()(obs, x, x + 9, r);
  Counter f = g_count;
  // residual_block.cc:131 squared norm (2 mul + 1 add); Huber outlier branch
  // (loss_function_cuda.h:64-80: sqrt, 2 mul, 1 sub, 1 div, 1 div + mul); Corrector ctor
  // (corrector.h:82-147: sqrt + (3 mul, 1 div, 1 add, sqrt, 1 sub, 1 sub, 2 div));
  // CorrectJacobian full path per column (corrector.h:199-211): kRes (mul+add) +
  // kRes (2 mul + 1 sub + 1 mul); CorrectResiduals kRes mul; J^T r: N * kRes (mul + add).
  const int kRes = 2;
  long epilogue = 3 + 8 + 12 + N * (2 * kRes + 4 * kRes) + kRes + N * 2 * kRes + 1 /* 0.5 rho */;
  std::printf("SnavelyReprojectionError<2,9,3> dense Jet<12>: add %ld mul %ld div %ld special %ld -> %ld\n",
              f.add, f.mul, f.div, f.special, f.total());
  std::printf("loss + corrector + gradient (full Triggs path): %ld\n", epilogue);
  std::printf("ALGORITHMIC FLOPs per residual block: %ld\n", f.total() + epilogue);
  return 0;
}
