// Cost functors used only by the parity tests: device restatements of the
// functors in the reference's own tests for this path.
#ifndef TESTS_DRIVER_TEST_FUNCTORS_H_
#define TESTS_DRIVER_TEST_FUNCTORS_H_

#include "ceres/internal/cuda_defs.h"
#include "ceres/loss_function_cuda.h"
#include "ceres/rotation.h"

namespace test_functors {

// internal/ceres/evaluator_cuda_test.cu.cc:115-164  <2, 7, 3>
struct SnavelyReprojectionErrorNoRadialDistortion {
  HOST_DEVICE SnavelyReprojectionErrorNoRadialDistortion(double x, double y)
      : observed_x(x), observed_y(y) {}
  template <typename T>
  HOST_DEVICE bool operator()(const T* const camera, const T* const point, T* residuals) const {
    T p[3];
    ceres::AngleAxisRotatePoint(camera, point, p);
    p[0] += camera[3];
    p[1] += camera[4];
    p[2] += camera[5];
    const T xp = -p[0] / p[2];
    const T yp = -p[1] / p[2];
    const T& focal = camera[6];
    residuals[0] = focal * xp - observed_x;
    residuals[1] = focal * yp - observed_y;
    return true;
  }
  double observed_x, observed_y;
};

// evaluator_cuda_test.cu.cc:83-109  <3, 3>
struct PointDisplacementError {
  HOST_DEVICE PointDisplacementError(double x, double y, double z) : x_(x), y_(y), z_(z) {}
  template <typename T>
  HOST_DEVICE bool operator()(const T* const point, T* residuals) const {
    using ceres::abs;
    residuals[0] = abs(x_) - abs(point[0]);
    residuals[1] = abs(y_) - abs(point[1]);
    residuals[2] = abs(z_) - abs(point[2]);
    return true;
  }
  double x_, y_, z_;
};

// autodiff_cost_function_cuda_test.cu.cc:42-54  <1, 2, 2>
struct BinaryScalarCost {
  HOST_DEVICE explicit BinaryScalarCost(double a) : a_(a) {}
  template <typename T>
  HOST_DEVICE bool operator()(const T* const x, const T* const y, T* cost) const {
    cost[0] = x[0] * y[0] + x[1] * y[1] - T(a_);
    return true;
  }
  double a_;
};

// autodiff_cost_function_cuda_test.cu.cc:128-146  <1, 1 x 10>
struct TenParameterCost {
  template <typename T>
  HOST_DEVICE bool operator()(const T* const x0, const T* const x1, const T* const x2,
                              const T* const x3, const T* const x4, const T* const x5,
                              const T* const x6, const T* const x7, const T* const x8,
                              const T* const x9, T* cost) const {
    cost[0] = *x0 + *x1 + *x2 + *x3 + *x4 + *x5 + *x6 + *x7 + *x8 + *x9;
    return true;
  }
  char unused = 0;
};

// autodiff_cost_function_cuda_test.cu.cc:230-237  <2, 1>
struct OnlyFillsOneOutputFunctor {
  template <typename T>
  HOST_DEVICE bool operator()(const T* x, T* output) const {
    output[0] = x[0];
    return true;
  }
  char unused = 0;
};

// <1, 1>  r = x + sqrt(T(c)): a constant sub-expression whose derivative the reference
// computes as 0 * inf = NaN when c == 0 (tests/test_gpu_robustness.py).
struct SqrtOfConstantCost {
  HOST_DEVICE explicit SqrtOfConstantCost(double c) : c(c) {}
  template <typename T>
  HOST_DEVICE bool operator()(const T* x, T* residual) const {
    residual[0] = x[0] + sqrt(T(c));
    return true;
  }
  double c;
};

// A user-defined loss with positive curvature, rho(s) = s + a s^2: the only way into the
// Corrector's alpha branch (corrector.cc:105-130); none of the library's losses has rho'' > 0.
class ConvexTestLoss : public ceres::LossFunctionCUDABase {
 public:
  HOST_DEVICE explicit ConvexTestLoss(double a) : a_(a) {}
  HOST_DEVICE void Evaluate(double s, double rho[3]) const {
    rho[0] = s + a_ * s * s;
    rho[1] = 1.0 + 2.0 * a_ * s;
    rho[2] = 2.0 * a_;
  }

 private:
  double a_;
};

// Autodiff form of evaluator_test.cc:59-100 ParameterIgnoringCostFunction:
// residual i = (i + 1) + sum_k sum_j kFactor (j + 1) x_k[j].
template <int kFactor, int kNumResiduals, bool kSucceeds, int... Ns>
struct AffineTestCost {
  template <typename T>
  HOST_DEVICE bool operator()(const T* const a, const T* const b, T* residuals) const {
    const T* params[] = {a, b};
    return Impl(params, residuals);
  }
  template <typename T>
  HOST_DEVICE bool operator()(const T* const a, const T* const b, const T* const c,
                              T* residuals) const {
    const T* params[] = {a, b, c};
    return Impl(params, residuals);
  }
  template <typename T>
  HOST_DEVICE static bool Impl(const T* const* params, T* residuals) {
    constexpr int sizes[] = {Ns...};
#pragma unroll
    for (int i = 0; i < kNumResiduals; ++i) {
      T r(static_cast<double>(i + 1));
#pragma unroll
      for (int k = 0; k < static_cast<int>(sizeof...(Ns)); ++k) {
#pragma unroll
        for (int j = 0; j < sizes[k]; ++j)
          r += params[k][j] * static_cast<double>(kFactor * (j + 1));
      }
      residuals[i] = r;
    }
    return kSucceeds;
  }
  char unused = 0;
};

// evaluator_test.cc:573-596  <2, 2>
struct ParameterSensitiveCost {
  template <typename T>
  HOST_DEVICE bool operator()(const T* const x, T* residuals) const {
    residuals[0] = x[0] * x[0];
    residuals[1] = x[1] * x[1];
    return true;
  }
  char unused = 0;
};

// Every Jet operation, in the order of oracle/ref_arith.cc ref_jet_battery  <40, 2>
// (the operations internal/ceres/jet_cuda_test.cu.cc:123-1056 exercises on the device).
struct JetBatteryCost {
  template <typename T>
  HOST_DEVICE bool operator()(const T* const p, T* r) const {
    // Qualified calls: for T = double they pick the HOST_DEVICE overloads of ceres/jet.h
    // (the global 3-argument hypot is host-only), for T = Jet the Jet overloads.
    namespace c = ceres;
    const T& x = p[0];
    const T& y = p[1];
    r[0] = x + y; r[1] = x - y; r[2] = x * y; r[3] = x / y; r[4] = -x;
    r[5] = x + 1.5; r[6] = 1.5 - x; r[7] = x * 1.5; r[8] = 1.5 / x; r[9] = x / 1.5;
    r[10] = c::sqrt(x); r[11] = c::exp(x); r[12] = c::log(x); r[13] = c::sin(x);
    r[14] = c::cos(x); r[15] = c::tan(x); r[16] = c::asin(x / 3.0); r[17] = c::acos(x / 3.0);
    r[18] = c::atan(x); r[19] = c::sinh(x); r[20] = c::cosh(x); r[21] = c::tanh(x);
    r[22] = c::abs(-x); r[23] = c::atan2(y, x); r[24] = c::pow(x, 1.7); r[25] = c::pow(x, y);
    r[26] = c::hypot(x, y); r[27] = c::hypot(x, y, x * y); r[28] = c::cbrt(x);
    r[29] = c::exp2(x); r[30] = c::log2(x); r[31] = c::log10(x); r[32] = c::log1p(x);
    r[33] = c::expm1(x); r[34] = c::fmax(x, y); r[35] = c::fmin(x, y); r[36] = c::erf(x);
    r[37] = c::erfc(x); r[38] = c::copysign(x, -y); r[39] = c::fma(x, y, x);
    return true;
  }
  char unused = 0;
};

}  // namespace test_functors

#endif  // TESTS_DRIVER_TEST_FUNCTORS_H_
