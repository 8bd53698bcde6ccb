// Loss functions evaluated on the device.
//
// Same classes and Evaluate(s, rho[3]) contract as the reference's
// include/ceres/loss_function_cuda.h:52-149 (which mirror the CPU losses of
// internal/ceres/loss_function.cc:44-82): rho[0] = rho(s), rho[1] = rho'(s),
// rho[2] = rho''(s) for s = ||r||^2.  No virtual Evaluate: the kernel is
// instantiated per loss type.  LossFunctionCUDABase exists only so ProblemCUDA can
// own heterogeneous loss objects (problem_cuda.h:455-460).
#ifndef CERES_B200_LOSS_FUNCTION_CUDA_H_
#define CERES_B200_LOSS_FUNCTION_CUDA_H_

#include <cmath>
#include <limits>

#include "ceres/internal/cuda_defs.h"

namespace ceres {

class LossFunctionCUDABase {
 public:
  virtual ~LossFunctionCUDABase() {}
};

namespace loss_internal {
// rho <- (value, first derivative, second derivative)
HOST_DEVICE inline void Set(double rho[3], double value, double d1, double d2) {
  rho[0] = value;
  rho[1] = d1;
  rho[2] = d2;
}
// Square root and its reciprocal of a positive finite x, for device code: one MUFU.RSQ64H with
// its Newton step (rsqrt) and one correction step for the root, instead of the library's
// sqrt followed by a division (each a 20-instruction sequence with a slow-path call).  The
// root is within 1 ulp; anything else (0, negative, Inf, NaN) takes the library functions.
HOST_DEVICE inline void RootAndReciprocal(double x, double* root, double* inv_root) {
#ifdef __CUDA_ARCH__
  if (x > 0.0 && x < 1.7976931348623157e308) {
    const double r = rsqrt(x);
    const double y = x * r;
    *root = fma(fma(-y, y, x), 0.5 * r, y);
    *inv_root = r;
    return;
  }
#endif
  *root = sqrt(x);
  *inv_root = 1.0 / *root;
}
// max(x, numeric_limits<double>::min()): rho' is kept strictly positive
// (loss_function.cc:56,70).
HOST_DEVICE inline double AtLeastTiny(double x) {
  const double tiny = 2.2250738585072014e-308;
  return x > tiny ? x : tiny;
}
}  // namespace loss_internal

// rho(s) = s
class TrivialLossCUDA : public LossFunctionCUDABase {
 public:
  // rho'' <= 0 everywhere: the evaluation kernel skips the Corrector's alpha branch
  // (corrector.cc:105-110 takes it only for rho'' > 0) and never needs rho[2].
  static constexpr bool kNonPositiveCurvature = true;
  HOST_DEVICE void Evaluate(double s, double rho[3]) const {
    loss_internal::Set(rho, s, 1.0, 0.0);
  }
};

// rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 beyond.
class HuberLossCUDA : public LossFunctionCUDABase {
 public:
  // rho'' <= 0 everywhere: the evaluation kernel skips the Corrector's alpha branch
  // (corrector.cc:105-110 takes it only for rho'' > 0) and never needs rho[2].
  static constexpr bool kNonPositiveCurvature = true;
  HOST_DEVICE explicit HuberLossCUDA(double a) : a_(a), b_(a * a) {}
  HOST_DEVICE void Evaluate(double s, double rho[3]) const {
    if (!(s > b_)) {  // inlier region (and NaN, as in the reference): the identity
      loss_internal::Set(rho, s, 1.0, 0.0);
      return;
    }
    double root, inv_root;
    loss_internal::RootAndReciprocal(s, &root, &inv_root);
#ifdef __CUDA_ARCH__
    const double slope = loss_internal::AtLeastTiny(a_ * inv_root);
#else
    // (host evaluation, e.g. Problem::EvaluateResidualBlock: the reference's expression)
    const double slope = loss_internal::AtLeastTiny(a_ / root);
#endif
    loss_internal::Set(rho, 2.0 * a_ * root - b_, slope, -slope / (2.0 * s));
  }

 private:
  double a_;
  double b_;
};

// rho(s) = a^2 log(1 + s / a^2)
class CauchyLossCUDA : public LossFunctionCUDABase {
 public:
  // rho'' <= 0 everywhere: the evaluation kernel skips the Corrector's alpha branch
  // (corrector.cc:105-110 takes it only for rho'' > 0) and never needs rho[2].
  static constexpr bool kNonPositiveCurvature = true;
  HOST_DEVICE explicit CauchyLossCUDA(double a) : b_(a * a), c_(1 / b_) {}
  HOST_DEVICE void Evaluate(double s, double rho[3]) const {
    const double one_plus = 1.0 + s * c_;
    const double reciprocal = 1.0 / one_plus;
    loss_internal::Set(rho, b_ * log(one_plus), loss_internal::AtLeastTiny(reciprocal),
                       -c_ * (reciprocal * reciprocal));
  }

 private:
  double b_;
  double c_;
};

// a * rho(s)
template <typename LossFunctionCUDA>
class ScaledLossCUDA : public LossFunctionCUDABase {
 public:
  HOST_DEVICE ScaledLossCUDA(const LossFunctionCUDA& rho, double a) : rho_(rho), a_(a) {}
  HOST_DEVICE void Evaluate(double s, double rho[3]) const {
    rho_.Evaluate(s, rho);
    loss_internal::Set(rho, rho[0] * a_, rho[1] * a_, rho[2] * a_);
  }

 private:
  LossFunctionCUDA rho_;
  double a_;
};

template <>
class ScaledLossCUDA<TrivialLossCUDA> : public LossFunctionCUDABase {
 public:
  HOST_DEVICE ScaledLossCUDA(const TrivialLossCUDA&, double a) : a_(a) {}
  HOST_DEVICE void Evaluate(double s, double rho[3]) const {
    loss_internal::Set(rho, a_ * s, a_, 0.0);
  }

 private:
  double a_;
};

}  // namespace ceres

#endif  // CERES_B200_LOSS_FUNCTION_CUDA_H_
