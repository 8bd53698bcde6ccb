// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// The REFERENCE's own arithmetic for the evaluation path, compiled from the sources
// where they lie under /root/reference (never copied) over oracle/eigen_shim:
//   include/ceres/jet.h, rotation.h, autodiff_cost_function.h (-> internal/autodiff.h),
//   internal/corrector.h, loss_function_cuda.h, examples/snavely_reprojection_error.h,
//   internal/ceres/autodiff_benchmarks/relative_pose_error.h,
//   examples/slam/pose_graph_3d/pose_graph_3d_error_term.h (+ types.h).
// Output: oracle/_ref/libref_arith.so (git-ignored).  Used by
// tests/golden/make_golden.py to validate the restatement (oracle_eval.cc) and to
// generate the golden vectors in tests/golden/.  /root/reference does not exist on
// the GPU box, so nothing at test/bench run time needs this library.
#include <cmath>

#include "ceres/autodiff_cost_function.h"
#include "ceres/internal/corrector.h"
#include "ceres/jet.h"
#include "ceres/loss_function_cuda.h"
#include "ceres/rotation.h"
#include "examples/slam/pose_graph_3d/pose_graph_3d_error_term.h"
#include "examples/snavely_reprojection_error.h"
#include "internal/ceres/autodiff_benchmarks/relative_pose_error.h"

namespace {
template <typename CostFunctionT>
int Eval(const CostFunctionT& f, const double* p0, const double* p1, double* residuals,
         double* j0, double* j1) {
  const double* params[2] = {p0, p1};
  double* jac[2] = {j0, j1};
  return f.Evaluate(params, residuals, (j0 || j1) ? jac : nullptr) ? 1 : 0;
}
}  // namespace

extern "C" {

// examples/snavely_reprojection_error.h:52-101 through AutoDiffCostFunction<.., 2, 9, 3>.
int ref_snavely(const double* camera, const double* point, const double* obs, double* residuals,
                double* jac_camera, double* jac_point) {
  ceres::AutoDiffCostFunction<ceres::examples::SnavelyReprojectionError, 2, 9, 3> f(
      new ceres::examples::SnavelyReprojectionError(obs[0], obs[1]));
  return Eval(f, camera, point, residuals, jac_camera, jac_point);
}

// examples/snavely_reprojection_error.h:108-179 <2, 10, 3>.
int ref_snavely_quaternions(const double* camera, const double* point, const double* obs,
                            double* residuals, double* jac_camera, double* jac_point) {
  ceres::AutoDiffCostFunction<ceres::examples::SnavelyReprojectionErrorWithQuaternions, 2, 10, 3>
      f(new ceres::examples::SnavelyReprojectionErrorWithQuaternions(obs[0], obs[1]));
  return Eval(f, camera, point, residuals, jac_camera, jac_point);
}

// include/ceres/loss_function_cuda.h:62-149.  kind: 1 trivial, 2 huber, 3 cauchy,
// 4 scaled huber, 5 scaled cauchy, 6 scaled trivial (same ids as problems.py).
void ref_loss(int kind, double a, double b, double s, double* rho) {
  switch (kind) {
    case 1: ceres::TrivialLossCUDA().Evaluate(s, rho); break;
    case 2: ceres::HuberLossCUDA(a).Evaluate(s, rho); break;
    case 3: ceres::CauchyLossCUDA(a).Evaluate(s, rho); break;
    case 4: ceres::ScaledLossCUDA<ceres::HuberLossCUDA>(ceres::HuberLossCUDA(a), b).Evaluate(s, rho); break;
    case 5: ceres::ScaledLossCUDA<ceres::CauchyLossCUDA>(ceres::CauchyLossCUDA(a), b).Evaluate(s, rho); break;
    case 6: ceres::ScaledLossCUDA<ceres::TrivialLossCUDA>(ceres::TrivialLossCUDA(), b).Evaluate(s, rho); break;
  }
}

// internal/ceres/autodiff_benchmarks/relative_pose_error.h:46-92 through
// AutoDiffCostFunction<RelativePoseError, 6, 7, 7>.  pose = [q(x, y, z, w), t];
// meas = [q_i_j(x, y, z, w), t_i_j].  Jacobians row-major 6 x 7.
int ref_relative_pose(const double* pose_i, const double* pose_j, const double* meas,
                      double* residuals, double* jac_i, double* jac_j) {
  ceres::AutoDiffCostFunction<ceres::RelativePoseError, 6, 7, 7> f(new ceres::RelativePoseError(
      Eigen::Quaterniond(meas[3], meas[0], meas[1], meas[2]),
      Eigen::Vector3d(meas[4], meas[5], meas[6])));
  return Eval(f, pose_i, pose_j, residuals, jac_i, jac_j);
}

// examples/slam/pose_graph_3d/pose_graph_3d_error_term.h:71-124 through its own Create():
// AutoDiffCostFunction<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4>.  data = [p_ab(3),
// q_ab(x, y, z, w), sqrt_information row-major 6 x 6].  jac: 4 row-major blocks
// 6 x 3, 6 x 4, 6 x 3, 6 x 4.
int ref_pose_graph_3d(const double* p_a, const double* q_a, const double* p_b, const double* q_b,
                      const double* data, double* residuals, double* j0, double* j1, double* j2,
                      double* j3) {
  ceres::examples::Pose3d t_ab;
  t_ab.p = Eigen::Vector3d(data[0], data[1], data[2]);
  t_ab.q = Eigen::Quaterniond(data[6], data[3], data[4], data[5]);
  Eigen::Matrix<double, 6, 6> sqrt_information;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) sqrt_information(i, j) = data[7 + 6 * i + j];
  std::unique_ptr<ceres::CostFunction> f(
      ceres::examples::PoseGraph3dErrorTerm::Create(t_ab, sqrt_information));
  const double* params[4] = {p_a, q_a, p_b, q_b};
  double* jac[4] = {j0, j1, j2, j3};
  return f->Evaluate(params, residuals, j0 ? jac : nullptr) ? 1 : 0;
}

// include/ceres/internal/corrector.h:82-213.
void ref_corrector(double sq_norm, const double* rho, int num_rows, int num_cols,
                   double* residuals, double* jacobian) {
  ceres::internal::Corrector c(sq_norm, rho);
  if (jacobian) c.CorrectJacobian(num_rows, num_cols, residuals, jacobian);
  c.CorrectResiduals(num_rows, residuals);
}

// include/ceres/rotation.h on plain doubles.
void ref_angle_axis_rotate_point(const double* aa, const double* pt, double* out) {
  ceres::AngleAxisRotatePoint(aa, pt, out);
}
void ref_quaternion_rotate_point(const double* q, const double* pt, double* out) {
  ceres::QuaternionRotatePoint(q, pt, out);
}
void ref_quaternion_to_angle_axis(const double* q, double* out) {
  ceres::QuaternionToAngleAxis(q, out);
}

// QuaternionToAngleAxis on Jets (rotation.h:356-406 with jet.h hypot/atan2):
// value[3] and the 3x4 Jacobian d angle_axis / d q, row-major.
void ref_quaternion_to_angle_axis_jet(const double* q, double* value, double* jac) {
  using J = ceres::Jet<double, 4>;
  J x[4], out[3];
  for (int i = 0; i < 4; ++i) x[i] = J(q[i], i);
  ceres::QuaternionToAngleAxis(x, out);
  for (int r = 0; r < 3; ++r) {
    value[r] = out[r].a;
    for (int c = 0; c < 4; ++c) jac[r * 4 + c] = out[r].v[c];
  }
}

// A battery of Jet<double, 2> operations at (x, y) = (in[0], in[1]) seeded as the two
// independent variables (inputs of internal/ceres/jet_cuda_test.cu.cc:108-110 are
// x = 2.3, y = 1.7).  out: 3 doubles (a, v0, v1) per operation.
int ref_jet_battery(const double* in, double* out) {
  using J = ceres::Jet<double, 2>;
  const J x(in[0], 0), y(in[1], 1);
  int k = 0;
  auto put = [&](const J& j) { out[k++] = j.a; out[k++] = j.v[0]; out[k++] = j.v[1]; };
  put(x + y); put(x - y); put(x * y); put(x / y); put(-x); put(x + 1.5); put(1.5 - x);
  put(x * 1.5); put(1.5 / x); put(x / 1.5);
  put(ceres::sqrt(x)); put(ceres::exp(x)); put(ceres::log(x)); put(ceres::sin(x));
  put(ceres::cos(x)); put(ceres::tan(x)); put(ceres::asin(x / 3.0)); put(ceres::acos(x / 3.0));
  put(ceres::atan(x)); put(ceres::sinh(x)); put(ceres::cosh(x)); put(ceres::tanh(x));
  put(ceres::abs(-x)); put(ceres::atan2(y, x)); put(ceres::pow(x, 1.7)); put(ceres::pow(x, y));
  put(ceres::hypot(x, y)); put(ceres::hypot(x, y, x * y)); put(ceres::cbrt(x));
  put(ceres::exp2(x)); put(ceres::log2(x)); put(ceres::log10(x)); put(ceres::log1p(x));
  put(ceres::expm1(x)); put(ceres::fmax(x, y)); put(ceres::fmin(x, y)); put(ceres::erf(x));
  put(ceres::erfc(x)); put(ceres::copysign(x, -y)); put(ceres::fma(x, y, x));
  return k / 3;
}

}  // extern "C"
