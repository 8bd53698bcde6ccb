// Manifolds whose PlusJacobian feeds the evaluation kernel's J * PlusJacobian.
//
// Same interface as the reference's ceres::Manifold (include/ceres/manifold.h)
// restricted to what the evaluation path and Program::Plus use: AmbientSize,
// TangentSize, Plus, PlusJacobian.  Implementations follow
// internal/ceres/manifold.cc:28-79 (quaternions), :184-214 (SubsetManifold) and
// include/ceres/product_manifold.h:117-124,200-218 (block-diagonal product).
// Jacobians are row-major AmbientSize x TangentSize.
#ifndef CERES_B200_MANIFOLD_H_
#define CERES_B200_MANIFOLD_H_

#include <cmath>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace ceres {

class Manifold {
 public:
  virtual ~Manifold() {}
  virtual int AmbientSize() const = 0;
  virtual int TangentSize() const = 0;
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool PlusJacobian(const double* x, double* jacobian) const = 0;
  // Extension (not in the reference): manifolds whose plus-Jacobian the evaluation
  // kernel can apply by itself describe themselves here (kind = CB200_MANIFOLD_* of
  // ceres_b200.h, param = kind-specific); the host then neither evaluates nor uploads
  // their PlusJacobian.  Returning false selects the generic path (PlusJacobian on the
  // host every Evaluate, as the reference does: parameter_block.h:312-338).
  virtual bool DeviceDescription(int* kind, int* param) const {
    (void)kind;
    (void)param;
    return false;
  }
};

namespace manifold_internal {
// Values of CB200_MANIFOLD_* (ceres_b200.h), repeated so this header stays stand-alone.
constexpr int kDeviceNone = 0, kDeviceSubset = 1, kDeviceQuaternionTail = 2,
              kDeviceEigenQuaternionTail = 3;
}  // namespace manifold_internal

class EuclideanManifoldBase : public Manifold {
 public:
  explicit EuclideanManifoldBase(int size) : size_(size) {}
  int AmbientSize() const override { return size_; }
  int TangentSize() const override { return size_; }
  bool Plus(const double* x, const double* delta, double* x_plus_delta) const override {
    for (int i = 0; i < size_; ++i) x_plus_delta[i] = x[i] + delta[i];
    return true;
  }
  bool PlusJacobian(const double*, double* jacobian) const override {
    for (int r = 0; r < size_; ++r)
      for (int c = 0; c < size_; ++c) jacobian[r * size_ + c] = (r == c) ? 1.0 : 0.0;
    return true;
  }
  bool DeviceDescription(int* kind, int* param) const override {
    *kind = manifold_internal::kDeviceNone;  // identity plus-Jacobian
    *param = 0;
    return true;
  }

 private:
  int size_;
};

template <int Size>
class EuclideanManifold : public EuclideanManifoldBase {
 public:
  EuclideanManifold() : EuclideanManifoldBase(Size) {}
  explicit EuclideanManifold(int size) : EuclideanManifoldBase(Size == -1 ? size : Size) {}
};

// Holds a subset of the coordinates constant.
class SubsetManifold : public Manifold {
 public:
  SubsetManifold(int size, const std::vector<int>& constant_parameters)
      : constancy_mask_(size, false), tangent_size_(size) {
    for (int i : constant_parameters) {
      if (!constancy_mask_[i]) --tangent_size_;
      constancy_mask_[i] = true;
    }
  }
  int AmbientSize() const override { return static_cast<int>(constancy_mask_.size()); }
  int TangentSize() const override { return tangent_size_; }
  bool Plus(const double* x, const double* delta, double* x_plus_delta) const override {
    for (int i = 0, j = 0; i < AmbientSize(); ++i)
      x_plus_delta[i] = constancy_mask_[i] ? x[i] : x[i] + delta[j++];
    return true;
  }
  bool PlusJacobian(const double*, double* jacobian) const override {
    if (tangent_size_ == 0) return true;
    const int n = AmbientSize();
    for (int i = 0; i < n * tangent_size_; ++i) jacobian[i] = 0.0;
    for (int r = 0, c = 0; r < n; ++r)
      if (!constancy_mask_[r]) jacobian[r * tangent_size_ + c++] = 1.0;
    return true;
  }
  const std::vector<bool>& constancy_mask() const { return constancy_mask_; }
  bool DeviceDescription(int* kind, int* param) const override {
    if (AmbientSize() > 31) return false;
    unsigned mask = 0;
    for (int i = 0; i < AmbientSize(); ++i)
      if (constancy_mask_[i]) mask |= 1u << i;
    *kind = manifold_internal::kDeviceSubset;
    *param = static_cast<int>(mask);
    return true;
  }

 private:
  std::vector<bool> constancy_mask_;
  int tangent_size_;
};

namespace manifold_internal {
// Order = position of (w, x, y, z) in storage.
template <int kW, int kX, int kY, int kZ>
inline void QuaternionPlus(const double* x, const double* delta, double* out) {
  const double norm_delta = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] +
                                      delta[2] * delta[2]);
  if (norm_delta == 0.0) {
    for (int i = 0; i < 4; ++i) out[i] = x[i];
    return;
  }
  const double k = std::sin(norm_delta) / norm_delta;
  double q[4];
  q[kW] = std::cos(norm_delta);
  q[kX] = k * delta[0];
  q[kY] = k * delta[1];
  q[kZ] = k * delta[2];
  // out = q * x
  out[kW] = q[kW] * x[kW] - q[kX] * x[kX] - q[kY] * x[kY] - q[kZ] * x[kZ];
  out[kX] = q[kW] * x[kX] + q[kX] * x[kW] + q[kY] * x[kZ] - q[kZ] * x[kY];
  out[kY] = q[kW] * x[kY] - q[kX] * x[kZ] + q[kY] * x[kW] + q[kZ] * x[kX];
  out[kZ] = q[kW] * x[kZ] + q[kX] * x[kY] - q[kY] * x[kX] + q[kZ] * x[kW];
}
template <int kW, int kX, int kY, int kZ>
inline void QuaternionPlusJacobian(const double* x, double* j /* 4x3 row-major */) {
  j[kW * 3 + 0] = -x[kX]; j[kW * 3 + 1] = -x[kY]; j[kW * 3 + 2] = -x[kZ];
  j[kX * 3 + 0] = x[kW];  j[kX * 3 + 1] = x[kZ];  j[kX * 3 + 2] = -x[kY];
  j[kY * 3 + 0] = -x[kZ]; j[kY * 3 + 1] = x[kW];  j[kY * 3 + 2] = x[kX];
  j[kZ * 3 + 0] = x[kY];  j[kZ * 3 + 1] = -x[kX]; j[kZ * 3 + 2] = x[kW];
}
}  // namespace manifold_internal

// Unit quaternions stored (w, x, y, z); Plus(x, delta) = exp(delta) * x.
class QuaternionManifold : public Manifold {
 public:
  int AmbientSize() const override { return 4; }
  int TangentSize() const override { return 3; }
  bool Plus(const double* x, const double* delta, double* out) const override {
    manifold_internal::QuaternionPlus<0, 1, 2, 3>(x, delta, out);
    return true;
  }
  bool PlusJacobian(const double* x, double* jacobian) const override {
    manifold_internal::QuaternionPlusJacobian<0, 1, 2, 3>(x, jacobian);
    return true;
  }
  bool DeviceDescription(int* kind, int* param) const override {
    *kind = manifold_internal::kDeviceQuaternionTail;
    *param = 0;
    return true;
  }
};

// Unit quaternions stored (x, y, z, w) as Eigen::Quaternion does.
class EigenQuaternionManifold : public Manifold {
 public:
  int AmbientSize() const override { return 4; }
  int TangentSize() const override { return 3; }
  bool Plus(const double* x, const double* delta, double* out) const override {
    manifold_internal::QuaternionPlus<3, 0, 1, 2>(x, delta, out);
    return true;
  }
  bool PlusJacobian(const double* x, double* jacobian) const override {
    manifold_internal::QuaternionPlusJacobian<3, 0, 1, 2>(x, jacobian);
    return true;
  }
  bool DeviceDescription(int* kind, int* param) const override {
    *kind = manifold_internal::kDeviceEigenQuaternionTail;
    *param = 0;
    return true;
  }
};

// Cartesian product; the plus-Jacobian is block diagonal.
template <typename... Ms>
class ProductManifold : public Manifold {
 public:
  ProductManifold() : manifolds_() { Init(); }
  explicit ProductManifold(Ms... ms) : manifolds_(std::move(ms)...) { Init(); }
  int AmbientSize() const override { return ambient_; }
  int TangentSize() const override { return tangent_; }
  // A quaternion followed by Euclidean factors is a shape the kernel knows.
  bool DeviceDescription(int* kind, int* param) const override {
    return DescribeQuaternionTail(kind, param, static_cast<std::tuple<Ms...>*>(nullptr));
  }
  bool Plus(const double* x, const double* delta, double* out) const override {
    bool ok = true;
    int a = 0, t = 0;
    std::apply(
        [&](const auto&... m) {
          ((ok = ok && m.Plus(x + a, delta + t, out + a), a += m.AmbientSize(),
            t += m.TangentSize()),
           ...);
        },
        manifolds_);
    return ok;
  }
  bool PlusJacobian(const double* x, double* jacobian) const override {
    for (int i = 0; i < ambient_ * tangent_; ++i) jacobian[i] = 0.0;
    bool ok = true;
    int a = 0, t = 0;
    std::vector<double> buffer;
    std::apply(
        [&](const auto&... m) {
          ((ok = ok && Place(m, x, jacobian, buffer, a, t)), ...);
        },
        manifolds_);
    return ok;
  }

 private:
  void Init() {
    ambient_ = tangent_ = 0;
    std::apply(
        [&](const auto&... m) {
          ((ambient_ += m.AmbientSize(), tangent_ += m.TangentSize()), ...);
        },
        manifolds_);
  }
  template <typename M>
  bool Place(const M& m, const double* x, double* jacobian, std::vector<double>& buffer,
             int& a, int& t) const {
    const int as = m.AmbientSize(), ts = m.TangentSize();
    buffer.assign(static_cast<size_t>(as) * ts + 1, 0.0);
    if (!m.PlusJacobian(x + a, buffer.data())) return false;
    for (int r = 0; r < as; ++r)
      for (int c = 0; c < ts; ++c) jacobian[(a + r) * tangent_ + (t + c)] = buffer[r * ts + c];
    a += as;
    t += ts;
    return true;
  }
  template <typename First, typename... Rest>
  static bool DescribeQuaternionTail(int* kind, int* param, std::tuple<First, Rest...>*) {
    constexpr bool tail_is_euclidean =
        (std::is_base_of<EuclideanManifoldBase, Rest>::value && ...);
    if (!tail_is_euclidean) return false;
    *param = 0;
    if (std::is_same<First, QuaternionManifold>::value) {
      *kind = manifold_internal::kDeviceQuaternionTail;
      return true;
    }
    if (std::is_same<First, EigenQuaternionManifold>::value) {
      *kind = manifold_internal::kDeviceEigenQuaternionTail;
      return true;
    }
    return false;
  }
  static bool DescribeQuaternionTail(int*, int*, std::tuple<>*) { return false; }
  std::tuple<Ms...> manifolds_;
  int ambient_ = 0, tangent_ = 0;
};

}  // namespace ceres

#endif  // CERES_B200_MANIFOLD_H_
