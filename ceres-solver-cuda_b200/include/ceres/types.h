// Enums of the public API that the evaluation path reads
// (reference: include/ceres/types.h; only the members the path dispatches on).
#ifndef CERES_B200_TYPES_H_
#define CERES_B200_TYPES_H_

namespace ceres {

// Argument type used in interfaces that can optionally take ownership of a
// passed in argument.
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };

// Selects the Jacobian container (internal/ceres/evaluator.cc:57-92).
enum LinearSolverType {
  DENSE_NORMAL_CHOLESKY,
  DENSE_QR,
  SPARSE_NORMAL_CHOLESKY,
  DENSE_SCHUR,
  SPARSE_SCHUR,
  ITERATIVE_SCHUR,
  CGNR
};

enum SparseLinearAlgebraLibraryType {
  SUITE_SPARSE,
  EIGEN_SPARSE,
  ACCELERATE_SPARSE,
  CUDA_SPARSE,
  NO_SPARSE
};

enum MinimizerType { LINE_SEARCH, TRUST_REGION };

constexpr int DYNAMIC = -1;

}  // namespace ceres

#endif  // CERES_B200_TYPES_H_
