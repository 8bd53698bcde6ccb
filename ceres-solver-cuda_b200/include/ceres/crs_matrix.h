// A compressed-row matrix handed to the user by Problem::Evaluate
// (reference: include/ceres/crs_matrix.h): row i holds cols[rows[i] .. rows[i + 1]) with
// the matching values.
#ifndef CERES_B200_CRS_MATRIX_H_
#define CERES_B200_CRS_MATRIX_H_

#include <vector>

namespace ceres {

struct CRSMatrix {
  int num_rows = 0;
  int num_cols = 0;
  std::vector<int> cols;
  std::vector<int> rows;
  std::vector<double> values;
};

}  // namespace ceres

#endif  // CERES_B200_CRS_MATRIX_H_
