// Jacobian containers with the reference's value layouts.
//
// BlockSparseMatrix: values laid out cell by cell, each cell row-major
// row_block_size x col_block_size, cells of the first num_eliminate_blocks column
// blocks (E) before all others (F) — internal/ceres/block_sparse_matrix.h:163-167,
// block_structure.h:52-90, block_jacobian_writer.cc:62-150,192-250.
// CompressedRowSparseMatrix: rows[num_rows + 1], cols[nnz], values allocated
// nnz + num_cols so the LM diagonal can be appended without reallocation —
// compressed_row_jacobian_writer.cc:93-193.
//
// Values live in pinned host memory (cb200_host_alloc) so the device->host copy of
// the Jacobian, the end-to-end bottleneck the reference names (README.md:198-200),
// runs at full PCIe rate.  The block structure is stored flat (CSR over cells)
// instead of one std::vector<Cell> per row.
#ifndef CERES_B200_INTERNAL_SPARSE_MATRIX_H_
#define CERES_B200_INTERNAL_SPARSE_MATRIX_H_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ceres_b200.h"

namespace ceres {
namespace internal {

struct Block {
  int32_t size = -1;
  int32_t position = -1;  // position along the row or column
};

struct Cell {
  int32_t block_id = -1;  // column block
  int32_t position = -1;  // where in the values array the cell starts
};

struct CompressedRowBlockStructure {
  std::vector<Block> cols;
  std::vector<Block> rows;              // row blocks
  std::vector<int32_t> row_cell_begin;  // size rows.size() + 1
  std::vector<Cell> cells;              // per row sorted by block_id
};

class SparseMatrix {
 public:
  virtual ~SparseMatrix() {
    if (values_) cb200_host_free(values_);
  }
  SparseMatrix(const SparseMatrix&) = delete;
  void operator=(const SparseMatrix&) = delete;
  double* mutable_values() { return values_; }
  const double* values() const { return values_; }
  int num_rows() const { return num_rows_; }
  int num_cols() const { return num_cols_; }
  int64_t num_nonzeros() const { return num_nonzeros_; }
  int64_t values_size() const { return values_size_; }
  void SetZero() { if (values_) std::memset(values_, 0, sizeof(double) * values_size_); }
  // Row-major num_rows x num_cols.
  virtual void ToDenseMatrix(std::vector<double>* dense) const = 0;
  // The operations the trust-region loop needs from a Jacobian
  // (internal/ceres/sparse_matrix.h): y += A x, y += A' x, squared column norms,
  // A <- A diag(scale).
  virtual void RightMultiplyAndAccumulate(const double* x, double* y) const = 0;
  virtual void LeftMultiplyAndAccumulate(const double* x, double* y) const = 0;
  virtual void SquaredColumnNorm(double* x) const = 0;
  virtual void ScaleColumns(const double* scale) = 0;

 protected:
  SparseMatrix() {}
  void Allocate(int64_t values_size) {
    values_size_ = values_size;
    values_ = static_cast<double*>(cb200_host_alloc(sizeof(double) * (values_size + 1)));
  }
  double* values_ = nullptr;
  int64_t values_size_ = 0;
  int64_t num_nonzeros_ = 0;
  int num_rows_ = 0, num_cols_ = 0;
};

class BlockSparseMatrix final : public SparseMatrix {
 public:
  BlockSparseMatrix(CompressedRowBlockStructure* block_structure, int64_t num_nonzeros)
      : block_structure_(block_structure) {
    num_nonzeros_ = num_nonzeros;
    for (const Block& b : block_structure_->rows) num_rows_ += b.size;
    for (const Block& b : block_structure_->cols) num_cols_ += b.size;
    Allocate(num_nonzeros);
  }
  ~BlockSparseMatrix() override { delete block_structure_; }
  const CompressedRowBlockStructure* block_structure() const { return block_structure_; }
  void ToDenseMatrix(std::vector<double>* dense) const override {
    dense->assign(static_cast<size_t>(num_rows_) * num_cols_, 0.0);
    const auto& bs = *block_structure_;
    for (size_t i = 0; i < bs.rows.size(); ++i) {
      for (int32_t c = bs.row_cell_begin[i]; c < bs.row_cell_begin[i + 1]; ++c) {
        const Block& col = bs.cols[bs.cells[c].block_id];
        const double* v = values_ + bs.cells[c].position;
        for (int r = 0; r < bs.rows[i].size; ++r)
          for (int k = 0; k < col.size; ++k)
            (*dense)[static_cast<size_t>(bs.rows[i].position + r) * num_cols_ + col.position + k] +=
                v[r * col.size + k];
      }
    }
  }

  template <typename F>  // f(row, col, value&) over every stored entry
  void ForEachEntry(F&& f) const {
    const auto& bs = *block_structure_;
    for (size_t i = 0; i < bs.rows.size(); ++i)
      for (int32_t c = bs.row_cell_begin[i]; c < bs.row_cell_begin[i + 1]; ++c) {
        const Block& col = bs.cols[bs.cells[c].block_id];
        double* v = values_ + bs.cells[c].position;
        for (int r = 0; r < bs.rows[i].size; ++r)
          for (int k = 0; k < col.size; ++k)
            f(bs.rows[i].position + r, col.position + k, v[r * col.size + k]);
      }
  }
  void RightMultiplyAndAccumulate(const double* x, double* y) const override {
    ForEachEntry([&](int r, int c, double& v) { y[r] += v * x[c]; });
  }
  void LeftMultiplyAndAccumulate(const double* x, double* y) const override {
    ForEachEntry([&](int r, int c, double& v) { y[c] += v * x[r]; });
  }
  void SquaredColumnNorm(double* x) const override {
    std::memset(x, 0, sizeof(double) * num_cols_);
    ForEachEntry([&](int, int c, double& v) { x[c] += v * v; });
  }
  void ScaleColumns(const double* scale) override {
    ForEachEntry([&](int, int c, double& v) { v *= scale[c]; });
  }

 private:
  CompressedRowBlockStructure* block_structure_;
};

class CompressedRowSparseMatrix final : public SparseMatrix {
 public:
  CompressedRowSparseMatrix(int num_rows, int num_cols, int64_t max_num_nonzeros)
      : rows_(num_rows + 1, 0), cols_(max_num_nonzeros, 0) {
    num_rows_ = num_rows;
    num_cols_ = num_cols;
    Allocate(max_num_nonzeros);
  }
  int* mutable_rows() { return rows_.data(); }
  int* mutable_cols() { return cols_.data(); }
  const int* rows() const { return rows_.data(); }
  const int* cols() const { return cols_.data(); }
  void set_num_nonzeros(int64_t n) { num_nonzeros_ = n; }
  std::vector<Block>* mutable_row_blocks() { return &row_blocks_; }
  std::vector<Block>* mutable_col_blocks() { return &col_blocks_; }
  const std::vector<Block>& row_blocks() const { return row_blocks_; }
  const std::vector<Block>& col_blocks() const { return col_blocks_; }
  void ToDenseMatrix(std::vector<double>* dense) const override {
    dense->assign(static_cast<size_t>(num_rows_) * num_cols_, 0.0);
    for (int r = 0; r < num_rows_; ++r)
      for (int k = rows_[r]; k < rows_[r + 1]; ++k)
        (*dense)[static_cast<size_t>(r) * num_cols_ + cols_[k]] += values_[k];
  }

  void RightMultiplyAndAccumulate(const double* x, double* y) const override {
    for (int r = 0; r < num_rows_; ++r) {
      double acc = 0.0;
      for (int k = rows_[r]; k < rows_[r + 1]; ++k) acc += values_[k] * x[cols_[k]];
      y[r] += acc;
    }
  }
  void LeftMultiplyAndAccumulate(const double* x, double* y) const override {
    for (int r = 0; r < num_rows_; ++r)
      for (int k = rows_[r]; k < rows_[r + 1]; ++k) y[cols_[k]] += values_[k] * x[r];
  }
  void SquaredColumnNorm(double* x) const override {
    std::memset(x, 0, sizeof(double) * num_cols_);
    for (int64_t k = 0; k < num_nonzeros_; ++k) x[cols_[k]] += values_[k] * values_[k];
  }
  void ScaleColumns(const double* scale) override {
    for (int64_t k = 0; k < num_nonzeros_; ++k) values_[k] *= scale[cols_[k]];
  }

 private:
  std::vector<int> rows_, cols_;
  std::vector<Block> row_blocks_, col_blocks_;
};

// A Jacobian whose values never leave the device: the
// evaluator leaves them in HBM (CB200_KEEP_JACOBIAN_ON_DEVICE) and every operation the
// trust-region loop needs runs there through the C ABI.  Counterpart of the reference's
// CudaSparseMatrix (internal/ceres/cuda_sparse_matrix.h) for the evaluator's own layout,
// so no conversion and no host copy of the values exists.  values() is null.
class DeviceResidentJacobian final : public SparseMatrix {
 public:
  DeviceResidentJacobian(cb200_engine* engine, int num_rows, int num_cols, int64_t num_nonzeros)
      : engine_(engine) {
    num_rows_ = num_rows;
    num_cols_ = num_cols;
    num_nonzeros_ = num_nonzeros;
  }
  cb200_engine* engine() const { return engine_; }
  void RightMultiplyAndAccumulate(const double* x, double* y) const override {
    std::vector<double> t(num_rows_);
    Check(cb200_engine_jacobian_multiply(engine_, 0, x, t.data()));
    for (int i = 0; i < num_rows_; ++i) y[i] += t[i];
  }
  void LeftMultiplyAndAccumulate(const double* x, double* y) const override {
    std::vector<double> t(num_cols_);
    Check(cb200_engine_jacobian_multiply(engine_, 1, x, t.data()));
    for (int i = 0; i < num_cols_; ++i) y[i] += t[i];
  }
  void SquaredColumnNorm(double* x) const override {
    Check(cb200_engine_jacobian_squared_column_norm(engine_, x));
  }
  void ScaleColumns(const double* scale) override {
    Check(cb200_engine_jacobian_scale_columns(engine_, scale));
  }
  // Column by column through J e_c; meant for tests on small problems.
  void ToDenseMatrix(std::vector<double>* dense) const override {
    dense->assign(static_cast<size_t>(num_rows_) * num_cols_, 0.0);
    std::vector<double> e(num_cols_, 0.0), col(num_rows_);
    for (int c = 0; c < num_cols_; ++c) {
      e[c] = 1.0;
      Check(cb200_engine_jacobian_multiply(engine_, 0, e.data(), col.data()));
      e[c] = 0.0;
      for (int r = 0; r < num_rows_; ++r) (*dense)[static_cast<size_t>(r) * num_cols_ + c] = col[r];
    }
  }

 private:
  void Check(int rc) const {
    if (rc == CB200_OK) return;
    std::fprintf(stderr, "device Jacobian operation failed: %s\n", cb200_engine_last_error(engine_));
    std::abort();  // like the reference's CHECK on CUDA failures
  }
  cb200_engine* engine_;
};

}  // namespace internal
}  // namespace ceres

#endif  // CERES_B200_INTERNAL_SPARSE_MATRIX_H_
