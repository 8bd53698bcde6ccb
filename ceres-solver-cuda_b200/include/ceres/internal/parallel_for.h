// Static-partition parallel loop for the host-side vector work around the evaluator
// (Program::Plus, norms and scalings in the trust-region loop); the reference runs the same
// loops through its ParallelFor with Solver::Options::num_threads
// (internal/ceres/program.cc:121-150, internal/ceres/parallel_vector_ops.h).
#ifndef CERES_B200_INTERNAL_PARALLEL_FOR_H_
#define CERES_B200_INTERNAL_PARALLEL_FOR_H_

#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace ceres {
namespace internal {

// Calls f(begin, end, thread_index) on num_threads contiguous slices of [0, n).
template <typename F>
void ParallelFor(int num_threads, int64_t n, F&& f) {
  const int threads =
      static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(num_threads, n / 4096)));
  if (threads == 1) {
    f(int64_t{0}, n, 0);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(threads - 1);
  for (int t = 1; t < threads; ++t)
    pool.emplace_back([&, t] { f(n * t / threads, n * (t + 1) / threads, t); });
  f(int64_t{0}, n / threads, 0);
  for (std::thread& th : pool) th.join();
}

}  // namespace internal
}  // namespace ceres

#endif  // CERES_B200_INTERNAL_PARALLEL_FOR_H_
