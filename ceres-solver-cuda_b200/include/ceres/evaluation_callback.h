// ceres::EvaluationCallback, same contract as the reference's
// include/ceres/evaluation_callback.h:63-78: Problem::Options::evaluation_callback
// (include/ceres/problem.h:183) is notified before every evaluation, after the user's
// parameter blocks were set to the evaluation point (program_evaluator_cuda.h:116-121).
#ifndef CERES_B200_EVALUATION_CALLBACK_H_
#define CERES_B200_EVALUATION_CALLBACK_H_

namespace ceres {

class EvaluationCallback {
 public:
  virtual ~EvaluationCallback() = default;
  // User parameters (the double* values given to the problem) hold the evaluation point and
  // stay fixed until the next call.  new_evaluation_point == false: the same point as the
  // previous evaluation (of residuals or Jacobians), cached results may be reused.
  virtual void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) = 0;
};

}  // namespace ceres

#endif  // CERES_B200_EVALUATION_CALLBACK_H_
