// oracle/orc_trust_region.h — TEST INFRASTRUCTURE ONLY.
//
// Restatement of ceres::internal::TrustRegionMinimizer + LevenbergMarquardtStrategy
// (Ceres Solver 2.1: internal/ceres/trust_region_minimizer.cc,
// levenberg_marquardt_strategy.cc, trust_region_step_evaluator.cc).  Ceres is a
// non-vendored dependency of the reference (README.md:34); the reference reaches it at
// bundle_optimizer.h:224 and keypoint_optimizer.h:92 (`ceres::Solve`).  The algorithm is
// restated from Ceres' published source/documentation; parity of end results is
// therefore "unpinned" (no Ceres in this container).
#pragma once
#include <cmath>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

namespace orc {

struct TROptions {
  int max_num_iterations = 50;
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  double min_relative_decrease = 1e-3;
  double initial_trust_region_radius = 1e4;
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;
  bool jacobi_scaling = true;
  int max_num_consecutive_invalid_steps = 5;
  bool use_inner_iterations = false;
  double inner_iteration_tolerance = 1e-3;
  int max_num_line_search_step_size_iterations = 20;
  bool use_nonmonotonic_steps = false;            // ceres::Solver::Options, default false; the reference's default.yaml sets true
  int max_consecutive_nonmonotonic_steps = 5;
  bool verbose = false;
};

// ceres::internal::TrustRegionStepEvaluator (trust_region_step_evaluator.cc; Conn, Gould & Toint, algorithm 10.1.2).
// max_consecutive_nonmonotonic_steps == 0 is the monotonic minimizer: the reference iterate is reset on every accepted step.
struct StepEvaluator {
  int max_nonmonotonic;
  double minimum_cost, current_cost, reference_cost, candidate_cost;
  double acc_reference_model_cost_change = 0.0, acc_candidate_model_cost_change = 0.0;
  int num_consecutive_nonmonotonic_steps = 0;
  StepEvaluator(double initial_cost, int max_nm)
      : max_nonmonotonic(max_nm), minimum_cost(initial_cost), current_cost(initial_cost), reference_cost(initial_cost),
        candidate_cost(initial_cost) {}
  double StepQuality(double cost, double model_cost_change) const {
    if (cost >= std::numeric_limits<double>::max()) return std::numeric_limits<double>::lowest();
    const double relative_decrease = (current_cost - cost) / model_cost_change;
    const double historical_relative_decrease = (reference_cost - cost) / (acc_reference_model_cost_change + model_cost_change);
    return std::max(relative_decrease, historical_relative_decrease);
  }
  void StepAccepted(double cost, double model_cost_change) {
    current_cost = cost;
    acc_candidate_model_cost_change += model_cost_change;
    acc_reference_model_cost_change += model_cost_change;
    if (current_cost < minimum_cost) {
      minimum_cost = current_cost;
      num_consecutive_nonmonotonic_steps = 0;
      candidate_cost = current_cost;
      acc_candidate_model_cost_change = 0.0;
    } else {
      ++num_consecutive_nonmonotonic_steps;
      if (current_cost > candidate_cost) { candidate_cost = current_cost; acc_candidate_model_cost_change = 0.0; }
    }
    if (num_consecutive_nonmonotonic_steps == max_nonmonotonic) {
      reference_cost = candidate_cost;
      acc_reference_model_cost_change = acc_candidate_model_cost_change;
    }
  }
};

struct TRIteration {
  int iteration = 0;
  bool step_is_valid = false, step_is_successful = false;
  double cost = 0, cost_change = 0, gradient_max_norm = 0, step_norm = 0, relative_decrease = 0,
         trust_region_radius = 0;
  int linear_solver_iterations = 0;
};

struct TRSummary {
  double initial_cost = 0, final_cost = 0;
  int num_successful_steps = 0, num_unsuccessful_steps = 0, num_inner_iteration_steps = 0;
  int termination_type = 1;  // 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE
  std::string message;
  std::vector<TRIteration> iterations;
};

// What the minimizer needs from a problem (ceres Evaluator + LinearSolver rolled together).
struct TREvaluator {
  virtual ~TREvaluator() {}
  virtual int NumParameters() const = 0;  // ambient size
  virtual int NumLocal() const = 0;       // tangent size
  // cost (0.5 * sum rho) at x; with_jacobian also linearises (residuals, J, gradient kept inside)
  virtual bool Evaluate(const double* x, double* cost, bool with_jacobian) = 0;
  virtual void Gradient(double* g) const = 0;           // J^T r, local
  virtual void SquaredColumnNorm(double* d) const = 0;  // diag(J^T J), local, unscaled
  // (J^T J + diag(D2)) delta = -g ; returns false on linear solver failure
  virtual bool SolveDamped(const double* D2, double* delta, int* iters) = 0;
  // -(J delta)^T (r + J delta / 2)
  virtual double ModelCostChange(const double* delta) const = 0;
  virtual void Plus(const double* x, const double* delta, double* x_plus) const = 0;
  virtual bool IsConstrained() const { return false; }
  // CoordinateDescentMinimizer::Minimize on the candidate (only called when enabled)
  virtual void InnerIterations(double* /*x*/) {}
};

namespace detail {
inline double Norm(const std::vector<double>& v) { double s = 0; for (double e : v) s += e * e; return std::sqrt(s); }
inline double NormDiff(const std::vector<double>& a, const std::vector<double>& b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) { const double d = a[i] - b[i]; s += d * d; } return std::sqrt(s); }
inline double MaxDiff(const std::vector<double>& a, const std::vector<double>& b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s = std::max(s, std::fabs(a[i] - b[i])); return s; }

// (ceres) MinimizeInterpolatingPolynomial for the two/three-sample CUBIC case used by
// ArmijoLineSearch (internal/ceres/polynomial.cc). Fits the polynomial through
// {value,gradient} samples by solving the Vandermonde system and minimises it on
// [xmin,xmax] over endpoints and real critical points.
struct FSample { double x, value, gradient; bool value_valid, gradient_valid; };

inline bool SolveDense(int n, std::vector<double>& A, std::vector<double>& b) {
  // Gaussian elimination with partial pivoting (A row-major n x n)
  for (int k = 0; k < n; ++k) {
    int p = k; double m = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i) if (std::fabs(A[i * n + k]) > m) { m = std::fabs(A[i * n + k]); p = i; }
    if (m == 0.0) return false;
    if (p != k) { for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]); std::swap(b[k], b[p]); }
    for (int i = k + 1; i < n; ++i) {
      const double f = A[i * n + k] / A[k * n + k];
      for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
      b[i] -= f * b[k];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * b[j];
    b[i] = s / A[i * n + i];
  }
  return true;
}

inline double PolyEval(const std::vector<double>& p, double x) {  // highest degree first
  double v = 0; for (double c : p) v = v * x + c; return v;
}

// Real roots of a polynomial (highest degree first) of degree <= 4 via bracketing +
// bisection on the derivative sign changes; adequate for the line-search use.
inline void RealRootsInInterval(const std::vector<double>& p, double lo, double hi, std::vector<double>* roots) {
  const int kGrid = 512;
  double xp = lo, fp = PolyEval(p, lo);
  for (int i = 1; i <= kGrid; ++i) {
    const double x = lo + (hi - lo) * i / kGrid, f = PolyEval(p, x);
    if (fp == 0.0) roots->push_back(xp);
    else if ((fp < 0) != (f < 0) && f != 0.0) {
      double a = xp, b = x, fa = fp;
      for (int it = 0; it < 200; ++it) {
        const double m = 0.5 * (a + b), fm = PolyEval(p, m);
        if ((fa < 0) != (fm < 0)) { b = m; } else { a = m; fa = fm; }
      }
      roots->push_back(0.5 * (a + b));
    }
    xp = x; fp = f;
  }
  if (fp == 0.0) roots->push_back(xp);
}

inline double MinimizeInterpolatingPolynomial(const std::vector<FSample>& samples, double xmin, double xmax) {
  int n = 0;
  for (auto& s : samples) { if (s.value_valid) ++n; if (s.gradient_valid) ++n; }
  std::vector<double> A(n * n, 0.0), b(n, 0.0);
  int row = 0;
  const int degree = n - 1;
  for (auto& s : samples) {
    if (s.value_valid) {
      for (int j = 0; j <= degree; ++j) A[row * n + j] = std::pow(s.x, degree - j);
      b[row++] = s.value;
    }
    if (s.gradient_valid) {
      for (int j = 0; j < degree; ++j) A[row * n + j] = (degree - j) * std::pow(s.x, degree - j - 1);
      b[row++] = s.gradient;
    }
  }
  double best_x = (xmin + xmax) / 2.0;
  if (!SolveDense(n, A, b)) return best_x;
  const std::vector<double>& poly = b;
  double best = std::numeric_limits<double>::max();
  auto consider = [&](double x) { const double v = PolyEval(poly, x); if (v < best) { best = v; best_x = x; } };
  consider(xmin); consider(xmax);
  std::vector<double> d(degree);
  for (int j = 0; j < degree; ++j) d[j] = (degree - j) * poly[j];
  std::vector<double> roots;
  RealRootsInInterval(d, xmin, xmax, &roots);
  for (double r : roots) consider(r);
  return best_x;
}
}  // namespace detail

class TrustRegionMinimizer {
 public:
  TROptions opt;
  explicit TrustRegionMinimizer(const TROptions& o) : opt(o) {}

  void Minimize(TREvaluator* ev, double* parameters, TRSummary* sum) {
    const int n = ev->NumParameters(), nl = ev->NumLocal();
    std::vector<double> x(parameters, parameters + n), candidate_x(n), inner_x(n);
    std::vector<double> gradient(nl), scale(nl, 1.0), diag(nl), D2(nl), step(nl), delta(nl);
    std::vector<double> neg_g(nl), proj(n);
    double x_cost = 0, candidate_cost = 0, model_cost_change = 0;
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    bool inner_enabled = opt.use_inner_iterations;
    int num_consecutive_invalid_steps = 0;
    sum->iterations.clear();
    sum->termination_type = 1;

    auto gradient_norms = [&](TRIteration* it) {
      for (int i = 0; i < nl; ++i) neg_g[i] = -gradient[i];
      ev->Plus(x.data(), neg_g.data(), proj.data());
      it->gradient_max_norm = detail::MaxDiff(x, proj);
    };

    // ---- IterationZero
    TRIteration it;
    double x_norm = detail::Norm(x);
    if (ev->IsConstrained()) {
      std::fill(delta.begin(), delta.end(), 0.0);
      ev->Plus(x.data(), delta.data(), candidate_x.data());
      x = candidate_x;
      x_norm = detail::Norm(x);
    }
    if (!ev->Evaluate(x.data(), &x_cost, true)) { sum->termination_type = 2; sum->message = "initial evaluation failed"; return; }
    ev->Gradient(gradient.data());
    if (opt.jacobi_scaling) {
      ev->SquaredColumnNorm(scale.data());
      for (int i = 0; i < nl; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(scale[i]));
    }
    it.cost = x_cost;
    gradient_norms(&it);
    it.trust_region_radius = radius;
    sum->initial_cost = x_cost;
    double minimum_cost = x_cost;
    std::vector<double> best_x = x;
    StepEvaluator step_evaluator(x_cost, opt.use_nonmonotonic_steps ? opt.max_consecutive_nonmonotonic_steps : 0);

    auto finalize = [&]() -> bool {  // FinalizeIterationAndCheckIfMinimizerCanContinue
      if (it.step_is_successful) {
        ++sum->num_successful_steps;
        if (x_cost < minimum_cost) { minimum_cost = x_cost; best_x = x; }
      } else if (it.iteration > 0) {
        ++sum->num_unsuccessful_steps;
      }
      it.trust_region_radius = radius;
      sum->iterations.push_back(it);
      if (opt.verbose)
        fprintf(stderr, "[orc] it %3d cost %.12e change %.3e |g| %.3e step %.3e rho %.3e radius %.3e %s\n",
                it.iteration, it.cost, it.cost_change, it.gradient_max_norm, it.step_norm,
                it.relative_decrease, it.trust_region_radius, it.step_is_successful ? "ok" : "--");
      if (it.iteration >= opt.max_num_iterations) { sum->termination_type = 1; sum->message = "Maximum number of iterations reached."; return false; }
      if (it.gradient_max_norm <= opt.gradient_tolerance) { sum->termination_type = 0; sum->message = "Gradient tolerance reached."; return false; }
      if (radius < opt.min_trust_region_radius) { sum->termination_type = 0; sum->message = "Minimum trust region radius reached."; return false; }
      return true;
    };

    while (finalize()) {
      const double previous_gradient_max_norm = it.gradient_max_norm;
      const int iteration = it.iteration + 1;
      it = TRIteration();
      it.iteration = iteration;

      // ---- ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep)
      ev->SquaredColumnNorm(diag.data());
      for (int i = 0; i < nl; ++i) {
        const double ds = diag[i] * scale[i] * scale[i];  // column norm of the scaled Jacobian
        const double c = std::min(std::max(ds, opt.min_lm_diagonal), opt.max_lm_diagonal);
        D2[i] = c / (radius * scale[i] * scale[i]);       // back in unscaled variables
      }
      int lin_iters = 0;
      bool solved = ev->SolveDamped(D2.data(), delta.data(), &lin_iters);
      it.linear_solver_iterations = lin_iters;
      if (solved) {
        for (int i = 0; i < nl; ++i) if (!std::isfinite(delta[i])) { solved = false; break; }
      }
      if (solved) {
        model_cost_change = ev->ModelCostChange(delta.data());
        it.step_is_valid = model_cost_change > 0.0;
      }
      if (!it.step_is_valid) {
        // ---- HandleInvalidStep
        ++num_consecutive_invalid_steps;
        if (num_consecutive_invalid_steps >= opt.max_num_consecutive_invalid_steps) {
          sum->termination_type = 2;
          sum->message = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps";
          break;
        }
        radius = radius / decrease_factor;  // StepIsInvalid == StepRejected
        decrease_factor *= 2.0;
        it.cost = x_cost; it.cost_change = 0; it.gradient_max_norm = previous_gradient_max_norm;
        it.step_norm = 0; it.relative_decrease = 0;
        continue;
      }
      num_consecutive_invalid_steps = 0;

      if (ev->IsConstrained() && opt.max_num_line_search_step_size_iterations > 0)
        DoLineSearch(ev, x, gradient, x_cost, &delta);

      // ---- ComputeCandidatePointAndEvaluateCost
      ev->Plus(x.data(), delta.data(), candidate_x.data());
      if (!ev->Evaluate(candidate_x.data(), &candidate_cost, false) || !std::isfinite(candidate_cost))
        candidate_cost = std::numeric_limits<double>::max();

      // ---- DoInnerIterationsIfNeeded
      bool inner_were_useful = false;
      if (inner_enabled && candidate_cost < std::numeric_limits<double>::max()) {
        ++sum->num_inner_iteration_steps;
        inner_x = candidate_x;
        ev->InnerIterations(inner_x.data());
        double inner_cost;
        if (ev->Evaluate(inner_x.data(), &inner_cost, false)) {
          candidate_x = inner_x;
          const double inner_cost_change = candidate_cost - inner_cost;
          model_cost_change += inner_cost_change;
          inner_were_useful = inner_cost < x_cost;
          const double rel = 1.0 - inner_cost / candidate_cost;
          inner_enabled = rel > opt.inner_iteration_tolerance;
          candidate_cost = inner_cost;
        }
      }

      // ---- ParameterToleranceReached
      it.step_norm = detail::NormDiff(x, candidate_x);
      if (it.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
        sum->termination_type = 0; sum->message = "Parameter tolerance reached.";
        break;
      }
      // ---- FunctionToleranceReached
      it.cost_change = x_cost - candidate_cost;
      if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) {
        sum->termination_type = 0; sum->message = "Function tolerance reached.";
        break;
      }
      // ---- IsStepSuccessful
      it.relative_decrease = step_evaluator.StepQuality(candidate_cost, model_cost_change);
      const bool ok = inner_were_useful || it.relative_decrease > opt.min_relative_decrease;
      if (ok) {
        // ---- HandleSuccessfulStep
        x = candidate_x;
        x_norm = detail::Norm(x);
        if (!ev->Evaluate(x.data(), &x_cost, true)) { sum->termination_type = 2; sum->message = "evaluation failed"; break; }
        ev->Gradient(gradient.data());
        it.cost = x_cost;
        gradient_norms(&it);
        it.step_is_successful = true;
        // LevenbergMarquardtStrategy::StepAccepted
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
        radius = std::min(opt.max_trust_region_radius, radius);
        decrease_factor = 2.0;
        step_evaluator.StepAccepted(candidate_cost, model_cost_change);
      } else {
        it.step_is_successful = false;
        it.cost = candidate_cost;
        it.gradient_max_norm = previous_gradient_max_norm;
        radius = radius / decrease_factor;  // StepRejected
        decrease_factor *= 2.0;
      }
    }
    // ceres keeps the lowest-cost successful iterate in `parameters`
    if (x_cost < minimum_cost) { minimum_cost = x_cost; best_x = x; }
    for (int i = 0; i < n; ++i) parameters[i] = best_x[i];
    sum->final_cost = minimum_cost;
  }

 private:
  // TrustRegionMinimizer::DoLineSearch + ArmijoLineSearch::DoSearch (CUBIC interpolation)
  void DoLineSearch(TREvaluator* ev, const std::vector<double>& x, const std::vector<double>& gradient,
                    double cost, std::vector<double>* delta) {
    const int n = ev->NumParameters(), nl = ev->NumLocal();
    const double sufficient_decrease = 1e-4, max_step_contraction = 1e-3, min_step_contraction = 0.6;
    const double min_step_size = 1e-9;
    double initial_gradient = 0, dir_max = 0;
    for (int i = 0; i < nl; ++i) { initial_gradient += gradient[i] * (*delta)[i]; dir_max = std::max(dir_max, std::fabs((*delta)[i])); }
    std::vector<double> xp(n), d(nl), g(nl);
    auto eval = [&](double a, detail::FSample* s) {
      for (int i = 0; i < nl; ++i) d[i] = a * (*delta)[i];
      ev->Plus(x.data(), d.data(), xp.data());
      double c;
      s->x = a;
      s->value_valid = ev->Evaluate(xp.data(), &c, true) && std::isfinite(c);
      s->value = c;
      if (s->value_valid) {
        ev->Gradient(g.data());
        double gg = 0; for (int i = 0; i < nl; ++i) gg += g[i] * (*delta)[i];
        s->gradient = gg; s->gradient_valid = std::isfinite(gg);
      } else s->gradient_valid = false;
    };
    detail::FSample initial{0.0, cost, initial_gradient, true, true};
    detail::FSample previous{0, 0, 0, false, false}, current;
    eval(1.0, &current);
    int iters = 0;
    bool success = true;
    while (!current.value_valid || current.value > cost + sufficient_decrease * initial_gradient * current.x) {
      ++iters;
      if (iters >= opt.max_num_line_search_step_size_iterations) { success = false; break; }
      std::vector<detail::FSample> samples;
      samples.push_back(initial);
      if (current.value_valid) samples.push_back(current);
      if (previous.value_valid) samples.push_back(previous);
      double step_size;
      if (!current.value_valid) step_size = 0.5 * (max_step_contraction * current.x + min_step_contraction * current.x);
      else step_size = detail::MinimizeInterpolatingPolynomial(samples, max_step_contraction * current.x, min_step_contraction * current.x);
      if (step_size * dir_max < min_step_size) { success = false; break; }
      previous = current;
      eval(step_size, &current);
    }
    if (success) for (int i = 0; i < nl; ++i) (*delta)[i] *= current.x;
    // restore the linearisation at x (the evaluator keeps the last Jacobian it computed)
    double c0; ev->Evaluate(x.data(), &c0, true);
  }
};

}  // namespace orc
