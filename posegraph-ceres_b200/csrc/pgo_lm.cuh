// pgo_lm.cuh -- the Levenberg-Marquardt bookkeeping of ceres::Solve as DEVICE-resident state.
//
// ceres::internal::TrustRegionMinimizer::Minimize + LevenbergMarquardtStrategy (ceres 1.13 flow, the reference's
// settings: REF/test/pose_graph_ceres_plus_finial.cpp:531-544) take one decision per iteration from a dozen scalars:
// accept / reject, the new trust-region radius, the termination tests.  lm_decide() below is that decision, written
// once as a __host__ __device__ function (the same sequence of tests as oracle/pgo_oracle.c:oracle_solve).  On the
// device it runs as a one-thread kernel after the speculative linearisation at the candidate, followed by a commit
// kernel (candidate poses and the candidate's H, g become current when the step was accepted).  The host therefore never
// waits for an iteration: it enqueues iterations ahead of the GPU and only watches a `done` flag two iterations behind.
#pragma once

#include "../../include/pgo_b200.h"
#include "pgo_common.cuh"

namespace pgo {

struct LmOptions {
  int max_num_iterations;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_radius, max_radius, min_radius, min_relative_decrease;
  int max_consecutive_invalid;
  int verbose;
};

enum LmReason {
  kLmRunning = 0, kLmMaxIterations, kLmGradientTolerance, kLmMinRadius, kLmInvalidSteps, kLmParameterTolerance,
  kLmFunctionTolerance, kLmInitialCostNotFinite
};

struct LmState {
  double radius, decrease_factor, x_cost, x_norm, gradient_max_norm, gradient_norm;
  double initial_cost;
  double reason_value;            // the number the termination message quotes
  double solver_ns, linearize_ns; // %globaltimer spans: linear solver, (plus + linearize + norms)
  unsigned long long t_mark;      // end of the previous iteration's commit
  unsigned long long t_solved;    // start of the plus kernel (= the linear solver is done)
  long long total_pcg_iterations;
  int iter;                       // iteration being worked on (1-based once the loop runs)
  int reuse_diagonal;
  int num_consecutive_invalid;
  int done;                       // every kernel of the loop returns at once when set
  int reason, termination_type;
  int accept;                     // last decision: the candidate becomes the current point
  int num_successful, num_unsuccessful;
  int num_rows;                   // iteration-log rows produced (incl. iteration 0)
};

__host__ __device__ inline double lm_bits_to_double(unsigned long long b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}

__host__ __device__ inline void lm_push(LmState& st, const pgo_iteration_summary& it, pgo_iteration_summary* log, int cap) {
  if (log && st.num_rows < cap) log[st.num_rows] = it;
  st.num_rows++;
}

// top of TrustRegionMinimizer's loop: the tests made before an iteration starts
__host__ __device__ inline void lm_top_of_loop(LmState& st, const LmOptions& o) {
  if (st.done) return;
  if (st.iter >= o.max_num_iterations) { st.done = 1; st.termination_type = PGO_NO_CONVERGENCE; st.reason = kLmMaxIterations; st.reason_value = st.iter; return; }
  if (st.gradient_max_norm <= o.gradient_tolerance) { st.done = 1; st.termination_type = PGO_CONVERGENCE; st.reason = kLmGradientTolerance; st.reason_value = st.gradient_max_norm; return; }
  if (st.radius < o.min_radius) { st.done = 1; st.termination_type = PGO_CONVERGENCE; st.reason = kLmMinRadius; st.reason_value = st.radius; return; }
  st.iter++;
}

// IterationZero: the scalars hold cost, |x|^2 and the gradient norms at the initial point
__host__ __device__ inline void lm_init(LmState& st, const DeviceScalars& sc, const LmOptions& o, pgo_iteration_summary* log, int cap) {
  st = LmState();
  st.x_cost = sc.cost; st.initial_cost = sc.cost;
  st.x_norm = sqrt(sc.x_norm2);
  st.radius = o.initial_radius; st.decrease_factor = 2.0;
  st.gradient_max_norm = lm_bits_to_double(sc.gmax_bits);
  st.gradient_norm = sqrt(sc.gnorm2);
  pgo_iteration_summary it = pgo_iteration_summary();
  it.cost = st.x_cost; it.trust_region_radius = st.radius;
  it.gradient_max_norm = st.gradient_max_norm; it.gradient_norm = st.gradient_norm;
  lm_push(st, it, log, cap);
  if (!isfinite(st.x_cost)) { st.done = 1; st.termination_type = PGO_FAILURE; st.reason = kLmInitialCostNotFinite; return; }
  lm_top_of_loop(st, o);
}

// One iteration's decision.  sc: step statistics of the linear solve (xtb, xtAx, xtDx, pcg_*), |step|^2, |x_cand|^2, and
// the cost / gradient norms of the speculative linearisation at the candidate.
__host__ __device__ inline void lm_decide(LmState& st, const DeviceScalars& sc, const LmOptions& o, pgo_iteration_summary* log, int cap) {
  if (st.done) return;
  st.accept = 0;
  pgo_iteration_summary it = pgo_iteration_summary();
  it.iteration = st.iter;
  it.gradient_max_norm = st.gradient_max_norm; it.gradient_norm = st.gradient_norm;
  st.total_pcg_iterations += sc.pcg_iterations;
  it.linear_solver_iterations = sc.pcg_iterations;
  it.pcg_relative_residual = sc.pcg_gamma0 > 0 ? sqrt(fabs(sc.pcg_gamma) / sc.pcg_gamma0) : 0.0;
  it.trust_region_radius = st.radius;
  st.reuse_diagonal = 1;
  // model_cost_change = -(J step)^T (r + J step / 2) = y^T g - y^T H y / 2
  const double model_cost_change = sc.xtb - 0.5 * (sc.xtAx - sc.xtDx);
  const bool step_valid = sc.pcg_flag < 2 && isfinite(model_cost_change) && model_cost_change > 0.0;   // 2: CG breakdown, 3: pivot failure
  if (!step_valid) {
    it.step_is_valid = 0; it.cost = st.x_cost;
    if (++st.num_consecutive_invalid >= o.max_consecutive_invalid) {
      st.done = 1; st.termination_type = PGO_FAILURE; st.reason = kLmInvalidSteps;
      lm_push(st, it, log, cap);
      return;
    }
    st.radius /= st.decrease_factor; st.decrease_factor *= 2.0;
    st.num_unsuccessful++;
    lm_push(st, it, log, cap);
    lm_top_of_loop(st, o);
    return;
  }
  st.num_consecutive_invalid = 0;
  it.step_is_valid = 1;
  double cand_cost = sc.cost;
  if (!isfinite(cand_cost)) cand_cost = 1.7976931348623157e308;
  const double step_norm = sqrt(sc.step_norm2);
  it.step_norm = step_norm;
  if (step_norm <= o.parameter_tolerance * (st.x_norm + o.parameter_tolerance)) {
    st.done = 1; st.termination_type = PGO_CONVERGENCE; st.reason = kLmParameterTolerance;
    st.reason_value = step_norm / (st.x_norm + o.parameter_tolerance);
    it.cost = st.x_cost;
    lm_push(st, it, log, cap);
    return;
  }
  const double cost_change = st.x_cost - cand_cost;
  it.cost_change = cost_change;
  if (fabs(cost_change) <= o.function_tolerance * st.x_cost) {
    st.done = 1; st.termination_type = PGO_CONVERGENCE; st.reason = kLmFunctionTolerance;
    st.reason_value = fabs(cost_change) / st.x_cost;
    it.cost = st.x_cost;
    lm_push(st, it, log, cap);
    return;
  }
  const double relative_decrease = cost_change / model_cost_change;
  it.relative_decrease = relative_decrease;
  if (relative_decrease > o.min_relative_decrease) {
    // HandleSuccessfulStep: x = candidate; its linearisation becomes the current system (commit kernel)
    st.accept = 1;
    st.x_norm = sqrt(sc.x_norm2);
    st.x_cost = sc.cost;
    st.gradient_max_norm = lm_bits_to_double(sc.gmax_bits);
    st.gradient_norm = sqrt(sc.gnorm2);
    it.gradient_max_norm = st.gradient_max_norm; it.gradient_norm = st.gradient_norm;
    it.step_is_successful = 1; it.cost = st.x_cost;
    st.num_successful++;
    double t = 2.0 * relative_decrease - 1.0;
    t = 1.0 - t * t * t;
    if (t < 1.0 / 3.0) t = 1.0 / 3.0;
    st.radius = fmin(st.radius / t, o.max_radius);
    st.decrease_factor = 2.0; st.reuse_diagonal = 0;
  } else {
    // HandleUnsuccessfulStep: the speculative system is dropped
    it.step_is_successful = 0; it.cost = st.x_cost;
    st.num_unsuccessful++;
    st.radius /= st.decrease_factor; st.decrease_factor *= 2.0;
  }
  lm_push(st, it, log, cap);
  lm_top_of_loop(st, o);
}

__device__ __forceinline__ unsigned long long lm_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void lm_init_kernel(LmState* st, DeviceScalars* sc, const LmOptions o, pgo_iteration_summary* log, int cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  LmState s;
  lm_init(s, *sc, o, log, cap);
  s.t_mark = lm_globaltimer();
  *st = s;
  // the scalars of the first iteration start from zero
  *sc = DeviceScalars();
}

__global__ void lm_decide_kernel(LmState* st, const DeviceScalars* sc, const LmOptions o, pgo_iteration_summary* log, int cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0 || st->done) return;
  LmState s = *st;
  const unsigned long long now = lm_globaltimer();
  s.solver_ns += (double)(s.t_solved - s.t_mark);
  s.linearize_ns += (double)(now - s.t_solved);
  lm_decide(s, *sc, o, log, cap);
  *st = s;
}

// Accepted step: candidate poses and the candidate's system become current.  Always: the scalars are zeroed for the next
// iteration (nobody reads them between the decision and the next linear solve).
__global__ void __launch_bounds__(256) lm_commit_kernel(LmState* st, DeviceScalars* sc, long long n_pose_words, const double* __restrict__ poses_cand,
                                                        double* __restrict__ poses, long long n_hd, const double* __restrict__ hd_alt,
                                                        double* __restrict__ hd, long long n_ho, const double* __restrict__ ho_alt,
                                                        double* __restrict__ ho, long long n_g, const double* __restrict__ g_alt,
                                                        double* __restrict__ g) {
  if (st->accept) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const double2* a; double2* b;
    a = reinterpret_cast<const double2*>(poses_cand); b = reinterpret_cast<double2*>(poses);
    for (long long k = tid; k < n_pose_words / 2; k += nt) b[k] = a[k];
    a = reinterpret_cast<const double2*>(hd_alt); b = reinterpret_cast<double2*>(hd);
    for (long long k = tid; k < n_hd / 2; k += nt) b[k] = a[k];
    a = reinterpret_cast<const double2*>(ho_alt); b = reinterpret_cast<double2*>(ho);
    for (long long k = tid; k < n_ho / 2; k += nt) b[k] = a[k];
    a = reinterpret_cast<const double2*>(g_alt); b = reinterpret_cast<double2*>(g);
    for (long long k = tid; k < n_g / 2; k += nt) b[k] = a[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *sc = DeviceScalars();
    st->t_mark = lm_globaltimer();
  }
}

}  // namespace pgo
