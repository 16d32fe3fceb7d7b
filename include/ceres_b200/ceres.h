// ceres_b200/ceres.h -- C++ host-side mirror of the slice of the Ceres Solver API that the reference's
// pose-graph path uses, implemented on top of the C-ABI of libpgo_b200.so (include/pgo_b200.h).
//
// REF = /root/reference/src/POSE_GRAPH_CERES_PLUS.  Every name below is one the reference calls:
//
//   ceres::Problem, Problem::AddResidualBlock(cost, loss, p_a, q_a, p_b, q_b)   REF/test/pose_graph_ceres_plus_finial.cpp:58,513-517
//   Problem::SetParameterization(q, EigenQuaternionParameterization)            :496-497,519-522
//   Problem::SetParameterBlockConstant(p) / (q)                                 :526-527
//   ceres::HuberLoss(1.0)                                                       :495
//   ceres::AutoDiffCostFunction<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4>            REF/include/PoseGraph3dError.h:56-61
//   ceres::Solver::Options {max_num_iterations, linear_solver_type}             :534-536
//   ceres::Solve(options, problem, &summary), Summary::FullReport(), IsSolutionUsable()   :538-543
//
// The reference builds a general ceres::Problem, but on this path every residual block is the SE(3)
// relative-pose term (AutoDiffCostFunction<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4>) over parameter blocks
// (p_a[3], q_a[4], p_b[3], q_b[4]) with q under EigenQuaternionParameterization.  This mirror accepts
// exactly that structure and hands the whole problem to the device solver -- including what Ceres allows per block:
// a different LossFunction on every residual block, and p or q of a pose held constant on its own.  Anything else (other
// cost functions, a q block without the Eigen parameterization) is rejected with std::invalid_argument at the call that
// introduces it or at Solve() -- there is no CPU fallback behind this header.
//
// Ownership follows Ceres: the Problem takes ownership of cost functions, loss functions and local
// parameterizations passed by pointer (each distinct pointer is deleted once).
//
// Usage in the reference: NO source change -- put include/ceres_b200/compat first on the include
// path (its ceres/ceres.h and ceres/autodiff_cost_function.h are this header with `namespace ceres = ceres_b200`) and link
// -lpgo_b200.  The reference's own PoseGraph3dError.h / types.h stay as they are: its functor is accepted by
// AutoDiffCostFunction below.  Helper types that Ceres does not have live in ceres_b200::pgo.
#ifndef CERES_B200_CERES_H_
#define CERES_B200_CERES_H_

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../pgo_b200.h"

namespace ceres_b200 {

// ---- enums (names and order of ceres/types.h; every linear solver type maps to the device solver) ----
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };

inline const char* TerminationTypeToString(TerminationType t) {
  switch (t) {
    case CONVERGENCE: return "CONVERGENCE";
    case NO_CONVERGENCE: return "NO_CONVERGENCE";
    case FAILURE: return "FAILURE";
    case USER_SUCCESS: return "USER_SUCCESS";
    default: return "USER_FAILURE";
  }
}

// ---- loss functions (ceres/loss_function.h) ----
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;   // rho, rho', rho''
  virtual int device_type() const = 0;                              // pgo_loss_type
  virtual double device_a() const { return 1.0; }
};
class TrivialLoss : public LossFunction {
 public:
  void Evaluate(double s, double rho[3]) const override { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  int device_type() const override { return PGO_LOSS_TRIVIAL; }
};
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a_ * r - b_; rho[1] = std::max(2.2250738585072014e-308, a_ / r); rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  }
  int device_type() const override { return PGO_LOSS_HUBER; }
  double device_a() const override { return a_; }
 private:
  const double a_, b_;
};
class CauchyLoss : public LossFunction {
 public:
  explicit CauchyLoss(double a) : a_(a), b_(a * a), c_(1.0 / (a * a)) {}
  void Evaluate(double s, double rho[3]) const override {
    const double sum = 1.0 + s * c_, inv = 1.0 / sum;
    rho[0] = b_ * std::log(sum); rho[1] = std::max(2.2250738585072014e-308, inv); rho[2] = -c_ * (inv * inv);
  }
  int device_type() const override { return PGO_LOSS_CAUCHY; }
  double device_a() const override { return a_; }
 private:
  const double a_, b_, c_;
};

// ---- local parameterizations (ceres/local_parameterization.h) ----
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
  virtual bool is_eigen_quaternion() const { return false; }
};
// Plus(x, delta) = [sin(|delta|)/|delta| delta ; cos(|delta|)] * x with x stored x,y,z,w (Eigen coeffs()).
class EigenQuaternionParameterization : public LocalParameterization {
 public:
  int GlobalSize() const override { return 4; }
  int LocalSize() const override { return 3; }
  bool is_eigen_quaternion() const override { return true; }
};

// ---- cost functions (ceres/cost_function.h, ceres/autodiff_cost_function.h) ----
class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual int num_residuals() const = 0;
  virtual const std::vector<int>& parameter_block_sizes() const = 0;
};

// Helper types that real Ceres does NOT have live in ceres_b200::pgo, so that `using namespace ceres;` next to the
// reference's own `using namespace POSE_GRAPH;` (REF/test/pose_graph_ceres_plus_finial.cpp:20-22) stays unambiguous.
namespace pgo {

// Eigen-free pose: p = x y z, q = coeffs() x y z w  (layout of POSE_GRAPH::Pose3d, REF/include/types.h:15-20)
struct PosePod {
  double p[3];
  double q[4];
};

// The one cost function of the path -- the SE(3) relative-pose residual of REF/include/PoseGraph3dError.h:21-54 --
// as data: the measurement t_ab and the square-root information.  It is evaluated on the device by the fused
// linearize kernel; this object only carries its constants.
class PoseGraph3dCost : public CostFunction {
 public:
  PoseGraph3dCost(const double t_ab[7], const double sqrt_information_row_major[36]) : sizes_{3, 4, 3, 4} {
    std::memcpy(t_ab_, t_ab, sizeof t_ab_);
    std::memcpy(sqrt_info_, sqrt_information_row_major, sizeof sqrt_info_);
  }
  int num_residuals() const override { return 6; }
  const std::vector<int>& parameter_block_sizes() const override { return sizes_; }
  const double* t_ab() const { return t_ab_; }
  const double* sqrt_information() const { return sqrt_info_; }
 protected:
  PoseGraph3dCost() : sizes_{3, 4, 3, 4} {}
  std::vector<int> sizes_;
  double t_ab_[7];
  double sqrt_info_[36];
};

// plain-array form: t_ab = x y z qx qy qz qw, sqrt_information ROW-major 6x6 (nullptr = identity)
inline CostFunction* MakePoseGraph3dCost(const double t_ab[7], const double* sqrt_information_row_major) {
  double s[36];
  if (sqrt_information_row_major) std::memcpy(s, sqrt_information_row_major, sizeof s);
  else for (int k = 0; k < 36; ++k) s[k] = (k / 6 == k % 6) ? 1.0 : 0.0;
  return new PoseGraph3dCost(t_ab, s);
}
// Eigen-style form: t.p is indexable as p(i), t.q has x() y() z() w(); Matrix6 is indexable as m(i, j)
template <typename PoseT, typename Matrix6,
          typename = typename std::enable_if<std::is_class<PoseT>::value && std::is_class<Matrix6>::value>::type>
inline CostFunction* MakePoseGraph3dCost(const PoseT& t, const Matrix6& sqrt_information) {
  double a[7] = {t.p(0), t.p(1), t.p(2), t.q.x(), t.q.y(), t.q.z(), t.q.w()}, s[36];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) s[i * 6 + j] = sqrt_information(i, j);
  return new PoseGraph3dCost(a, s);
}

// ---- small dense helpers for the functor probe below ----
inline double det3(const double a[3], const double b[3], const double c[3]) {
  return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}
// closed form of the residual with q_m unit: r = S [R(q_a)^T (p_b - p_a) - p_m ; 2 vec(q_m * conj(conj(q_a) * q_b))]
inline void pose_graph_residual(const double t_ab[7], const double S[36], const double* pa, const double* qa, const double* pb,
                                const double* qb, double r[6]) {
  auto qmul = [](const double* a, const double* b, double* o) {   // x y z w
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  };
  const double qi[4] = {-qa[0], -qa[1], -qa[2], qa[3]};
  const double d[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
  // Eigen's quaternion * vector: v + w t + u x t with t = 2 u x v
  const double tx = 2 * (qi[1] * d[2] - qi[2] * d[1]), ty = 2 * (qi[2] * d[0] - qi[0] * d[2]), tz = 2 * (qi[0] * d[1] - qi[1] * d[0]);
  double e[6];
  e[0] = d[0] + qi[3] * tx + (qi[1] * tz - qi[2] * ty) - t_ab[0];
  e[1] = d[1] + qi[3] * ty + (qi[2] * tx - qi[0] * tz) - t_ab[1];
  e[2] = d[2] + qi[3] * tz + (qi[0] * ty - qi[1] * tx) - t_ab[2];
  double qab[4], dq[4];
  qmul(qi, qb, qab);
  const double qabc[4] = {-qab[0], -qab[1], -qab[2], qab[3]};
  qmul(t_ab + 3, qabc, dq);
  e[3] = 2 * dq[0]; e[4] = 2 * dq[1]; e[5] = 2 * dq[2];
  for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < 6; ++k) s += S[i * 6 + k] * e[k]; r[i] = s; }
}

}  // namespace pgo

// ceres::AutoDiffCostFunction<Functor, 6, 3, 4, 3, 4>(new Functor(...)) -- what PoseGraph3dErrorTerm::Create builds
// (REF/include/PoseGraph3dError.h:56-61).  The device path does not differentiate a functor: it evaluates the SE(3)
// relative-pose residual in closed form.  So the functor is PROBED once, on the host, at construction: with p_a = 0 and
// q_a = identity the functor is f(p_b, q_b) = S_p (p_b - p_m) + 2 S_r L(q_m) q_b with L = [-w_m I - [v_m]x | v_m]
// (quaternion products are bilinear and the functor never normalises), so nine evaluations give S_p (columns of
// f(e_k, 0) - f(0, 0)), G = 2 S_r L (columns f(0, e_b) - f(0, 0)), q_m as the null vector of G (L q_m = 0, L L^T = I),
// S_r = G L^T / 2 and p_m from S_p p_m = -f(0, 0).  The recovered constants are then checked against the functor at
// random poses; a functor that is not this residual is rejected with std::invalid_argument.  (Problem setup, not a CPU
// fallback: no residual or Jacobian of the solve is ever evaluated on the host.)
template <typename CostFunctor, int kNumResiduals, int N0 = 0, int N1 = 0, int N2 = 0, int N3 = 0, int N4 = 0, int N5 = 0,
          int N6 = 0, int N7 = 0, int N8 = 0, int N9 = 0>
class AutoDiffCostFunction : public pgo::PoseGraph3dCost {
  static_assert(kNumResiduals == 6 && N0 == 3 && N1 == 4 && N2 == 3 && N3 == 4 && N4 == 0,
                "ceres_b200: only the SE(3) relative-pose residual <6, 3, 4, 3, 4> runs on the device path");
 public:
  explicit AutoDiffCostFunction(CostFunctor* functor) : functor_(functor) {
    if (!functor) throw std::invalid_argument("ceres_b200::AutoDiffCostFunction: null functor");
    probe();
  }
  ~AutoDiffCostFunction() override { delete functor_; }
  AutoDiffCostFunction(const AutoDiffCostFunction&) = delete;
  AutoDiffCostFunction& operator=(const AutoDiffCostFunction&) = delete;

 private:
  void eval(const double* pa, const double* qa, const double* pb, const double* qb, double* r) const {
    if (!(*functor_)(pa, qa, pb, qb, r)) throw std::invalid_argument("ceres_b200::AutoDiffCostFunction: the functor returned false");
  }
  void probe() {
    const double zero3[3] = {0, 0, 0}, ident[4] = {0, 0, 0, 1}, zero4[4] = {0, 0, 0, 0};
    double c0[6], f[6], Sp[6][3], G[6][4];
    eval(zero3, ident, zero3, zero4, c0);
    for (int k = 0; k < 3; ++k) {
      double e[3] = {0, 0, 0};
      e[k] = 1.0;
      eval(zero3, ident, e, zero4, f);
      for (int i = 0; i < 6; ++i) Sp[i][k] = f[i] - c0[i];
    }
    for (int b = 0; b < 4; ++b) {
      double e[4] = {0, 0, 0, 0};
      e[b] = 1.0;
      eval(zero3, ident, zero3, e, f);
      for (int i = 0; i < 6; ++i) G[i][b] = f[i] - c0[i];
    }
    // q_m: null vector of G = the 4-D cross product of three independent rows; take the best-conditioned triple
    double qm[4] = {0, 0, 0, 1}, best = -1.0;
    for (int i = 0; i < 6; ++i) for (int j = i + 1; j < 6; ++j) for (int k = j + 1; k < 6; ++k) {
      double n[4];
      for (int c = 0; c < 4; ++c) {
        double ra[3], rb[3], rc[3];
        for (int t = 0, u = 0; t < 4; ++t) if (t != c) { ra[u] = G[i][t]; rb[u] = G[j][t]; rc[u] = G[k][t]; ++u; }
        n[c] = ((c & 1) ? -1.0 : 1.0) * pgo::det3(ra, rb, rc);
      }
      const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3];
      if (nn > best) { best = nn; std::memcpy(qm, n, sizeof qm); }
    }
    if (!(best > 0.0)) throw std::invalid_argument("ceres_b200::AutoDiffCostFunction: rank-deficient orientation information, cannot recover the measurement");
    {
      const double inv = 1.0 / std::sqrt(best) * (qm[3] < 0 ? -1.0 : 1.0);
      for (double& v : qm) v *= inv;
    }
    // S_r = G L^T / 2 with L = [-w I - [v]x | v]
    const double vx = qm[0], vy = qm[1], vz = qm[2], w = qm[3];
    const double L[3][4] = {{-w, vz, -vy, vx}, {-vz, -w, vx, vy}, {vy, -vx, -w, vz}};
    double S[36];
    for (int i = 0; i < 6; ++i) {
      for (int k = 0; k < 3; ++k) S[i * 6 + k] = Sp[i][k];
      for (int k = 0; k < 3; ++k) { double s = 0; for (int b = 0; b < 4; ++b) s += G[i][b] * L[k][b]; S[i * 6 + 3 + k] = 0.5 * s; }
    }
    // p_m from the normal equations of S_p p_m = -c0
    double A[3][3], rhs[3];
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) { double s = 0; for (int i = 0; i < 6; ++i) s += Sp[i][a] * Sp[i][b]; A[a][b] = s; }
      double s = 0;
      for (int i = 0; i < 6; ++i) s -= Sp[i][a] * c0[i];
      rhs[a] = s;
    }
    const double det = pgo::det3(A[0], A[1], A[2]);
    if (!(std::fabs(det) > 0.0)) throw std::invalid_argument("ceres_b200::AutoDiffCostFunction: rank-deficient position information, cannot recover the measurement");
    double pm[3];
    for (int c = 0; c < 3; ++c) {
      double M[3][3];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) M[a][b] = (b == c) ? rhs[a] : A[a][b];
      pm[c] = pgo::det3(M[0], M[1], M[2]) / det;
    }
    // one step of refinement of p_m against the rounding of the normal equations
    {
      double res[3] = {0, 0, 0};
      for (int a = 0; a < 3; ++a) { double s = rhs[a]; for (int b = 0; b < 3; ++b) s -= A[a][b] * pm[b]; res[a] = s; }
      for (int c = 0; c < 3; ++c) {
        double M[3][3];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) M[a][b] = (b == c) ? res[a] : A[a][b];
        pm[c] += pgo::det3(M[0], M[1], M[2]) / det;
      }
    }
    t_ab_[0] = pm[0]; t_ab_[1] = pm[1]; t_ab_[2] = pm[2];
    std::memcpy(t_ab_ + 3, qm, sizeof qm);
    std::memcpy(sqrt_info_, S, sizeof S);
    // the functor must BE this residual: compare at poses that exercise every term
    const double pa[3] = {0.3, -1.1, 0.7}, pb[3] = {-0.4, 0.9, 1.6};
    const double qa[4] = {0.18257418583505536, -0.3651483716701107, 0.5477225575051661, 0.7302967433402214};
    const double qb[4] = {-0.4, 0.2, 0.1, 0.8888194417315589};
    double want[6], got[6], scale = 0.0, err = 0.0;
    eval(pa, qa, pb, qb, got);
    pgo::pose_graph_residual(t_ab_, sqrt_info_, pa, qa, pb, qb, want);
    for (int i = 0; i < 6; ++i) { scale = std::max(scale, std::fabs(want[i])); err = std::max(err, std::fabs(want[i] - got[i])); }
    if (!(err <= 1e-9 * std::max(scale, 1.0)))
      throw std::invalid_argument("ceres_b200::AutoDiffCostFunction: the functor is not the SE(3) relative-pose residual of "
                                  "PoseGraph3dError.h (the only cost function the device path evaluates)");
  }
  CostFunctor* functor_;
};

class Problem;
struct Solver {
  // ceres::Solver::Options: the fields that act on this path, Ceres' defaults.
  struct Options {
    int max_num_iterations = 50;
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    double function_tolerance = 1e-6;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4;
    double max_trust_region_radius = 1e16;
    double min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3;
    double min_lm_diagonal = 1e-6;
    double max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
    bool jacobi_scaling = true;
    bool minimizer_progress_to_stdout = false;
    int num_threads = 1;                 // accepted, unused: the device path is always parallel
    // device-specific knobs (not in Ceres)
    int device = 0;
    int device_linear_solver = PGO_LINEAR_AUTO;   // pgo_linear_solver_type
    double pcg_tolerance = 1e-8;
    int pcg_max_iterations = 20000;
    double direct_residual_accept = 1e-8;
  };
  struct IterationSummary {
    int iteration = 0;
    bool step_is_valid = false, step_is_successful = false;
    double cost = 0, cost_change = 0, gradient_max_norm = 0, gradient_norm = 0, step_norm = 0;
    double relative_decrease = 0, trust_region_radius = 0;
    int linear_solver_iterations = 0;
  };
  struct Summary {
    TerminationType termination_type = FAILURE;
    std::string message = "ceres_b200::Solve was not called.";
    double initial_cost = -1, final_cost = -1;
    int num_successful_steps = -1, num_unsuccessful_steps = -1;
    std::vector<IterationSummary> iterations;
    double total_time_in_seconds = -1;
    int num_parameter_blocks = -1, num_parameters = -1, num_effective_parameters = -1;
    int num_residual_blocks = -1, num_residuals = -1;
    int num_parameter_blocks_reduced = -1, num_parameters_reduced = -1, num_effective_parameters_reduced = -1;
    pgo_solver_summary device;            // device-side counters and timings
    bool IsSolutionUsable() const {
      return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE || termination_type == USER_SUCCESS;
    }
    std::string BriefReport() const {
      char buf[256];
      snprintf(buf, sizeof buf, "Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
               num_successful_steps + num_unsuccessful_steps, initial_cost, final_cost, TerminationTypeToString(termination_type));
      return buf;
    }
    std::string FullReport() const {
      char buf[2048];
      const char* ls = device.linear_solver_used == PGO_LINEAR_PCG_LEVEL_CHOLESKY ? "PCG + LEVEL_SCHEDULED_BLOCK_CHOLESKY (sm_100a)"
                                                                                  : "PCG + BLOCK_JACOBI (sm_100a)";
      snprintf(buf, sizeof buf,
               "\nSolver Summary (v ceres_b200 / libpgo_b200 ABI %d)\n\n"
               "                                     Original                  Reduced\n"
               "Parameter blocks                 %12d             %12d\n"
               "Parameters                       %12d             %12d\n"
               "Effective parameters             %12d             %12d\n"
               "Residual blocks                  %12d             %12d\n"
               "Residuals                        %12d             %12d\n\n"
               "Minimizer                        TRUST_REGION\n"
               "Trust region strategy     LEVENBERG_MARQUARDT\n"
               "Linear solver          %s\n"
               "Hessian blocks (6x6)             %12lld\nFactor blocks (6x6)              %12lld (levels %d)\n\n"
               "Cost:\nInitial                          %e\nFinal                            %e\nChange                           %e\n\n"
               "Minimizer iterations             %12d\nSuccessful steps                 %12d\nUnsuccessful steps               %12d\n"
               "Linear solver iterations (PCG)   %12lld\nKernel launches                  %12lld\n\n"
               "Time (in seconds):\n  Setup (analysis + upload)      %12.6f\n  Residual+Jacobian kernels      %12.6f\n"
               "  Linear solver kernels          %12.6f\nTotal                            %12.6f\n\n"
               "Termination:                     %s (%s)\n",
               pgo_abi_version(), num_parameter_blocks, num_parameter_blocks_reduced, num_parameters, num_parameters_reduced,
               num_effective_parameters, num_effective_parameters_reduced, num_residual_blocks, num_residual_blocks,
               num_residuals, num_residuals, ls, device.hessian_blocks, device.factor_blocks, device.factor_levels,
               initial_cost, final_cost, initial_cost - final_cost, num_successful_steps + num_unsuccessful_steps + 1,
               num_successful_steps, num_unsuccessful_steps, device.total_pcg_iterations, device.kernel_launches,
               device.time_setup_s, device.time_linearize_ms * 1e-3, device.time_linear_solver_ms * 1e-3,
               total_time_in_seconds, TerminationTypeToString(termination_type), message.c_str());
      return buf;
    }
  };
};

class Problem {
 public:
  struct Options {
    Ownership cost_function_ownership = TAKE_OWNERSHIP;
    Ownership loss_function_ownership = TAKE_OWNERSHIP;
    Ownership local_parameterization_ownership = TAKE_OWNERSHIP;
  };
  Problem() {}
  explicit Problem(const Options& o) : options_(o) {}
  Problem(const Problem&) = delete;
  Problem& operator=(const Problem&) = delete;
  ~Problem() {
    if (options_.cost_function_ownership == TAKE_OWNERSHIP) for (CostFunction* c : owned_costs_) delete c;
    if (options_.loss_function_ownership == TAKE_OWNERSHIP) for (LossFunction* l : owned_losses_) delete l;
    if (options_.local_parameterization_ownership == TAKE_OWNERSHIP) for (LocalParameterization* l : owned_params_) delete l;
  }

  typedef int ResidualBlockId;

  // problem->AddResidualBlock(cost_function, loss_function, p_a, q_a, p_b, q_b)   (REF ...plus_finial.cpp:513-517)
  ResidualBlockId AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function, double* p_a, double* q_a,
                                   double* p_b, double* q_b) {
    // ownership first (Ceres owns what it is handed, also when the call fails afterwards); validate before mutating
    if (cost_function) owned_costs_.insert(cost_function);
    if (loss_function) owned_losses_.insert(loss_function);
    pgo::PoseGraph3dCost* c = dynamic_cast<pgo::PoseGraph3dCost*>(cost_function);
    if (!c) throw std::invalid_argument("ceres_b200::Problem::AddResidualBlock: only SE(3) relative-pose cost functions (AutoDiffCostFunction<F, 6, 3, 4, 3, 4> / pgo::MakePoseGraph3dCost) run on the device path");
    if (!p_a || !q_a || !p_b || !q_b) throw std::invalid_argument("ceres_b200::Problem::AddResidualBlock: null parameter block");
    if (p_a == p_b || q_a == q_b || p_a == q_a || p_a == q_b || p_b == q_a || p_b == q_b)
      throw std::invalid_argument("ceres_b200::Problem::AddResidualBlock: duplicate parameter blocks in one residual block");
    check_pairing(p_a, q_a);          // nothing is registered unless the whole call is valid
    check_pairing(p_b, q_b);
    const int a = pose_of(p_a, q_a), b = pose_of(p_b, q_b);
    edges_.push_back({a, b, c, loss_function});
    return (int)edges_.size() - 1;
  }

  ResidualBlockId AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function, const std::vector<double*>& blocks) {
    if (blocks.size() != 4) throw std::invalid_argument("ceres_b200::Problem::AddResidualBlock: the pose-graph residual takes 4 parameter blocks");
    return AddResidualBlock(cost_function, loss_function, blocks[0], blocks[1], blocks[2], blocks[3]);
  }

  void AddParameterBlock(double* values, int size) { block_of(values, size); }
  void AddParameterBlock(double* values, int size, LocalParameterization* lp) { block_of(values, size); SetParameterization(values, lp); }

  void SetParameterization(double* values, LocalParameterization* local_parameterization) {
    Block& blk = find_block(values, "SetParameterization");
    if (!local_parameterization) throw std::invalid_argument("ceres_b200::Problem::SetParameterization: null parameterization");
    if (local_parameterization->GlobalSize() != blk.size) throw std::invalid_argument("ceres_b200::Problem::SetParameterization: size mismatch");
    blk.param = local_parameterization;
    owned_params_.insert(local_parameterization);
  }
  void SetParameterBlockConstant(double* values) { find_block(values, "SetParameterBlockConstant").constant = true; }
  void SetParameterBlockVariable(double* values) { find_block(values, "SetParameterBlockVariable").constant = false; }
  bool IsParameterBlockConstant(double* values) const {
    auto it = blocks_.find(values);
    if (it == blocks_.end()) throw std::invalid_argument("ceres_b200::Problem::IsParameterBlockConstant: unknown parameter block");
    return it->second.constant;
  }
  bool HasParameterBlock(const double* values) const { return blocks_.count(const_cast<double*>(values)) != 0; }

  int NumParameterBlocks() const { return (int)blocks_.size(); }
  int NumParameters() const { int n = 0; for (auto& kv : blocks_) n += kv.second.size; return n; }
  int NumResidualBlocks() const { return (int)edges_.size(); }
  int NumResiduals() const { return 6 * (int)edges_.size(); }

  // Problem::Evaluate at the current parameter values: cost, robustified residuals [6 * #blocks] and the
  // gradient in local (tangent) coordinates, per pose [6] in the order poses were first seen.
  bool Evaluate(double* cost, std::vector<double>* residuals, std::vector<double>* gradient, int device = 0) {
    Packed pk;
    pack(&pk);
    pgo_graph* g = nullptr;
    int rc = pgo_graph_create(&g, device, pk.n_poses, pk.n_edges, pk.poses.data(), pk.ids.data(), pk.meas.data(),
                              pk.identity_info ? nullptr : pk.sqrt_info.data(), pk.pose_const.data());
    if (rc != PGO_OK) throw std::runtime_error(std::string("ceres_b200::Problem::Evaluate: ") + pgo_last_error());
    if (residuals) residuals->assign((size_t)6 * pk.n_edges, 0.0);
    if (gradient) gradient->assign((size_t)6 * pk.n_poses, 0.0);
    if (!pk.uniform_loss) {
      rc = pgo_graph_set_edge_losses(g, pk.edge_loss_type.data(), pk.edge_loss_a.data());
      if (rc != PGO_OK) { pgo_graph_destroy(g); throw std::runtime_error(std::string("ceres_b200::Problem::Evaluate: ") + pgo_last_error()); }
    }
    rc = pgo_graph_evaluate(g, pk.loss_type, pk.loss_a, cost, residuals ? residuals->data() : nullptr,
                            gradient ? gradient->data() : nullptr, nullptr);
    pgo_graph_destroy(g);
    if (rc != PGO_OK) throw std::runtime_error(std::string("ceres_b200::Problem::Evaluate: ") + pgo_last_error());
    return true;
  }

 private:
  friend void Solve(const Solver::Options&, Problem*, Solver::Summary*);
  struct Block { int size = 0; LocalParameterization* param = nullptr; bool constant = false; int pose = -1; };
  struct Edge { int a, b; pgo::PoseGraph3dCost* cost; LossFunction* loss; };
  struct PoseRef { double* p; double* q; };
  struct Packed {
    int n_poses = 0, n_edges = 0, loss_type = PGO_LOSS_TRIVIAL;
    double loss_a = 1.0;
    bool identity_info = true;
    bool uniform_loss = true;                 // every residual block carries the same loss (type and scale)
    std::vector<int> edge_loss_type;          // per residual block (ceres::Problem::AddResidualBlock takes the loss per block)
    std::vector<double> edge_loss_a;
    std::vector<double> poses, meas, sqrt_info;
    std::vector<int> ids;
    std::vector<unsigned char> pose_const;
  };

  Block& block_of(double* values, int size) {
    if (!values) throw std::invalid_argument("ceres_b200::Problem: null parameter block");
    auto it = blocks_.find(values);
    if (it == blocks_.end()) { Block b; b.size = size; it = blocks_.emplace(values, b).first; }
    else if (it->second.size != size) throw std::invalid_argument("ceres_b200::Problem: parameter block re-added with a different size");
    return it->second;
  }
  Block& find_block(double* values, const char* who) {
    auto it = blocks_.find(values);
    if (it == blocks_.end()) throw std::invalid_argument(std::string("ceres_b200::Problem::") + who + ": parameter block not found (add a residual block first, as Ceres requires)");
    return it->second;
  }
  // would pose_of(p, q) throw?  (same rules, no side effects)
  void check_pairing(double* p, double* q) const {
    auto ip = blocks_.find(p), iq = blocks_.find(q);
    if ((ip != blocks_.end() && ip->second.size != 3) || (iq != blocks_.end() && iq->second.size != 4))
      throw std::invalid_argument("ceres_b200::Problem: parameter block re-added with a different size");
    const int pp = ip == blocks_.end() ? -1 : ip->second.pose, pq = iq == blocks_.end() ? -1 : iq->second.pose;
    if (pp != pq && (pp >= 0 || pq >= 0))
      throw std::invalid_argument("ceres_b200::Problem::AddResidualBlock: a position block must always be paired with the same quaternion block");
  }
  int pose_of(double* p, double* q) {
    Block& bp = block_of(p, 3);
    Block& bq = block_of(q, 4);
    if (bp.pose < 0 && bq.pose < 0) {
      bp.pose = bq.pose = (int)poses_.size();
      poses_.push_back({p, q});
    } else if (bp.pose != bq.pose) {
      throw std::invalid_argument("ceres_b200::Problem::AddResidualBlock: a position block must always be paired with the same quaternion block");
    }
    return bp.pose;
  }

  void pack(Packed* pk) const {
    pk->n_poses = (int)poses_.size();
    pk->n_edges = (int)edges_.size();
    if (pk->n_poses == 0) throw std::invalid_argument("ceres_b200: the problem has no parameter blocks");
    pk->poses.resize((size_t)7 * pk->n_poses);
    pk->pose_const.assign(pk->n_poses, 0);
    for (int i = 0; i < pk->n_poses; ++i) {
      const Block& bp = blocks_.at(poses_[i].p);
      const Block& bq = blocks_.at(poses_[i].q);
      if (!bq.param || !bq.param->is_eigen_quaternion())
        throw std::invalid_argument("ceres_b200: every quaternion block needs SetParameterization(q, new EigenQuaternionParameterization)");
      if (bp.param) throw std::invalid_argument("ceres_b200: position blocks must not carry a local parameterization");
      // SetParameterBlockConstant acts per block (REF :526-527 are two calls): 1 = both, 2 = only p, 3 = only q
      pk->pose_const[i] = (bp.constant && bq.constant) ? 1 : bp.constant ? 2 : bq.constant ? 3 : 0;
      std::memcpy(&pk->poses[7 * (size_t)i], poses_[i].p, 24);
      std::memcpy(&pk->poses[7 * (size_t)i + 3], poses_[i].q, 32);
    }
    pk->ids.resize((size_t)2 * pk->n_edges);
    pk->meas.resize((size_t)7 * pk->n_edges);
    pk->sqrt_info.resize((size_t)36 * pk->n_edges);
    bool have_loss = false;
    pk->edge_loss_type.resize((size_t)pk->n_edges);
    pk->edge_loss_a.resize((size_t)pk->n_edges);
    for (int e = 0; e < pk->n_edges; ++e) {
      const Edge& ed = edges_[e];
      pk->ids[2 * (size_t)e] = ed.a; pk->ids[2 * (size_t)e + 1] = ed.b;
      std::memcpy(&pk->meas[7 * (size_t)e], ed.cost->t_ab(), 56);
      std::memcpy(&pk->sqrt_info[36 * (size_t)e], ed.cost->sqrt_information(), 288);
      for (int k = 0; k < 36 && pk->identity_info; ++k)
        if (ed.cost->sqrt_information()[k] != ((k / 6 == k % 6) ? 1.0 : 0.0)) pk->identity_info = false;
      const int lt = ed.loss ? ed.loss->device_type() : (int)PGO_LOSS_TRIVIAL;
      const double la = ed.loss ? ed.loss->device_a() : 1.0;
      pk->edge_loss_type[e] = lt; pk->edge_loss_a[e] = la;
      if (!have_loss) { pk->loss_type = lt; pk->loss_a = la; have_loss = true; }
      else if (lt != pk->loss_type || (lt != PGO_LOSS_TRIVIAL && la != pk->loss_a)) pk->uniform_loss = false;
    }
  }
  void unpack(const std::vector<double>& poses) {
    for (size_t i = 0; i < poses_.size(); ++i) {
      std::memcpy(poses_[i].p, &poses[7 * i], 24);
      std::memcpy(poses_[i].q, &poses[7 * i + 3], 32);
    }
  }

  Options options_;
  std::unordered_map<double*, Block> blocks_;
  std::vector<PoseRef> poses_;
  std::vector<Edge> edges_;
  std::set<CostFunction*> owned_costs_;
  std::set<LossFunction*> owned_losses_;
  std::set<LocalParameterization*> owned_params_;
};

// ceres::Solve(options, problem, &summary)   (REF test/pose_graph_ceres_plus_finial.cpp:538-539)
inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  if (!problem || !summary) throw std::invalid_argument("ceres_b200::Solve: null argument");
  *summary = Solver::Summary();
  Problem::Packed pk;
  problem->pack(&pk);
  pgo_solver_options o;
  pgo_default_options(&o);
  o.max_num_iterations = options.max_num_iterations;
  o.function_tolerance = options.function_tolerance;
  o.gradient_tolerance = options.gradient_tolerance;
  o.parameter_tolerance = options.parameter_tolerance;
  o.initial_trust_region_radius = options.initial_trust_region_radius;
  o.max_trust_region_radius = options.max_trust_region_radius;
  o.min_trust_region_radius = options.min_trust_region_radius;
  o.min_relative_decrease = options.min_relative_decrease;
  o.min_lm_diagonal = options.min_lm_diagonal;
  o.max_lm_diagonal = options.max_lm_diagonal;
  o.max_num_consecutive_invalid_steps = options.max_num_consecutive_invalid_steps;
  o.jacobi_scaling = options.jacobi_scaling ? 1 : 0;
  o.loss_type = pk.loss_type;
  o.loss_a = pk.loss_a;
  o.linear_solver_type = options.device_linear_solver;
  o.pcg_tolerance = options.pcg_tolerance;
  o.pcg_max_iterations = options.pcg_max_iterations;
  o.direct_residual_accept = options.direct_residual_accept;
  if (!pk.uniform_loss) { o.edge_loss_type = pk.edge_loss_type.data(); o.edge_loss_a = pk.edge_loss_a.data(); }
  const int cap = options.max_num_iterations + 2;
  std::vector<pgo_iteration_summary> log((size_t)std::max(cap, 2));
  std::memset(&summary->device, 0, sizeof summary->device);
  const int rc = pgo_solve_pose_graph(options.device, pk.n_poses, pk.poses.data(), pk.n_edges, pk.ids.data(), pk.meas.data(),
                                      pk.identity_info ? nullptr : pk.sqrt_info.data(), pk.pose_const.data(), &o,
                                      &summary->device, log.data(), (int)log.size());
  int n_const = 0;
  for (unsigned char c : pk.pose_const) n_const += c;
  summary->num_parameter_blocks = 2 * pk.n_poses;
  summary->num_parameters = 7 * pk.n_poses;
  summary->num_effective_parameters = 6 * pk.n_poses;
  summary->num_residual_blocks = pk.n_edges;
  summary->num_residuals = 6 * pk.n_edges;
  summary->num_parameter_blocks_reduced = 2 * (pk.n_poses - n_const);
  summary->num_parameters_reduced = 7 * (pk.n_poses - n_const);
  summary->num_effective_parameters_reduced = 6 * (pk.n_poses - n_const);
  if (rc != PGO_OK) {
    summary->termination_type = FAILURE;
    summary->message = std::string("libpgo_b200: ") + pgo_last_error();
    return;
  }
  const pgo_solver_summary& d = summary->device;
  summary->termination_type = d.termination_type == PGO_CONVERGENCE ? CONVERGENCE : d.termination_type == PGO_NO_CONVERGENCE ? NO_CONVERGENCE : FAILURE;
  if (summary->IsSolutionUsable()) problem->unpack(pk.poses);   // Ceres updates the user's state only for a usable solution
  summary->message = d.message;
  summary->initial_cost = d.initial_cost;
  summary->final_cost = d.final_cost;
  summary->num_successful_steps = d.num_successful_steps;
  summary->num_unsuccessful_steps = d.num_unsuccessful_steps;
  summary->total_time_in_seconds = d.time_total_s + d.time_setup_s;
  const int rows = std::min<int>(d.num_iterations, (int)log.size());
  for (int k = 0; k < rows; ++k) {
    Solver::IterationSummary it;
    it.iteration = log[k].iteration; it.step_is_valid = log[k].step_is_valid != 0; it.step_is_successful = log[k].step_is_successful != 0;
    it.cost = log[k].cost; it.cost_change = log[k].cost_change; it.gradient_max_norm = log[k].gradient_max_norm;
    it.gradient_norm = log[k].gradient_norm; it.step_norm = log[k].step_norm; it.relative_decrease = log[k].relative_decrease;
    it.trust_region_radius = log[k].trust_region_radius; it.linear_solver_iterations = log[k].linear_solver_iterations;
    summary->iterations.push_back(it);
  }
  if (options.minimizer_progress_to_stdout) {
    printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter\n");
    for (const Solver::IterationSummary& it : summary->iterations)
      printf("%4d % .6e  % .2e  % .2e  % .2e  % .2e % .2e  %7d\n", it.iteration, it.cost, it.cost_change, it.gradient_max_norm,
             it.step_norm, it.relative_decrease, it.trust_region_radius, it.linear_solver_iterations);
  }
}

}  // namespace ceres_b200

#ifdef CERES_B200_AS_CERES
namespace ceres = ceres_b200;
#endif

#endif  // CERES_B200_CERES_H_
