// ref_functor.cpp -- TEST INFRASTRUCTURE (oracle/_ref): evaluates the REFERENCE's own cost functor,
//   POSE_GRAPH::PoseGraph3dErrorTerm::operator()   /root/reference/src/POSE_GRAPH_CERES_PLUS/include/PoseGraph3dError.h:21-54
// compiled UNMODIFIED from where it lies (oracle/Makefile adds the reference's include directory), against the minimal
// Eigen / Ceres stand-ins under oracle/ref_shim/ (neither library is in this image).  What comes from the reference is the
// functor's algebra -- which quantities are composed, in which order and with which signs; what comes from this repo is
// the quaternion / vector arithmetic underneath it (ref_shim/Eigen/Core, following Eigen 3.3) and the forward-mode Jet
// below (Ceres' AutoDiffCostFunction evaluates the same functor on ceres::Jet<double, 14>).
// tests/test_oracle_cpu.py compares oracle_evaluate()'s residuals and Jacobians with these on random edges.
#include "PoseGraph3dError.h"

namespace {
struct Jet {                       // value + derivatives w.r.t. (p_a[3], q_a[4], p_b[3], q_b[4])
  double a;
  double v[14];
  Jet() : a(0.0) { for (int k = 0; k < 14; ++k) v[k] = 0.0; }
  explicit Jet(double x) : a(x) { for (int k = 0; k < 14; ++k) v[k] = 0.0; }
};
inline Jet operator+(const Jet& x, const Jet& y) { Jet r; r.a = x.a + y.a; for (int k = 0; k < 14; ++k) r.v[k] = x.v[k] + y.v[k]; return r; }
inline Jet operator-(const Jet& x, const Jet& y) { Jet r; r.a = x.a - y.a; for (int k = 0; k < 14; ++k) r.v[k] = x.v[k] - y.v[k]; return r; }
inline Jet operator-(const Jet& x) { Jet r; r.a = -x.a; for (int k = 0; k < 14; ++k) r.v[k] = -x.v[k]; return r; }
inline Jet operator*(const Jet& x, const Jet& y) { Jet r; r.a = x.a * y.a; for (int k = 0; k < 14; ++k) r.v[k] = x.a * y.v[k] + x.v[k] * y.a; return r; }
}  // namespace

static POSE_GRAPH::PoseGraph3dErrorTerm make_term(const double* meas, const double* sqrt_info) {
  POSE_GRAPH::Pose3d t;
  t.p = Eigen::Vector3d(meas[0], meas[1], meas[2]);
  t.q = Eigen::Quaterniond(meas[6], meas[3], meas[4], meas[5]);      // (w, x, y, z) constructor; meas is x y z qx qy qz qw
  Eigen::Matrix<double, 6, 6> S;
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) S(i, j) = sqrt_info[6 * i + j];   // row-major input
  return POSE_GRAPH::PoseGraph3dErrorTerm(t, S);
}

extern "C" {

// residuals[6] of one edge; pose_a / pose_b are x y z qx qy qz qw (p block, then q block in Eigen coeffs() order)
void ref_edge_residual(const double* pose_a, const double* pose_b, const double* meas, const double* sqrt_info, double* residuals) {
  const POSE_GRAPH::PoseGraph3dErrorTerm term = make_term(meas, sqrt_info);
  term(pose_a, pose_a + 3, pose_b, pose_b + 3, residuals);
}

// residuals[6] and the 6 x 14 row-major Jacobian w.r.t. the ambient parameters (p_a, q_a, p_b, q_b), by forward-mode
// differentiation of the reference's functor -- what AutoDiffCostFunction<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4> computes
void ref_edge_jacobian(const double* pose_a, const double* pose_b, const double* meas, const double* sqrt_info,
                       double* residuals, double* jacobian) {
  const POSE_GRAPH::PoseGraph3dErrorTerm term = make_term(meas, sqrt_info);
  Jet x[14], r[6];
  for (int k = 0; k < 7; ++k) { x[k] = Jet(pose_a[k]); x[k].v[k] = 1.0; x[7 + k] = Jet(pose_b[k]); x[7 + k].v[7 + k] = 1.0; }
  term(x, x + 3, x + 7, x + 10, r);
  for (int i = 0; i < 6; ++i) {
    residuals[i] = r[i].a;
    for (int k = 0; k < 14; ++k) jacobian[14 * i + k] = r[i].v[k];
  }
}

}  // extern "C"
