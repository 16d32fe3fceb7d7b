/*
 * pgo_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See pgo_oracle.h.
 *
 * Every function cites the reference line (or the Ceres Solver component the reference calls)
 * it restates.  REF = /root/reference/src/POSE_GRAPH_CERES_PLUS.
 *
 * PARITY PARTIALLY PINNED: no real Ceres run is available to check the iterate sequence against; the cost
 * function is pinned against the reference's own functor source (oracle/_ref, ref_functor.cpp) and against the
 * reference's Ceres output, the end result against that output, the candidate search against the reference's
 * committed file (pgo_oracle.h).
 */
#include "pgo_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------
 * Forward-mode jets: what ceres::AutoDiffCostFunction<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4>
 * (REF/include/PoseGraph3dError.h:59) evaluates the functor with -- ceres::Jet<double, 14>,
 * one derivative slot per ambient parameter: p_a[0..2] q_a[3..6] p_b[7..9] q_b[10..13].
 * ---------------------------------------------------------------------------------------- */
#define NJ 14
typedef struct { double a; double v[NJ]; } jet;

static jet j_const(double c) { jet r; r.a = c; memset(r.v, 0, sizeof r.v); return r; }
static jet j_var(double c, int k) { jet r = j_const(c); r.v[k] = 1.0; return r; }
static jet j_add(jet x, jet y) { jet r; int i; r.a = x.a + y.a; for (i = 0; i < NJ; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
static jet j_sub(jet x, jet y) { jet r; int i; r.a = x.a - y.a; for (i = 0; i < NJ; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
static jet j_neg(jet x) { jet r; int i; r.a = -x.a; for (i = 0; i < NJ; ++i) r.v[i] = -x.v[i]; return r; }
static jet j_mul(jet x, jet y) { jet r; int i; r.a = x.a * y.a; for (i = 0; i < NJ; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
static jet j_scale(double s, jet x) { jet r; int i; r.a = s * x.a; for (i = 0; i < NJ; ++i) r.v[i] = s * x.v[i]; return r; }

/* Eigen::Quaternion stored as coeffs() = x,y,z,w (REF/include/types.h:18). */
typedef struct { jet x, y, z, w; } jquat;
typedef struct { jet x, y, z; } jvec3;

/* Eigen::QuaternionBase::conjugate() */
static jquat jq_conj(jquat q) { jquat r; r.x = j_neg(q.x); r.y = j_neg(q.y); r.z = j_neg(q.z); r.w = q.w; return r; }

/* Eigen::internal::quat_product (Hamilton product a*b) */
static jquat jq_mul(jquat a, jquat b) {
  jquat r;
  r.w = j_sub(j_sub(j_sub(j_mul(a.w, b.w), j_mul(a.x, b.x)), j_mul(a.y, b.y)), j_mul(a.z, b.z));
  r.x = j_sub(j_add(j_add(j_mul(a.w, b.x), j_mul(a.x, b.w)), j_mul(a.y, b.z)), j_mul(a.z, b.y));
  r.y = j_sub(j_add(j_add(j_mul(a.w, b.y), j_mul(a.y, b.w)), j_mul(a.z, b.x)), j_mul(a.x, b.z));
  r.z = j_sub(j_add(j_add(j_mul(a.w, b.z), j_mul(a.z, b.w)), j_mul(a.x, b.y)), j_mul(a.y, b.x));
  return r;
}

static jvec3 jv_cross(jvec3 a, jvec3 b) {
  jvec3 r;
  r.x = j_sub(j_mul(a.y, b.z), j_mul(a.z, b.y));
  r.y = j_sub(j_mul(a.z, b.x), j_mul(a.x, b.z));
  r.z = j_sub(j_mul(a.x, b.y), j_mul(a.y, b.x));
  return r;
}

/* Eigen::QuaternionBase::_transformVector: uv = vec x v; uv += uv; v + w*uv + vec x uv.
 * (No normalisation: the derivative w.r.t. a non-unit q is that of this exact polynomial.) */
static jvec3 jq_rotate(jquat q, jvec3 v) {
  jvec3 u, uv, c, r;
  u.x = q.x; u.y = q.y; u.z = q.z;
  uv = jv_cross(u, v);
  uv.x = j_add(uv.x, uv.x); uv.y = j_add(uv.y, uv.y); uv.z = j_add(uv.z, uv.z);
  c = jv_cross(u, uv);
  r.x = j_add(j_add(v.x, j_mul(q.w, uv.x)), c.x);
  r.y = j_add(j_add(v.y, j_mul(q.w, uv.y)), c.y);
  r.z = j_add(j_add(v.z, j_mul(q.w, uv.z)), c.z);
  return r;
}

/* PoseGraph3dErrorTerm::operator() with T = Jet  (REF/include/PoseGraph3dError.h:21-54).
 * Outputs: res[6] (sqrt-information applied, no loss yet), jamb[6][14] ambient Jacobian. */
static void edge_functor_jet(const double* pa, const double* pb, const double* meas,
                             const double* S, double* res, double* jamb) {
  jvec3 p_a, p_b, d, p_ab;
  jquat q_a, q_b, q_m, q_a_inv, q_ab, dq;
  jet r[6], o[6];
  int i, k;
  p_a.x = j_var(pa[0], 0); p_a.y = j_var(pa[1], 1); p_a.z = j_var(pa[2], 2);
  q_a.x = j_var(pa[3], 3); q_a.y = j_var(pa[4], 4); q_a.z = j_var(pa[5], 5); q_a.w = j_var(pa[6], 6);
  p_b.x = j_var(pb[0], 7); p_b.y = j_var(pb[1], 8); p_b.z = j_var(pb[2], 9);
  q_b.x = j_var(pb[3], 10); q_b.y = j_var(pb[4], 11); q_b.z = j_var(pb[5], 12); q_b.w = j_var(pb[6], 13);
  q_m.x = j_const(meas[3]); q_m.y = j_const(meas[4]); q_m.z = j_const(meas[5]); q_m.w = j_const(meas[6]);

  q_a_inv = jq_conj(q_a);                       /* :33 */
  q_ab = jq_mul(q_a_inv, q_b);                  /* :34 */
  d.x = j_sub(p_b.x, p_a.x); d.y = j_sub(p_b.y, p_a.y); d.z = j_sub(p_b.z, p_a.z);
  p_ab = jq_rotate(q_a_inv, d);                 /* :37 */
  dq = jq_mul(q_m, jq_conj(q_ab));              /* :40-41 */
  r[0] = j_sub(p_ab.x, j_const(meas[0]));       /* :47-48 */
  r[1] = j_sub(p_ab.y, j_const(meas[1]));
  r[2] = j_sub(p_ab.z, j_const(meas[2]));
  r[3] = j_scale(2.0, dq.x);                    /* :49 */
  r[4] = j_scale(2.0, dq.y);
  r[5] = j_scale(2.0, dq.z);
  for (i = 0; i < 6; ++i) {                     /* :52 residuals.applyOnTheLeft(sqrt_information) */
    o[i] = j_const(0.0);
    for (k = 0; k < 6; ++k) o[i] = j_add(o[i], j_scale(S[i * 6 + k], r[k]));
  }
  for (i = 0; i < 6; ++i) {
    res[i] = o[i].a;
    for (k = 0; k < NJ; ++k) jamb[i * NJ + k] = o[i].v[k];
  }
}

/* Same functor with T = double (cost-only evaluations). */
static void edge_functor_double(const double* pa, const double* pb, const double* meas,
                                const double* S, double* res) {
  /* q_a_inverse = conj(q_a) */
  const double ax = -pa[3], ay = -pa[4], az = -pa[5], aw = pa[6];
  const double bx = pb[3], by = pb[4], bz = pb[5], bw = pb[6];
  /* q_ab = q_a_inverse * q_b */
  const double ew = aw * bw - ax * bx - ay * by - az * bz;
  const double ex = aw * bx + ax * bw + ay * bz - az * by;
  const double ey = aw * by + ay * bw + az * bx - ax * bz;
  const double ez = aw * bz + az * bw + ax * by - ay * bx;
  const double dx = pb[0] - pa[0], dy = pb[1] - pa[1], dz = pb[2] - pa[2];
  double ux = ay * dz - az * dy, uy = az * dx - ax * dz, uz = ax * dy - ay * dx;
  double r[6];
  int i, k;
  ux += ux; uy += uy; uz += uz;
  r[0] = dx + aw * ux + (ay * uz - az * uy) - meas[0];
  r[1] = dy + aw * uy + (az * ux - ax * uz) - meas[1];
  r[2] = dz + aw * uz + (ax * uy - ay * ux) - meas[2];
  {
    /* delta_q = q_meas * conj(q_ab) */
    const double mx = meas[3], my = meas[4], mz = meas[5], mw = meas[6];
    const double cx = -ex, cy = -ey, cz = -ez, cw = ew;
    r[3] = 2.0 * (mw * cx + mx * cw + my * cz - mz * cy);
    r[4] = 2.0 * (mw * cy + my * cw + mz * cx - mx * cz);
    r[5] = 2.0 * (mw * cz + mz * cw + mx * cy - my * cx);
  }
  for (i = 0; i < 6; ++i) {
    double s = 0.0;
    for (k = 0; k < 6; ++k) s += S[i * 6 + k] * r[k];
    res[i] = s;
  }
}

/* ceres::EigenQuaternionParameterization::ComputeJacobian (4x3, rows x,y,z,w).
 * The reference attaches it to every q block: pose_graph_ceres_plus_finial.cpp:464-465,487-490. */
static void eigen_quat_plus_jacobian(const double* q, double* P) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  P[0] = w;  P[1] = z;   P[2] = -y;
  P[3] = -z; P[4] = w;   P[5] = x;
  P[6] = y;  P[7] = -x;  P[8] = w;
  P[9] = -x; P[10] = -y; P[11] = -z;
}

/* ceres::EigenQuaternionParameterization::Plus: q+ = Quaternion(cos|d|, sin|d|/|d| * d) * q,
 * and the default (identity) parameterization of the p block: p+ = p + d. */
static void pose_plus(const double* x, const double* delta, double* out) {
  const double n = sqrt(delta[3] * delta[3] + delta[4] * delta[4] + delta[5] * delta[5]);
  out[0] = x[0] + delta[0]; out[1] = x[1] + delta[1]; out[2] = x[2] + delta[2];
  if (n > 0.0) {
    const double s = sin(n) / n;
    const double ax = s * delta[3], ay = s * delta[4], az = s * delta[5], aw = cos(n);
    const double bx = x[3], by = x[4], bz = x[5], bw = x[6];
    out[6] = aw * bw - ax * bx - ay * by - az * bz;
    out[3] = aw * bx + ax * bw + ay * bz - az * by;
    out[4] = aw * by + ay * bw + az * bx - ax * bz;
    out[5] = aw * bz + az * bw + ax * by - ay * bx;
  } else {
    out[3] = x[3]; out[4] = x[4]; out[5] = x[5]; out[6] = x[6];
  }
}

void oracle_plus(int n_poses, const double* poses, const double* delta, double* out) {
  int i;
  for (i = 0; i < n_poses; ++i) pose_plus(poses + 7 * i, delta + 6 * i, out + 7 * i);
}

/* ceres::HuberLoss::Evaluate / CauchyLoss::Evaluate / NULL loss. rho = {rho, rho', rho''}.
 * The reference uses new ceres::HuberLoss(1.0): pose_graph_ceres_plus_finial.cpp:463. */
static void loss_evaluate(int type, double a, double s, double* rho) {
  if (type == ORACLE_LOSS_HUBER) {
    const double b = a * a;
    if (s > b) {
      const double r = sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = a / r; if (rho[1] < DBL_MIN) rho[1] = DBL_MIN;
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else if (type == ORACLE_LOSS_CAUCHY) {
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c, inv = 1.0 / sum;
    rho[0] = b * log(sum);
    rho[1] = inv; if (rho[1] < DBL_MIN) rho[1] = DBL_MIN;
    rho[2] = -c * (inv * inv);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

/* ceres::internal::Corrector (corrector.cc): rescale residual and Jacobian so that the
 * Gauss-Newton model of 1/2 rho(|r|^2) is J^T J / J^T r of the corrected quantities. */
static void corrector_apply(double sq_norm, const double* rho, double* res, double* Ja, double* Jb) {
  const double sqrt_rho1 = sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  int r, c, blk;
  if (sq_norm == 0.0 || rho[2] <= 0.0) {
    residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0;
  } else {
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - sqrt(D);
    residual_scaling = sqrt_rho1 / (1.0 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  for (blk = 0; blk < 2; ++blk) {
    double* J = blk ? Jb : Ja;
    if (!J) continue;
    if (alpha_sq_norm == 0.0) {
      for (r = 0; r < 36; ++r) J[r] *= sqrt_rho1;
    } else {
      for (c = 0; c < 6; ++c) {
        double rtj = 0.0;
        for (r = 0; r < 6; ++r) rtj += J[r * 6 + c] * res[r];
        for (r = 0; r < 6; ++r) J[r * 6 + c] = sqrt_rho1 * (J[r * 6 + c] - alpha_sq_norm * res[r] * rtj);
      }
    }
  }
  for (r = 0; r < 6; ++r) res[r] *= residual_scaling;
}

/* ceres::internal::ResidualBlock::Evaluate for one edge: functor (autodiff) -> project the q
 * columns through the local parameterization -> loss correction.  Ja/Jb: row-major 6x6 local. */
static double edge_evaluate(const double* pa, const double* pb, const double* meas, const double* S,
                            int loss_type, double loss_a, double* res, double* Ja, double* Jb) {
  double rho[3], sq = 0.0;
  int r, c, k;
  if (Ja || Jb) {
    double jamb[6 * NJ], Pa[12], Pb[12];
    edge_functor_jet(pa, pb, meas, S, res, jamb);
    eigen_quat_plus_jacobian(pa + 3, Pa);
    eigen_quat_plus_jacobian(pb + 3, Pb);
    for (r = 0; r < 6; ++r) {
      for (c = 0; c < 3; ++c) {
        double sa = 0.0, sb = 0.0;
        if (Ja) Ja[r * 6 + c] = jamb[r * NJ + c];
        if (Jb) Jb[r * 6 + c] = jamb[r * NJ + 7 + c];
        for (k = 0; k < 4; ++k) {
          sa += jamb[r * NJ + 3 + k] * Pa[k * 3 + c];
          sb += jamb[r * NJ + 10 + k] * Pb[k * 3 + c];
        }
        if (Ja) Ja[r * 6 + 3 + c] = sa;
        if (Jb) Jb[r * 6 + 3 + c] = sb;
      }
    }
  } else {
    edge_functor_double(pa, pb, meas, S, res);
  }
  for (r = 0; r < 6; ++r) sq += res[r] * res[r];
  loss_evaluate(loss_type, loss_a, sq, rho);
  if (loss_type != ORACLE_LOSS_TRIVIAL) corrector_apply(sq, rho, res, Ja, Jb);
  return 0.5 * rho[0];
}

/* ceres::Problem::Evaluate / ProgramEvaluator::Evaluate over all residual blocks. */
static int g_oracle_threads = 1;
/* Threads for the per-edge evaluation loop (Ceres' Solver::Options::num_threads; the reference
 * leaves it at its default of 1). The sparse Cholesky stays serial, as in Ceres. */
void oracle_set_num_threads(int n) { g_oracle_threads = n > 0 ? n : 1; }
int oracle_get_num_threads(void) { return g_oracle_threads; }

/* ceres::Problem::AddResidualBlock takes the loss function PER residual block (REF :513-517 passes one HuberLoss to every
 * block, but the surface allows any mixture): when set, edge e uses (type[e], a[e]) instead of the options' loss. */
static const int* g_edge_loss_type = NULL;
static const double* g_edge_loss_a = NULL;
static int g_edge_loss_n = 0;
void oracle_set_edge_losses(int n_edges, const int* type, const double* a) {
  g_edge_loss_type = type; g_edge_loss_a = a; g_edge_loss_n = (type && a) ? n_edges : 0;
}

/* pose_const: 0 = variable, 1 = SetParameterBlockConstant(p) and (q), 2 = p only, 3 = q only (REF :526-527 are two
 * separate calls).  A constant block's Jacobian columns are dropped, as Ceres' program preprocessing does. */
static void drop_constant_columns(double* J, int code) {
  int r, c;
  if (code == 1) { memset(J, 0, 36 * sizeof(double)); return; }
  if (code != 2 && code != 3) return;
  for (r = 0; r < 6; ++r) for (c = 0; c < 3; ++c) J[r * 6 + (code == 3 ? 3 : 0) + c] = 0.0;
}

int oracle_evaluate(int n_poses, const double* poses, const unsigned char* pose_const,
                    int n_edges, const int* edge_ids, const double* edge_meas,
                    const double* edge_sqrt_info, int loss_type, double loss_a,
                    double* cost, double* residuals, double* gradient, double* jac) {
  double total = 0.0;
  int e, r, c;
  const int need_j = (gradient != NULL) || (jac != NULL);
  double* jbuf = jac;
  double* rbuf = residuals;
  if (gradient) memset(gradient, 0, sizeof(double) * 6 * (size_t)n_poses);
  for (e = 0; e < n_edges; ++e) {
    const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
    if (a < 0 || a >= n_poses || b < 0 || b >= n_poses) return -1;
  }
  if (need_j && !jbuf) jbuf = (double*)malloc(sizeof(double) * 72 * (size_t)(n_edges + 1));
  if (gradient && !rbuf) rbuf = (double*)malloc(sizeof(double) * 6 * (size_t)(n_edges + 1));
#pragma omp parallel for schedule(static) reduction(+ : total) num_threads(g_oracle_threads) if (g_oracle_threads > 1)
  for (e = 0; e < n_edges; ++e) {
    const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
    double res[6];
    double* Ja = need_j ? jbuf + 72 * (size_t)e : NULL;
    double* Jb = need_j ? Ja + 36 : NULL;
    const int lt = (g_edge_loss_n == n_edges) ? g_edge_loss_type[e] : loss_type;
    const double la = (g_edge_loss_n == n_edges) ? g_edge_loss_a[e] : loss_a;
    total += edge_evaluate(poses + 7 * a, poses + 7 * b, edge_meas + 7 * e, edge_sqrt_info + 36 * e,
                           lt, la, res, Ja, Jb);
    if (need_j) {
      if (pose_const) { drop_constant_columns(Ja, pose_const[a]); drop_constant_columns(Jb, pose_const[b]); }
    }
    if (rbuf) memcpy(rbuf + 6 * (size_t)e, res, sizeof res);
  }
  if (gradient) {
    for (e = 0; e < n_edges; ++e) {
      const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
      const double* Ja = jbuf + 72 * (size_t)e; const double* Jb = Ja + 36; const double* res = rbuf + 6 * (size_t)e;
      for (c = 0; c < 6; ++c) {
        double ga = 0.0, gb = 0.0;
        for (r = 0; r < 6; ++r) { ga += Ja[r * 6 + c] * res[r]; gb += Jb[r * 6 + c] * res[r]; }
        gradient[6 * a + c] += ga; gradient[6 * b + c] += gb;
      }
    }
  }
  if (jbuf != jac) free(jbuf);
  if (rbuf != residuals) free(rbuf);
  if (cost) *cost = total;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * SPARSE_NORMAL_CHOLESKY restated: form H = J^T J + D^T D on the pose graph's block pattern,
 * fill-reducing ordering (Ceres uses SuiteSparse/Eigen AMD; here exact minimum degree on the
 * block graph), block up-looking Cholesky (the 6x6-block form of CSparse cs_chol), two
 * triangular solves.  Constant poses are removed from the system, as Ceres' program
 * preprocessing does.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int n_poses, n_edges, nb;
  int* var_of_pose;   /* pose -> variable block (original numbering) or -1 */
  int* pose_of_var;
  int* perm;          /* new -> old variable */
  int* iperm;         /* old -> new */
  int* Ap; int* Ai; double* Ax;   /* upper block CSC of permuted H, blocks row-major 6x6 */
  int* e_slot;        /* [E][3]: diag a, diag b, offdiag; -1 when the pose is constant */
  unsigned char* e_tr;/* offdiag stored as (new b, new a): accumulate Jb^T Ja instead of Ja^T Jb */
  int* parent;
  int* Lp; int* Li; double* Lx;   /* lower block CSC, diagonal block first */
  int* colfill;       /* running fill pointer per column during numeric */
  int* stack; int* mark; double* xw;
  long long lnz;
} chol_t;

static int cmp_ll(const void* a, const void* b) {
  const long long x = *(const long long*)a, y = *(const long long*)b;
  return (x > y) - (x < y);
}

/* exact minimum-degree ordering on the elimination graph (lazy binary heap, sorted adjacency). */
typedef struct { int deg, node; } heap_ent;
static void heap_push(heap_ent** h, int* n, int* cap, int deg, int node) {
  int i;
  if (*n == *cap) { *cap = *cap * 2 + 16; *h = (heap_ent*)realloc(*h, sizeof(heap_ent) * (size_t)*cap); }
  i = (*n)++;
  while (i > 0) {
    const int p = (i - 1) / 2;
    if ((*h)[p].deg < deg || ((*h)[p].deg == deg && (*h)[p].node < node)) break;
    (*h)[i] = (*h)[p]; i = p;
  }
  (*h)[i].deg = deg; (*h)[i].node = node;
}
static heap_ent heap_pop(heap_ent* h, int* n) {
  heap_ent top = h[0], last = h[--(*n)];
  int i = 0;
  for (;;) {
    int c = 2 * i + 1;
    if (c >= *n) break;
    if (c + 1 < *n && (h[c + 1].deg < h[c].deg || (h[c + 1].deg == h[c].deg && h[c + 1].node < h[c].node))) ++c;
    if (last.deg < h[c].deg || (last.deg == h[c].deg && last.node < h[c].node)) break;
    h[i] = h[c]; i = c;
  }
  h[i] = last;
  return top;
}

static void min_degree_order(int nb, const int* ep /*pairs*/, int npairs, int* perm) {
  int** adj = (int**)calloc((size_t)nb, sizeof(int*));
  int* len = (int*)calloc((size_t)nb, sizeof(int));
  int* cap = (int*)calloc((size_t)nb, sizeof(int));
  unsigned char* done = (unsigned char*)calloc((size_t)nb, 1);
  int* tmp = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  heap_ent* heap = NULL; int hn = 0, hcap = 0;
  int i, k, out = 0;
  long long* keys = (long long*)malloc(sizeof(long long) * (size_t)(2 * npairs + 1));
  int nk = 0;
  for (i = 0; i < npairs; ++i) {
    const int a = ep[2 * i], b = ep[2 * i + 1];
    if (a == b) continue;
    keys[nk++] = (long long)a * nb + b; keys[nk++] = (long long)b * nb + a;
  }
  qsort(keys, (size_t)nk, sizeof(long long), cmp_ll);
  for (i = 0; i < nk; ++i) if (i == 0 || keys[i] != keys[i - 1]) len[keys[i] / nb]++;
  for (i = 0; i < nb; ++i) { cap[i] = len[i] + 4; adj[i] = (int*)malloc(sizeof(int) * (size_t)cap[i]); len[i] = 0; }
  for (i = 0; i < nk; ++i) if (i == 0 || keys[i] != keys[i - 1]) { const int a = (int)(keys[i] / nb); adj[a][len[a]++] = (int)(keys[i] % nb); }
  free(keys);
  for (i = 0; i < nb; ++i) heap_push(&heap, &hn, &hcap, len[i], i);
  while (out < nb) {
    heap_ent t = heap_pop(heap, &hn);
    const int v = t.node;
    int nv; int* Nv;
    if (done[v] || t.deg != len[v]) continue;   /* stale entry */
    done[v] = 1; perm[out++] = v;
    Nv = adj[v]; nv = len[v];
    for (k = 0; k < nv; ++k) {
      const int u = Nv[k];
      /* adj[u] = (adj[u] U Nv) \ {u, v}; both sorted */
      int ia = 0, ib = 0, m = 0;
      const int* A = adj[u]; const int na = len[u];
      while (ia < na || ib < nv) {
        int x;
        if (ib >= nv || (ia < na && A[ia] < Nv[ib])) x = A[ia++];
        else if (ia >= na || Nv[ib] < A[ia]) x = Nv[ib++];
        else { x = A[ia]; ++ia; ++ib; }
        if (x != u && x != v) tmp[m++] = x;
      }
      if (m > cap[u]) { cap[u] = m + m / 2 + 4; free(adj[u]); adj[u] = (int*)malloc(sizeof(int) * (size_t)cap[u]); }
      memcpy(adj[u], tmp, sizeof(int) * (size_t)m);
      len[u] = m;
      heap_push(&heap, &hn, &hcap, m, u);
    }
    free(adj[v]); adj[v] = NULL; len[v] = 0;
  }
  for (i = 0; i < nb; ++i) free(adj[i]);
  free(adj); free(len); free(cap); free(done); free(tmp); free(heap);
}

static void chol_free(chol_t* C) {
  if (!C) return;
  free(C->var_of_pose); free(C->pose_of_var); free(C->perm); free(C->iperm);
  free(C->Ap); free(C->Ai); free(C->Ax); free(C->e_slot); free(C->e_tr); free(C->parent);
  free(C->Lp); free(C->Li); free(C->Lx); free(C->colfill); free(C->stack); free(C->mark); free(C->xw);
  free(C);
}

/* cs_ereach on the block pattern: nonzero pattern of row k of L, topologically ordered. */
static int ereach(const chol_t* C, int k, int* s, int* mark) {
  int top = C->nb, p;
  mark[k] = k;
  for (p = C->Ap[k]; p < C->Ap[k + 1]; ++p) {
    int i = C->Ai[p], len = 0;
    if (i > k) continue;
    for (; mark[i] != k; i = C->parent[i]) { s[len++] = i; mark[i] = k; }
    while (len > 0) s[--top] = s[--len];
  }
  return top;
}

static chol_t* chol_analyze(int n_poses, const unsigned char* pose_const, int n_edges,
                            const int* edge_ids, int ordering) {
  chol_t* C = (chol_t*)calloc(1, sizeof(chol_t));
  int i, e, k, nb = 0, npairs = 0;
  int* pairs; long long* keys; int nk = 0, nnzb = 0;
  unsigned char* used = (unsigned char*)calloc((size_t)n_poses, 1);
  C->n_poses = n_poses; C->n_edges = n_edges;
  C->var_of_pose = (int*)malloc(sizeof(int) * (size_t)n_poses);
  for (e = 0; e < n_edges; ++e) { used[edge_ids[2 * e]] = 1; used[edge_ids[2 * e + 1]] = 1; }
  for (i = 0; i < n_poses; ++i) C->var_of_pose[i] = (used[i] && !(pose_const && pose_const[i] == 1)) ? nb++ : -1;
  free(used);
  C->nb = nb;
  C->pose_of_var = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  for (i = 0; i < n_poses; ++i) if (C->var_of_pose[i] >= 0) C->pose_of_var[C->var_of_pose[i]] = i;
  pairs = (int*)malloc(sizeof(int) * 2 * (size_t)(n_edges + 1));
  for (e = 0; e < n_edges; ++e) {
    const int va = C->var_of_pose[edge_ids[2 * e]], vb = C->var_of_pose[edge_ids[2 * e + 1]];
    if (va >= 0 && vb >= 0 && va != vb) { pairs[2 * npairs] = va; pairs[2 * npairs + 1] = vb; ++npairs; }
  }
  C->perm = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  C->iperm = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  if (ordering == 1) min_degree_order(nb, pairs, npairs, C->perm);
  else for (i = 0; i < nb; ++i) C->perm[i] = i;
  for (i = 0; i < nb; ++i) C->iperm[C->perm[i]] = i;
  /* upper block pattern, column-major keys col*nb+row */
  keys = (long long*)malloc(sizeof(long long) * (size_t)(nb + npairs + 1));
  for (i = 0; i < nb; ++i) keys[nk++] = (long long)i * nb + i;
  for (i = 0; i < npairs; ++i) {
    int a = C->iperm[pairs[2 * i]], b = C->iperm[pairs[2 * i + 1]];
    if (a > b) { int t = a; a = b; b = t; }
    keys[nk++] = (long long)b * nb + a;
  }
  free(pairs);
  qsort(keys, (size_t)nk, sizeof(long long), cmp_ll);
  for (i = 0; i < nk; ++i) if (i == 0 || keys[i] != keys[i - 1]) keys[nnzb++] = keys[i];
  C->Ap = (int*)calloc((size_t)nb + 1, sizeof(int));
  C->Ai = (int*)malloc(sizeof(int) * (size_t)(nnzb + 1));
  C->Ax = (double*)calloc((size_t)nnzb * 36 + 1, sizeof(double));
  for (i = 0; i < nnzb; ++i) { C->Ap[keys[i] / nb + 1]++; C->Ai[i] = (int)(keys[i] % nb); }
  for (i = 0; i < nb; ++i) C->Ap[i + 1] += C->Ap[i];
  /* edge -> slot */
  C->e_slot = (int*)malloc(sizeof(int) * 3 * (size_t)(n_edges + 1));
  C->e_tr = (unsigned char*)calloc((size_t)n_edges + 1, 1);
  for (e = 0; e < n_edges; ++e) {
    const int va = C->var_of_pose[edge_ids[2 * e]], vb = C->var_of_pose[edge_ids[2 * e + 1]];
    const int na = va >= 0 ? C->iperm[va] : -1, nbb = vb >= 0 ? C->iperm[vb] : -1;
    int* s = C->e_slot + 3 * e;
    s[0] = s[1] = s[2] = -1;
    if (na >= 0) { long long key = (long long)na * nb + na; s[0] = (int)((long long*)bsearch(&key, keys, (size_t)nnzb, sizeof(long long), cmp_ll) - keys); }
    if (nbb >= 0) { long long key = (long long)nbb * nb + nbb; s[1] = (int)((long long*)bsearch(&key, keys, (size_t)nnzb, sizeof(long long), cmp_ll) - keys); }
    if (na >= 0 && nbb >= 0 && na != nbb) {
      const int lo = na < nbb ? na : nbb, hi = na < nbb ? nbb : na;
      long long key = (long long)hi * nb + lo;
      s[2] = (int)((long long*)bsearch(&key, keys, (size_t)nnzb, sizeof(long long), cmp_ll) - keys);
      C->e_tr[e] = (unsigned char)(na > nbb);
    }
  }
  free(keys);
  /* elimination tree (cs_etree) */
  C->parent = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  {
    int* anc = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
    for (k = 0; k < nb; ++k) {
      int p;
      C->parent[k] = -1; anc[k] = -1;
      for (p = C->Ap[k]; p < C->Ap[k + 1]; ++p) {
        int ii = C->Ai[p];
        while (ii != -1 && ii < k) {
          const int inext = anc[ii];
          anc[ii] = k;
          if (inext == -1) C->parent[ii] = k;
          ii = inext;
        }
      }
    }
    free(anc);
  }
  /* column counts by a symbolic up-looking pass */
  C->stack = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  C->mark = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  C->Lp = (int*)calloc((size_t)nb + 1, sizeof(int));
  for (k = 0; k < nb; ++k) C->mark[k] = -1;
  for (k = 0; k < nb; ++k) {
    const int top = ereach(C, k, C->stack, C->mark);
    for (i = top; i < nb; ++i) C->Lp[C->stack[i] + 1]++;
    C->Lp[k + 1]++;  /* diagonal */
  }
  for (k = 0; k < nb; ++k) C->Lp[k + 1] += C->Lp[k];
  C->lnz = C->Lp[nb];
  C->Li = (int*)malloc(sizeof(int) * (size_t)(C->lnz + 1));
  C->Lx = (double*)malloc(sizeof(double) * 36 * (size_t)(C->lnz + 1));
  C->colfill = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
  C->xw = (double*)calloc((size_t)nb * 36 + 36, sizeof(double));
  return C;
}

/* H(upper, permuted) = sum_e J_e^T J_e + diag(d).  jac: [E][2][36] local blocks, d: [6N]. */
static void chol_assemble(chol_t* C, const int* edge_ids, const double* jac, const double* d) {
  int e, r, c, k, v;
  const int nnzb = C->Ap[C->nb];
  memset(C->Ax, 0, sizeof(double) * 36 * (size_t)nnzb);
  for (e = 0; e < C->n_edges; ++e) {
    const int* s = C->e_slot + 3 * e;
    const double* Ja = jac + 72 * (size_t)e;
    const double* Jb = Ja + 36;
    if (s[0] >= 0) { double* H = C->Ax + 36 * (size_t)s[0]; for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) { double t = 0; for (k = 0; k < 6; ++k) t += Ja[k * 6 + r] * Ja[k * 6 + c]; H[r * 6 + c] += t; } }
    if (s[1] >= 0) { double* H = C->Ax + 36 * (size_t)s[1]; for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) { double t = 0; for (k = 0; k < 6; ++k) t += Jb[k * 6 + r] * Jb[k * 6 + c]; H[r * 6 + c] += t; } }
    if (s[2] >= 0) {
      double* H = C->Ax + 36 * (size_t)s[2];
      const double* X = C->e_tr[e] ? Jb : Ja;   /* row side */
      const double* Y = C->e_tr[e] ? Ja : Jb;   /* column side */
      for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) { double t = 0; for (k = 0; k < 6; ++k) t += X[k * 6 + r] * Y[k * 6 + c]; H[r * 6 + c] += t; }
    }
  }
  if (d) {
    for (v = 0; v < C->nb; ++v) {
      const int nv = C->iperm[v], pose = C->pose_of_var[v];
      int p;
      for (p = C->Ap[nv]; p < C->Ap[nv + 1]; ++p) if (C->Ai[p] == nv) { for (r = 0; r < 6; ++r) C->Ax[36 * (size_t)p + r * 7] += d[6 * pose + r]; break; }
    }
  }
}

/* dense 6x6 lower Cholesky in place (row-major); upper part zeroed. */
static int chol6(double* A) {
  int i, j, k;
  for (j = 0; j < 6; ++j) {
    double s = A[j * 6 + j];
    for (k = 0; k < j; ++k) s -= A[j * 6 + k] * A[j * 6 + k];
    if (!(s > 0.0)) return -1;
    s = sqrt(s); A[j * 6 + j] = s;
    for (i = j + 1; i < 6; ++i) {
      double t = A[i * 6 + j];
      for (k = 0; k < j; ++k) t -= A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = t / s;
    }
    for (i = 0; i < j; ++i) A[i * 6 + j] = 0.0;
  }
  return 0;
}

/* block up-looking numeric factorisation (block form of CSparse cs_chol). */
static int chol_factor(chol_t* C) {
  const int nb = C->nb;
  int k, i, p, r, c, q;
  double* x = C->xw;
  for (k = 0; k < nb; ++k) { C->colfill[k] = C->Lp[k]; C->mark[k] = -1; }
  for (k = 0; k < nb; ++k) {
    double dkk[36];
    const int top = ereach(C, k, C->stack, C->mark);
    memset(dkk, 0, sizeof dkk);
    for (p = C->Ap[k]; p < C->Ap[k + 1]; ++p) {
      const int ii = C->Ai[p];
      if (ii < k) memcpy(x + 36 * (size_t)ii, C->Ax + 36 * (size_t)p, 36 * sizeof(double));
      else if (ii == k) memcpy(dkk, C->Ax + 36 * (size_t)p, sizeof dkk);
    }
    for (i = top; i < nb; ++i) {
      const int ci = C->stack[i];
      const double* Lii = C->Lx + 36 * (size_t)C->Lp[ci];
      double* xi = x + 36 * (size_t)ci;
      double X[36], Xt[36];
      /* X = Lii^{-1} * xi  (forward substitution per column) */
      for (c = 0; c < 6; ++c) {
        for (r = 0; r < 6; ++r) {
          double t = xi[r * 6 + c];
          for (q = 0; q < r; ++q) t -= Lii[r * 6 + q] * X[q * 6 + c];
          X[r * 6 + c] = t / Lii[r * 6 + r];
        }
      }
      memset(xi, 0, 36 * sizeof(double));
      for (p = C->Lp[ci] + 1; p < C->colfill[ci]; ++p) {
        const double* Lri = C->Lx + 36 * (size_t)p;
        double* xr = x + 36 * (size_t)C->Li[p];
        for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) {
          double t = 0; for (q = 0; q < 6; ++q) t += Lri[r * 6 + q] * X[q * 6 + c];
          xr[r * 6 + c] -= t;
        }
      }
      for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) {
        double t = 0; for (q = 0; q < 6; ++q) t += X[q * 6 + r] * X[q * 6 + c];
        dkk[r * 6 + c] -= t; Xt[r * 6 + c] = X[c * 6 + r];
      }
      p = C->colfill[ci]++;
      C->Li[p] = k; memcpy(C->Lx + 36 * (size_t)p, Xt, sizeof Xt);
    }
    if (chol6(dkk)) return -1;
    p = C->colfill[k]++;
    C->Li[p] = k; memcpy(C->Lx + 36 * (size_t)p, dkk, sizeof dkk);
  }
  return 0;
}

/* y = H^{-1} rhs; rhs and y are [6 * n_poses], constant poses get 0. */
static void chol_solve(const chol_t* C, const double* rhs, double* y) {
  const int nb = C->nb;
  double* w = (double*)malloc(sizeof(double) * 6 * (size_t)(nb + 1));
  int j, p, r, q, v;
  for (v = 0; v < nb; ++v) memcpy(w + 6 * (size_t)C->iperm[v], rhs + 6 * (size_t)C->pose_of_var[v], 6 * sizeof(double));
  for (j = 0; j < nb; ++j) {           /* L z = w */
    const double* Ljj = C->Lx + 36 * (size_t)C->Lp[j];
    double* wj = w + 6 * j;
    for (r = 0; r < 6; ++r) { double t = wj[r]; for (q = 0; q < r; ++q) t -= Ljj[r * 6 + q] * wj[q]; wj[r] = t / Ljj[r * 6 + r]; }
    for (p = C->Lp[j] + 1; p < C->Lp[j + 1]; ++p) {
      const double* L = C->Lx + 36 * (size_t)p; double* wr = w + 6 * (size_t)C->Li[p];
      for (r = 0; r < 6; ++r) { double t = 0; for (q = 0; q < 6; ++q) t += L[r * 6 + q] * wj[q]; wr[r] -= t; }
    }
  }
  for (j = nb - 1; j >= 0; --j) {      /* L^T y = z */
    const double* Ljj = C->Lx + 36 * (size_t)C->Lp[j];
    double* wj = w + 6 * j;
    for (p = C->Lp[j] + 1; p < C->Lp[j + 1]; ++p) {
      const double* L = C->Lx + 36 * (size_t)p; const double* wr = w + 6 * (size_t)C->Li[p];
      for (q = 0; q < 6; ++q) { double t = 0; for (r = 0; r < 6; ++r) t += L[r * 6 + q] * wr[r]; wj[q] -= t; }
    }
    for (r = 5; r >= 0; --r) { double t = wj[r]; for (q = r + 1; q < 6; ++q) t -= Ljj[q * 6 + r] * wj[q]; wj[r] = t / Ljj[r * 6 + r]; }
  }
  memset(y, 0, sizeof(double) * 6 * (size_t)C->n_poses);
  for (v = 0; v < nb; ++v) memcpy(y + 6 * (size_t)C->pose_of_var[v], w + 6 * (size_t)C->iperm[v], 6 * sizeof(double));
  free(w);
}

int oracle_normal_solve(int n_poses, const unsigned char* pose_const, int n_edges,
                        const int* edge_ids, const double* jac, const double* d,
                        const double* rhs, double* y, int ordering) {
  chol_t* C = chol_analyze(n_poses, pose_const, n_edges, edge_ids, ordering);
  int rc;
  chol_assemble(C, edge_ids, jac, d);
  rc = chol_factor(C);
  if (rc == 0) chol_solve(C, rhs, y);
  chol_free(C);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * ceres::internal::TrustRegionMinimizer::Minimize + LevenbergMarquardtStrategy, as of
 * ceres-solver 1.13 (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc), with the
 * options the reference sets at pose_graph_ceres_plus_finial.cpp:503-505 and Ceres defaults
 * for everything else.
 * ---------------------------------------------------------------------------------------- */
void oracle_default_options(oracle_options* o) {
  o->max_num_iterations = 1000;   /* REF test/pose_graph_ceres_plus_finial.cpp:504 */
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->loss_type = ORACLE_LOSS_HUBER;  /* :463 */
  o->loss_a = 1.0;
  o->ordering = 1;
}

static void col_sq_norms(int n_poses, int n_edges, const int* ids, const double* jac, double* out) {
  int e, r, c;
  memset(out, 0, sizeof(double) * 6 * (size_t)n_poses);
  for (e = 0; e < n_edges; ++e) {
    const double* Ja = jac + 72 * (size_t)e; const double* Jb = Ja + 36;
    double* oa = out + 6 * (size_t)ids[2 * e]; double* ob = out + 6 * (size_t)ids[2 * e + 1];
    for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) { oa[c] += Ja[r * 6 + c] * Ja[r * 6 + c]; ob[c] += Jb[r * 6 + c] * Jb[r * 6 + c]; }
  }
}

static void scale_columns(int n_edges, const int* ids, double* jac, const double* scale) {
  int e, r, c;
  for (e = 0; e < n_edges; ++e) {
    double* Ja = jac + 72 * (size_t)e; double* Jb = Ja + 36;
    const double* sa = scale + 6 * (size_t)ids[2 * e]; const double* sb = scale + 6 * (size_t)ids[2 * e + 1];
    for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) { Ja[r * 6 + c] *= sa[c]; Jb[r * 6 + c] *= sb[c]; }
  }
}

/* gradient_max_norm / gradient_norm: |x - Plus(x, -g)| (trust_region_minimizer.cc,
 * EvaluateGradientAndJacobian). Only non-constant poses are part of the reduced program. */
static void projected_gradient_norms(int n_poses, const double* x, const unsigned char* active,
                                     const double* g, double* max_norm, double* l2) {
  double mx = 0.0, ss = 0.0;
  int i, k;
  for (i = 0; i < n_poses; ++i) {
    double ng[6], xp[7];
    if (!active[i]) continue;
    for (k = 0; k < 6; ++k) ng[k] = -g[6 * i + k];
    pose_plus(x + 7 * i, ng, xp);
    for (k = 0; k < 7; ++k) { const double dlt = fabs(x[7 * i + k] - xp[k]); if (dlt > mx) mx = dlt; ss += dlt * dlt; }
  }
  *max_norm = mx; *l2 = sqrt(ss);
}

static void log_iter(oracle_iteration* log, int cap, oracle_summary* s, const oracle_iteration* it) {
  if (log && s->num_iterations < cap) log[s->num_iterations] = *it;
  s->num_iterations++;
}

/* is component e (0..6) of pose k in a variable parameter block?  (x_norm runs over the variable blocks only) */
#define BLOCK_VARIABLE(pc, k, e) (!(pc) || (pc)[k] == 0 || ((pc)[k] == 2 && (e) >= 3) || ((pc)[k] == 3 && (e) < 3))
int oracle_solve(int n_poses, double* poses, const unsigned char* pose_const,
                 int n_edges, const int* edge_ids, const double* edge_meas,
                 const double* edge_sqrt_info, const oracle_options* opt,
                 oracle_summary* summary, oracle_iteration* iter_log, int iter_log_cap) {
  const size_t nv = 6 * (size_t)n_poses;
  const double t_begin = now_s();
  double t0;
  double* residuals = (double*)malloc(sizeof(double) * 6 * (size_t)(n_edges + 1));
  double* jac = (double*)malloc(sizeof(double) * 72 * (size_t)(n_edges + 1));
  double* gradient = (double*)calloc(nv + 6, sizeof(double));
  double* scale = (double*)malloc(sizeof(double) * (nv + 6));
  double* diagonal = (double*)malloc(sizeof(double) * (nv + 6));
  double* lm_diag_sq = (double*)malloc(sizeof(double) * (nv + 6));
  double* rhs = (double*)malloc(sizeof(double) * (nv + 6));
  double* step = (double*)malloc(sizeof(double) * (nv + 6));
  double* delta = (double*)malloc(sizeof(double) * (nv + 6));
  double* cand = (double*)malloc(sizeof(double) * 7 * (size_t)(n_poses + 1));
  unsigned char* active = (unsigned char*)calloc((size_t)n_poses + 1, 1);
  chol_t* C;
  oracle_iteration it;
  double x_cost = 0.0, x_norm, radius, decrease_factor = 2.0;
  int reuse_diagonal = 0, num_consecutive_invalid = 0, iter = 0, rc = 0;
  size_t i; int e, k;

  memset(summary, 0, sizeof *summary);
  C = chol_analyze(n_poses, pose_const, n_edges, edge_ids, opt->ordering);
  for (k = 0; k < n_poses; ++k) active[k] = (unsigned char)(C->var_of_pose[k] >= 0);
  summary->factor_nnz_blocks = C->lnz;

  /* ---- IterationZero ---- */
  t0 = now_s();
  if (oracle_evaluate(n_poses, poses, pose_const, n_edges, edge_ids, edge_meas, edge_sqrt_info,
                      opt->loss_type, opt->loss_a, &x_cost, residuals, gradient, jac)) { rc = -1; goto done; }
  summary->time_jacobian_s += now_s() - t0; summary->num_jacobian_evals++;
  for (i = 0; i < nv; ++i) scale[i] = 1.0;
  if (opt->jacobi_scaling) {
    col_sq_norms(n_poses, n_edges, edge_ids, jac, scale);
    for (i = 0; i < nv; ++i) scale[i] = 1.0 / (1.0 + sqrt(scale[i]));
    scale_columns(n_edges, edge_ids, jac, scale);
  }
  x_norm = 0.0;
  for (k = 0; k < n_poses; ++k) if (active[k]) for (e = 0; e < 7; ++e) if (BLOCK_VARIABLE(pose_const, k, e)) x_norm += poses[7 * k + e] * poses[7 * k + e];
  x_norm = sqrt(x_norm);
  radius = opt->initial_trust_region_radius;
  memset(&it, 0, sizeof it);
  it.cost = x_cost; it.trust_region_radius = radius; it.step_is_valid = 0; it.step_is_successful = 0;
  projected_gradient_norms(n_poses, poses, active, gradient, &it.gradient_max_norm, &it.gradient_norm);
  summary->initial_cost = x_cost;
  log_iter(iter_log, iter_log_cap, summary, &it);

  for (;;) {
    double model_cost_change, cand_cost, step_norm, cost_change, relative_decrease;
    int step_valid = 1;
    /* FinalizeIterationAndCheckIfMinimizerCanContinue */
    if (iter >= opt->max_num_iterations) { summary->termination_type = ORACLE_NO_CONVERGENCE; snprintf(summary->message, sizeof summary->message, "Maximum number of iterations reached. Number of iterations: %d.", iter); break; }
    if (it.gradient_max_norm <= opt->gradient_tolerance) { summary->termination_type = ORACLE_CONVERGENCE; snprintf(summary->message, sizeof summary->message, "Gradient tolerance reached. Gradient max norm: %e <= %e", it.gradient_max_norm, opt->gradient_tolerance); break; }
    if (radius < opt->min_trust_region_radius) { summary->termination_type = ORACLE_CONVERGENCE; snprintf(summary->message, sizeof summary->message, "Minimum trust region radius reached."); break; }
    ++iter;
    {
      const double gmax = it.gradient_max_norm, gn = it.gradient_norm;
      memset(&it, 0, sizeof it);
      it.iteration = iter; it.gradient_max_norm = gmax; it.gradient_norm = gn;
    }

    /* ---- ComputeTrustRegionStep -> LevenbergMarquardtStrategy::ComputeStep ---- */
    t0 = now_s();
    if (!reuse_diagonal) {
      col_sq_norms(n_poses, n_edges, edge_ids, jac, diagonal);
      for (i = 0; i < nv; ++i) { double dd = diagonal[i]; if (dd < opt->min_lm_diagonal) dd = opt->min_lm_diagonal; if (dd > opt->max_lm_diagonal) dd = opt->max_lm_diagonal; diagonal[i] = dd; }
    }
    for (i = 0; i < nv; ++i) { const double l = sqrt(diagonal[i] / radius); lm_diag_sq[i] = l * l; }
    /* rhs = J^T r with the (scaled) Jacobian */
    memset(rhs, 0, sizeof(double) * nv);
    for (e = 0; e < n_edges; ++e) {
      const double* Ja = jac + 72 * (size_t)e; const double* Jb = Ja + 36; const double* r = residuals + 6 * (size_t)e;
      double* ra = rhs + 6 * (size_t)edge_ids[2 * e]; double* rb = rhs + 6 * (size_t)edge_ids[2 * e + 1];
      int rr, cc;
      for (cc = 0; cc < 6; ++cc) { double sa = 0, sb = 0; for (rr = 0; rr < 6; ++rr) { sa += Ja[rr * 6 + cc] * r[rr]; sb += Jb[rr * 6 + cc] * r[rr]; } ra[cc] += sa; rb[cc] += sb; }
    }
    chol_assemble(C, edge_ids, jac, lm_diag_sq);
    if (chol_factor(C) != 0) step_valid = 0;
    else {
      chol_solve(C, rhs, step);
      for (i = 0; i < nv; ++i) { step[i] = -step[i]; if (!isfinite(step[i])) step_valid = 0; }
    }
    reuse_diagonal = 1;
    summary->time_linear_solver_s += now_s() - t0;

    /* model_cost_change = -(J step)^T (r + J step / 2) */
    model_cost_change = 0.0;
    if (step_valid) {
      for (e = 0; e < n_edges; ++e) {
        const double* Ja = jac + 72 * (size_t)e; const double* Jb = Ja + 36; const double* r = residuals + 6 * (size_t)e;
        const double* sa = step + 6 * (size_t)edge_ids[2 * e]; const double* sb = step + 6 * (size_t)edge_ids[2 * e + 1];
        int rr, cc;
        for (rr = 0; rr < 6; ++rr) { double m = 0; for (cc = 0; cc < 6; ++cc) m += Ja[rr * 6 + cc] * sa[cc] + Jb[rr * 6 + cc] * sb[cc]; model_cost_change -= m * (r[rr] + m / 2.0); }
      }
      step_valid = model_cost_change > 0.0;
    }
    it.trust_region_radius = radius;
    if (!step_valid) {
      /* HandleInvalidStep */
      it.step_is_valid = 0; it.cost = x_cost;
      if (++num_consecutive_invalid >= opt->max_num_consecutive_invalid_steps) {
        summary->termination_type = ORACLE_FAILURE;
        snprintf(summary->message, sizeof summary->message, "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %d", opt->max_num_consecutive_invalid_steps);
        log_iter(iter_log, iter_log_cap, summary, &it);
        break;
      }
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;   /* StepIsInvalid */
      summary->num_unsuccessful_steps++;
      log_iter(iter_log, iter_log_cap, summary, &it);
      continue;
    }
    num_consecutive_invalid = 0;
    it.step_is_valid = 1;
    for (i = 0; i < nv; ++i) delta[i] = step[i] * scale[i];

    /* ---- ComputeCandidatePointAndEvaluateCost ---- */
    oracle_plus(n_poses, poses, delta, cand);
    for (k = 0; k < n_poses; ++k) if (!active[k]) memcpy(cand + 7 * k, poses + 7 * k, 7 * sizeof(double));
    t0 = now_s();
    oracle_evaluate(n_poses, cand, pose_const, n_edges, edge_ids, edge_meas, edge_sqrt_info,
                    opt->loss_type, opt->loss_a, &cand_cost, NULL, NULL, NULL);
    summary->time_residual_s += now_s() - t0; summary->num_residual_evals++;
    if (!isfinite(cand_cost)) cand_cost = DBL_MAX;

    /* ---- ParameterToleranceReached ---- */
    step_norm = 0.0;
    for (k = 0; k < n_poses; ++k) if (active[k]) for (e = 0; e < 7; ++e) { const double dd = poses[7 * k + e] - cand[7 * k + e]; step_norm += dd * dd; }
    step_norm = sqrt(step_norm);
    it.step_norm = step_norm;
    if (step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
      summary->termination_type = ORACLE_CONVERGENCE;
      snprintf(summary->message, sizeof summary->message, "Parameter tolerance reached. Relative step_norm: %e <= %e.", step_norm / (x_norm + opt->parameter_tolerance), opt->parameter_tolerance);
      it.cost = x_cost; log_iter(iter_log, iter_log_cap, summary, &it);
      break;
    }
    /* ---- FunctionToleranceReached ---- */
    cost_change = x_cost - cand_cost;
    it.cost_change = cost_change;
    if (fabs(cost_change) <= opt->function_tolerance * x_cost) {
      summary->termination_type = ORACLE_CONVERGENCE;
      snprintf(summary->message, sizeof summary->message, "Function tolerance reached. |cost_change|/cost: %e <= %e", fabs(cost_change) / x_cost, opt->function_tolerance);
      it.cost = x_cost; log_iter(iter_log, iter_log_cap, summary, &it);
      break;
    }
    /* ---- IsStepSuccessful ---- */
    relative_decrease = cost_change / model_cost_change;
    it.relative_decrease = relative_decrease;
    if (relative_decrease > opt->min_relative_decrease) {
      /* HandleSuccessfulStep */
      double t;
      memcpy(poses, cand, sizeof(double) * 7 * (size_t)n_poses);
      x_norm = 0.0;
      for (k = 0; k < n_poses; ++k) if (active[k]) for (e = 0; e < 7; ++e) if (BLOCK_VARIABLE(pose_const, k, e)) x_norm += poses[7 * k + e] * poses[7 * k + e];
      x_norm = sqrt(x_norm);
      t0 = now_s();
      oracle_evaluate(n_poses, poses, pose_const, n_edges, edge_ids, edge_meas, edge_sqrt_info,
                      opt->loss_type, opt->loss_a, &x_cost, residuals, gradient, jac);
      summary->time_jacobian_s += now_s() - t0; summary->num_jacobian_evals++;
      if (opt->jacobi_scaling) scale_columns(n_edges, edge_ids, jac, scale);
      projected_gradient_norms(n_poses, poses, active, gradient, &it.gradient_max_norm, &it.gradient_norm);
      it.step_is_successful = 1; it.cost = x_cost;
      summary->num_successful_steps++;
      /* LevenbergMarquardtStrategy::StepAccepted */
      t = 2.0 * relative_decrease - 1.0;
      t = 1.0 - t * t * t; if (t < 1.0 / 3.0) t = 1.0 / 3.0;
      radius = radius / t; if (radius > opt->max_trust_region_radius) radius = opt->max_trust_region_radius;
      decrease_factor = 2.0; reuse_diagonal = 0;
    } else {
      /* HandleUnsuccessfulStep -> StepRejected */
      it.step_is_successful = 0; it.cost = x_cost;
      summary->num_unsuccessful_steps++;
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
    }
    log_iter(iter_log, iter_log_cap, summary, &it);
  }
  summary->final_cost = x_cost;
done:
  summary->time_total_s = now_s() - t_begin;
  chol_free(C);
  free(residuals); free(jac); free(gradient); free(scale); free(diagonal); free(lm_diag_sq);
  free(rhs); free(step); free(delta); free(cand); free(active);
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Loop-edge candidate lists: generate_edges_from_trajectory_origion.cpp:58-110.
 * For frame c: first c-1 (:61), then, if c > gap, every i in [0, c-gap) (:65, the inner "c - i > gap" test :70 is
 * implied) whose centre is in range.  isInSearchRange (:84-110) works on float (CV_32F) coordinates:
 * dist = dA*dA + dB*dB + dC*dC, each operation rounded to float (x86-64 build of the reference: no FMA), and the
 * frame is rejected when dist > radius*radius.  volatile keeps this compiler from contracting into FMAs.
 * ------------------------------------------------------------------------------------------------ */
static int oracle_in_search_range(const float* c, const float* p, float r2) {
  volatile float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
  volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  volatile float s = xx + yy;
  volatile float dist = s + zz;
  return !(dist > r2);
}

long long oracle_edge_candidates(int n_frames, const double* positions, double search_radius, int min_frame_gap,
                                 long long* row_ptr, int* candidates, long long capacity) {
  float* pf = (float*)malloc(sizeof(float) * 3 * (size_t)(n_frames > 0 ? n_frames : 1));
  const float r = (float)search_radius;
  const float r2 = r * r;
  long long total = 0;
  for (int i = 0; i < 3 * n_frames; ++i) pf[i] = (float)positions[i];
  row_ptr[0] = 0;
  if (n_frames > 0) row_ptr[1] = 0;               /* frame 0 has no line in the file (:38 starts at id = 1) */
  for (int c = 1; c < n_frames; ++c) {
    if (candidates && total < capacity) candidates[total] = c - 1;
    ++total;
    for (int i = 0; i < c - min_frame_gap; ++i)
      if (oracle_in_search_range(pf + 3 * c, pf + 3 * i, r2)) {
        if (candidates && total < capacity) candidates[total] = i;
        ++total;
      }
    row_ptr[c + 1] = total;
  }
  free(pf);
  return total;
}
