// pgo_edge_math.cuh -- closed-form SE(3) relative-pose residual and its two 6x6 local Jacobians.
//
// Restates what ceres::AutoDiffCostFunction<PoseGraph3dErrorTerm,6,3,4,3,4> followed by
// EigenQuaternionParameterization::ComputeJacobian produces for one edge
// (REF/include/PoseGraph3dError.h:21-54), but analytically instead of with jets:
//
//   r_p = conj(q_a) * (p_b - p_a) - p_m          (Eigen _transformVector, q NOT normalised)
//   r_q = 2 vec( q_m * conj(conj(q_a) * q_b) ) = 2 vec( q_m * conj(q_b) * q_a )
//   r   = S [r_p; r_q]
//
// With u = -vec(q_a), w = q_a.w, d = p_b - p_a, c = q_m * conj(q_b) and Pa the 4x3 plus-Jacobian
// of q_a (columns f_k = (e_k, 0) * q_a):
//   M    = d r_p / d p_b = (1 - 2|u|^2) I + 2 u u^T + 2 w [u]x           d r_p / d p_a = -M
//   T    = d r_p / d delta_a = [-Gu | 2 u x d] Pa,   Gu = -2 w [d]x + 2((u.d) I + u d^T - 2 d u^T)
//   Arot = d r_q / d delta_a,  column k = 2 (c.w f_k.v + f_k.w c.v + c.v x f_k.v)   d r_q / d delta_b = -Arot
//   Ja = S [[-M, T], [0, Arot]]        Jb = S [[M, 0], [0, -Arot]]
// so both Jacobians are carried by three 6x3 panels B1 = S[:,0:3] M, B2 = S[:,3:6] Arot,
// C = S[:,0:3] T + B2:   Ja = [-B1 | C],  Jb = [B1 | -B2].
#pragma once

#include "pgo_common.cuh"

namespace pgo {

struct EdgePanels {
  double r[6];
  double B1[6][3];
  double B2[6][3];
  double C[6][3];
};

// S(i,k) accessor: sqrt_information row i, column k.
template <bool kIdentityInfo, typename SFn>
__device__ __forceinline__ void edge_residual_only(const double* pa, const double* pb, const double* m, SFn S,
                                                   double* r) {
  const double ux = -pa[3], uy = -pa[4], uz = -pa[5], w = pa[6];
  const double dx = pb[0] - pa[0], dy = pb[1] - pa[1], dz = pb[2] - pa[2];
  double cx = uy * dz - uz * dy, cy = uz * dx - ux * dz, cz = ux * dy - uy * dx;
  cx += cx; cy += cy; cz += cz;
  double t[6];
  t[0] = dx + w * cx + (uy * cz - uz * cy) - m[0];
  t[1] = dy + w * cy + (uz * cx - ux * cz) - m[1];
  t[2] = dz + w * cz + (ux * cy - uy * cx) - m[2];
  // c = q_m * conj(q_b)
  const double bx = pb[3], by = pb[4], bz = pb[5], bw = pb[6];
  const double mx = m[3], my = m[4], mz = m[5], mw = m[6];
  const double qw = mw * bw + mx * bx + my * by + mz * bz;
  const double qx = -mw * bx + bw * mx - (my * bz - mz * by);
  const double qy = -mw * by + bw * my - (mz * bx - mx * bz);
  const double qz = -mw * bz + bw * mz - (mx * by - my * bx);
  // dq = c * q_a
  const double ax = pa[3], ay = pa[4], az = pa[5], aw = pa[6];
  t[3] = 2.0 * (qw * ax + aw * qx + (qy * az - qz * ay));
  t[4] = 2.0 * (qw * ay + aw * qy + (qz * ax - qx * az));
  t[5] = 2.0 * (qw * az + aw * qz + (qx * ay - qy * ax));
  if (kIdentityInfo) {
#pragma unroll
    for (int i = 0; i < 6; ++i) r[i] = t[i];
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s = fma(S(i, k), t[k], s);
      r[i] = s;
    }
  }
}

template <bool kIdentityInfo, typename SFn>
__device__ __forceinline__ void edge_linearize(const double* pa, const double* pb, const double* m, SFn S,
                                               EdgePanels& o) {
  const double ax = pa[3], ay = pa[4], az = pa[5], aw = pa[6];
  const double ux = -ax, uy = -ay, uz = -az, w = aw;
  const double dx = pb[0] - pa[0], dy = pb[1] - pa[1], dz = pb[2] - pa[2];
  // uv = 2 u x d
  const double vx = 2.0 * (uy * dz - uz * dy), vy = 2.0 * (uz * dx - ux * dz), vz = 2.0 * (ux * dy - uy * dx);
  double t[6];
  t[0] = dx + w * vx + (uy * vz - uz * vy) - m[0];
  t[1] = dy + w * vy + (uz * vx - ux * vz) - m[1];
  t[2] = dz + w * vz + (ux * vy - uy * vx) - m[2];

  // M = (1 - 2|u|^2) I + 2 u u^T + 2 w [u]x
  double M[3][3];
  {
    const double uu = ux * ux + uy * uy + uz * uz;
    const double k = 1.0 - 2.0 * uu;
    const double w2 = 2.0 * w;
    M[0][0] = k + 2.0 * ux * ux; M[0][1] = 2.0 * ux * uy - w2 * uz; M[0][2] = 2.0 * ux * uz + w2 * uy;
    M[1][0] = 2.0 * uy * ux + w2 * uz; M[1][1] = k + 2.0 * uy * uy; M[1][2] = 2.0 * uy * uz - w2 * ux;
    M[2][0] = 2.0 * uz * ux - w2 * uy; M[2][1] = 2.0 * uz * uy + w2 * ux; M[2][2] = k + 2.0 * uz * uz;
  }
  // nGu = -Gu = 2 w [d]x - 2((u.d) I + u d^T - 2 d u^T)
  double nG[3][3];
  {
    const double ud = ux * dx + uy * dy + uz * dz;
    const double w2 = 2.0 * w;
    nG[0][0] = -2.0 * (ud + ux * dx - 2.0 * dx * ux);
    nG[0][1] = -w2 * dz - 2.0 * (ux * dy - 2.0 * dx * uy);
    nG[0][2] = w2 * dy - 2.0 * (ux * dz - 2.0 * dx * uz);
    nG[1][0] = w2 * dz - 2.0 * (uy * dx - 2.0 * dy * ux);
    nG[1][1] = -2.0 * (ud + uy * dy - 2.0 * dy * uy);
    nG[1][2] = -w2 * dx - 2.0 * (uy * dz - 2.0 * dy * uz);
    nG[2][0] = -w2 * dy - 2.0 * (uz * dx - 2.0 * dz * ux);
    nG[2][1] = w2 * dx - 2.0 * (uz * dy - 2.0 * dz * uy);
    nG[2][2] = -2.0 * (ud + uz * dz - 2.0 * dz * uz);
  }
  // Pa columns f_k = (e_k,0) * q_a : rows x,y,z,w
  const double P[4][3] = {{aw, az, -ay}, {-az, aw, ax}, {ay, -ax, aw}, {-ax, -ay, -az}};
  double T[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T[0][k] = nG[0][0] * P[0][k] + nG[0][1] * P[1][k] + nG[0][2] * P[2][k] + vx * P[3][k];
    T[1][k] = nG[1][0] * P[0][k] + nG[1][1] * P[1][k] + nG[1][2] * P[2][k] + vy * P[3][k];
    T[2][k] = nG[2][0] * P[0][k] + nG[2][1] * P[1][k] + nG[2][2] * P[2][k] + vz * P[3][k];
  }
  // c = q_m * conj(q_b)
  const double bx = pb[3], by = pb[4], bz = pb[5], bw = pb[6];
  const double mx = m[3], my = m[4], mz = m[5], mw = m[6];
  const double qw = mw * bw + mx * bx + my * by + mz * bz;
  const double qx = -mw * bx + bw * mx - (my * bz - mz * by);
  const double qy = -mw * by + bw * my - (mz * bx - mx * bz);
  const double qz = -mw * bz + bw * mz - (mx * by - my * bx);
  t[3] = 2.0 * (qw * ax + aw * qx + (qy * az - qz * ay));
  t[4] = 2.0 * (qw * ay + aw * qy + (qz * ax - qx * az));
  t[5] = 2.0 * (qw * az + aw * qz + (qx * ay - qy * ax));
  double A[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double fx = P[0][k], fy = P[1][k], fz = P[2][k], fw = P[3][k];
    A[0][k] = 2.0 * (qw * fx + fw * qx + (qy * fz - qz * fy));
    A[1][k] = 2.0 * (qw * fy + fw * qy + (qz * fx - qx * fz));
    A[2][k] = 2.0 * (qw * fz + fw * qz + (qx * fy - qy * fx));
  }
  if (kIdentityInfo) {
#pragma unroll
    for (int i = 0; i < 6; ++i) o.r[i] = t[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        o.B1[i][k] = M[i][k]; o.B1[i + 3][k] = 0.0;
        o.B2[i][k] = 0.0;     o.B2[i + 3][k] = A[i][k];
        o.C[i][k] = T[i][k];  o.C[i + 3][k] = A[i][k];
      }
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double s0 = S(i, 0), s1 = S(i, 1), s2 = S(i, 2), s3 = S(i, 3), s4 = S(i, 4), s5 = S(i, 5);
      o.r[i] = s0 * t[0] + s1 * t[1] + s2 * t[2] + s3 * t[3] + s4 * t[4] + s5 * t[5];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double b2 = s3 * A[0][k] + s4 * A[1][k] + s5 * A[2][k];
        o.B1[i][k] = s0 * M[0][k] + s1 * M[1][k] + s2 * M[2][k];
        o.B2[i][k] = b2;
        o.C[i][k] = s0 * T[0][k] + s1 * T[1][k] + s2 * T[2][k] + b2;
      }
    }
  }
}

}  // namespace pgo
