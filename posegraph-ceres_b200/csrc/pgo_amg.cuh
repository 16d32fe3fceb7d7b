// pgo_amg.cuh -- PCG for the damped normal equations preconditioned by an aggregation multigrid cycle whose coarse
// spaces are the rigid-body modes of pose patches (hierarchy: pgo_amg_host.hpp).  Solver type PGO_LINEAR_PCG_AMG: the
// solver of mesh-like graphs (sphere, grids, dense random loops) and of every multi-GPU solve.  Included by pgo_b200.cu.
//
// Per LM step (amg_setup_numeric): patch centroids from the current poses, the Galerkin operators A_{l+1} = P^T A_l P
// gathered block by block in a fixed order (no atomics: every rank computes bit-identical replicated levels), 6x6
// block-Jacobi inverses, fp32 copies of the operators for the sweeps inside the cycle, and the dense inverse of the last
// level (<= 512 nodes: block Gauss-Jordan by a cooperative grid).
// Per PCG iteration (Chronopoulos-Gear form, ONE reduction per iteration): cycle u = M^-1 r (V, or W on the first coarse
// levels of large graphs: amg_vcycle), halo exchange of u, w = A u with the partial sum of w.u in its epilogue, r.u,
// all-reduce of the two scalars, vector update fused with the first pre-smoothing sweep of the next cycle.  The whole
// iteration is one CUDA graph; on several GPUs its exchanges run over NVLink peer memory (pgo_peer.cuh).
//
// Prolongator of fine node i in aggregate I (d = p_i - c_I, c_I the patch centroid, S the Jacobi column scaling):
//   P_i = S_i^-1 [[I, X], [0, I]],  X = -2 [d]x      (a patch rotation delta moves p_i by 2 delta x d: the local rotation
//   coordinates of EigenQuaternionParameterization are HALF rotation vectors applied on the left, i.e. in the world frame)
// so P is never stored: restriction and prolongation cost one cross product per node.
#pragma once

#include <functional>

#include "pgo_amg_host.hpp"

namespace pgo {

constexpr int kAmgThreads = 256;
constexpr int kAmgDenseSmemNodes = 16;     // coarsest level of at most this many nodes: Gauss-Jordan in the shared memory of one CTA
constexpr int kAmgDenseMaxNodes = 512;     // ... up to this many: block Gauss-Jordan in global memory by a cooperative grid (3072^2 fp64 = 75 MB)
constexpr int kGjThreads = 512;
constexpr int kGjMaxOwnRows = 6 * 8;       // block rows of the dense system a CTA may own (x 6 scalar rows)

template <typename T> struct BsrViewT { int n; const T* Hdiag; const T* Hoff; const int* row_ptr; const int* col_idx; };
struct AmgLevelDev {
  bool replicated = false;
  int n_own = 0, n_halo = 0;
  // operator of the stored rows (panel layout, like the graph's Hessian); level 0 aliases the graph's arrays
  double* Adiag = nullptr;
  double* Aoff = nullptr;
  int* row_ptr = nullptr;
  int* col_idx = nullptr;
  long long nnz = 0;
  float *Adiag_f = nullptr, *Aoff_f = nullptr;   // fp32 copy of the operator: what the sweeps INSIDE the preconditioner read
  double* Dinv = nullptr;                  // [n_own][36] row-major inverses of the diagonal blocks (level 0: Minv)
  double *r = nullptr, *x = nullptr, *y = nullptr;   // [n_own + n_halo][6]
  double* z = nullptr;                     // third iterate buffer of the levels a W-cycle visits twice
  double* pos = nullptr;                   // [n_own + n_halo][pos_stride] (level 0: the pose array)
  int pos_stride = 3;
  // coarsening towards the next level
  int* agg = nullptr;
  int c_row0 = 0, c_row1 = 0, c_slot0 = 0, n_cblk = 0;
  int *mem_ptr = nullptr, *mem_idx = nullptr, *gal_ptr = nullptr, *gal_row = nullptr, *gal_slot = nullptr;
  // exchange plans (host copies drive the NCCL calls)
  std::vector<int> nbr, send_ptr, recv_ptr, gather_off, gather_slot_off;
  int* send_idx = nullptr;
  // the same exchanges over peer memory (pgo_peer.cuh): halo of this level, residual gather INTO this level
  int peer_ch = -1, gather_ch = -1;
  PeerPushArgs push, gpush;
  PeerWaitArgs wait, gwait;
  int push_ctas = 1, wait_ctas = 1, gpush_ctas = 1, gwait_ctas = 1;
};

struct Amg {
  int num_levels = 0;
  std::vector<AmgLevelDev> lv;
  double* dense_inv = nullptr;             // [6 n][6 n] of the coarsest level (n <= kAmgDenseMaxNodes), else nullptr
  double* gj_rows = nullptr;               // [2][6][6 n] pivot rows of the block Gauss-Jordan (n > kAmgDenseSmemNodes)
  unsigned int* gj_counter = nullptr;
  double omega = 0.85;                     // damped block-Jacobi smoother (measured on the 1M grid and the 100k torus: 0.6 < 0.7 < 0.85)
  int nu = 1;
  int coarse_sweeps = 4;
  int gamma = 1, gamma_depth = 0;          // cycle shape, see amg_vcycle
  bool fp32_ops = true;                    // smoothing / residual sweeps of the cycle read the fp32 operator copies
  float* Minv_f = nullptr;                 // fp32 copy of level 0's block inverses (large level 0 only), else nullptr
  pgo_graph* owner = nullptr;
  PcgMultiState* state = nullptr;
  PcgMultiState* state_h = nullptr;        // pinned, two slots
  cudaEvent_t ev[2] = {nullptr, nullptr};
  double* part = nullptr;                  // per-CTA partial sums
  int part_cap = 0;
  bool warp_spmv = false;
  double* red = nullptr;                   // [8] reduced scalars (all-reduced across ranks)
  // multi-GPU: the replicated tail of the V-cycle as a CUDA graph
  cudaGraphExec_t tail_graph = nullptr;
  int tail_graph_kernels = 0, tail_calls = 0;
  bool tail_graph_failed = false;
  std::vector<double*> tail_cur;
  struct IterGraph { const void* key[4]; int max_it; double tol; cudaGraphExec_t exec; int kernels; long long pushes, push_bytes; };
  std::vector<IterGraph> graphs;           // captured PCG iteration per (H, poses) buffer pair (the LM loop alternates two)
  // multi-GPU: per-iteration exchanges over peer memory instead of NCCL (nullptr: NCCL)
  PeerCtx* peer = nullptr;
  PeerPushArgs rpush;                      // the 2-scalar all-reduce
  PeerWaitArgs rwait;
  bool whole_iteration_graph = false;      // the PCG iteration is being captured / replayed as ONE graph
  unsigned long long* prof = nullptr;      // [64] PGO_AMG_PROFILE stage times (ns), [63] = the previous mark
  int peer_exchanges_per_iteration = 0;
  long long blocks_all_levels = 0;
  long long comm_bytes_per_iteration = 0;  // payload this rank sends per PCG iteration (halo + gathers + scalars)
  int comm_calls_per_iteration = 0;
};

// --------------------------------------------------------------------------------------------
// kernels
// --------------------------------------------------------------------------------------------
// B(k, r) of the mode matrix [[I, X], [0, I]], X = -2 [d]x
__device__ __forceinline__ double amg_mode(const double* d, int k, int r) {
  if (k == r) return 1.0;
  if (k < 3 && r >= 3) {
    const int c = r - 3;
    // X = [[0, 2dz, -2dy], [-2dz, 0, 2dx], [2dy, -2dx, 0]]
    if (k == 0) return c == 1 ? 2.0 * d[2] : (c == 2 ? -2.0 * d[1] : 0.0);
    if (k == 1) return c == 0 ? -2.0 * d[2] : (c == 2 ? 2.0 * d[0] : 0.0);
    return c == 0 ? 2.0 * d[1] : (c == 1 ? -2.0 * d[0] : 0.0);
  }
  return 0.0;
}

__device__ __forceinline__ void amg_delta(const double* pos, int stride, int i, const double* cpos, int I, double* d) {
  d[0] = pos[(size_t)stride * i] - cpos[3 * (size_t)I];
  d[1] = pos[(size_t)stride * i + 1] - cpos[3 * (size_t)I + 1];
  d[2] = pos[(size_t)stride * i + 2] - cpos[3 * (size_t)I + 2];
}

// centroids of the computed coarse rows
__global__ void amg_centroid_kernel(int ncomp, int c_row0, const int* __restrict__ mem_ptr, const int* __restrict__ mem_idx,
                                    const double* __restrict__ pos, int stride, double* __restrict__ cpos) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncomp) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  const int m0 = mem_ptr[k], m1 = mem_ptr[k + 1];
  for (int m = m0; m < m1; ++m) {
    const int i = mem_idx[m];
    s0 += pos[(size_t)stride * i]; s1 += pos[(size_t)stride * i + 1]; s2 += pos[(size_t)stride * i + 2];
  }
  const double inv = 1.0 / (double)max(m1 - m0, 1);
  double* o = cpos + 3 * (size_t)(c_row0 + k);
  o[0] = s0 * inv; o[1] = s1 * inv; o[2] = s2 * inv;
}

struct GalerkinParams {
  int n_cblk, ncomp, c_row0, c_slot0;
  const int *gal_ptr, *gal_row, *gal_slot;
  const double *Adiag, *Aoff;
  const int* col_idx;
  const double* dlm;        // level 0: LM diagonal added to the diagonal blocks; else nullptr
  const double* scale;      // level 0: Jacobi column scaling [n_loc][6]; else nullptr (identity)
  const int* agg;
  const double* pos; int pos_stride;
  const double* cpos;
  double *Cdiag, *Coff;
};

// Six lanes per coarse block (lane r owns row r of it), five blocks per warp -- the lane layout of bsr6_row, so every lane
// loads only ITS row of a fine block (three 128-bit loads, 96 contiguous bytes per load across the group):
//   Z = sum over the gather list of P_i^T A P_j,   P = S^-1 [[I, X], [0, I]],
//   y_r = row r of P_i^T A = A(r,:)/s_i[r] + [r >= 3] sum_{k<3} X_i(k, r-3)/s_i[k] A(k,:)     (rows 0..2 come by shuffle)
//   Z(r,c) = y_r(c)/s_j[c] + [c >= 3] sum_{m<3} y_r(m) X_j(m, c-3)/s_j[m]
// summed in list order (deterministic).  (Round 2 first had one thread per ELEMENT of the coarse block: 36 threads redid
// the index chasing and the mode algebra of every contribution -- 19 ms on the 1M-pose level, instruction bound.)
__global__ void __launch_bounds__(kAmgThreads) amg_galerkin_kernel(const GalerkinParams P) {
  const int lane = threadIdx.x & 31, grp = lane / 6, r = lane - grp * 6, g0 = grp * 6;
  const int cb = (blockIdx.x * (kAmgThreads / 32) + (threadIdx.x >> 5)) * kRowsPerWarp + grp;
  const bool valid = grp < kRowsPerWarp && cb < P.n_cblk;
  const int q0 = valid ? __ldg(P.gal_ptr + cb) : 0, q1 = valid ? __ldg(P.gal_ptr + cb + 1) : 0;
  const unsigned full = 0xffffffffu;
  int maxlen = q1 - q0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(full, maxlen, o));
  double z[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int t = 0; t < maxlen; ++t) {
    const bool on = q0 + t < q1;
    double a[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double si[3] = {1.0, 1.0, 1.0}, sir = 1.0, sj[6] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
    double di[3] = {0.0, 0.0, 0.0}, dj[3] = {0.0, 0.0, 0.0};
    if (on) {
      const int i = __ldg(P.gal_row + q0 + t), p = __ldg(P.gal_slot + q0 + t);
      const int j = p < 0 ? i : __ldg(P.col_idx + p);
      const double2* blk = reinterpret_cast<const double2*>(p < 0 ? P.Adiag + 36 * (size_t)i : P.Aoff + 36 * (size_t)p) + r;
      const double2 a0 = __ldg(blk), a1 = __ldg(blk + 6), a2 = __ldg(blk + 12);
      a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y;
      if (p < 0 && P.dlm) {
        const double dl = __ldg(P.dlm + 6 * (size_t)i + r);
#pragma unroll
        for (int c = 0; c < 6; ++c) if (c == r) a[c] += dl;
      }
      amg_delta(P.pos, P.pos_stride, i, P.cpos, __ldg(P.agg + i), di);
      amg_delta(P.pos, P.pos_stride, j, P.cpos, __ldg(P.agg + j), dj);
      if (P.scale) {
        const double* sc = P.scale + 6 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double v = __ldg(sc + k); si[k] = v > 0.0 ? 1.0 / v : 0.0; }
        { const double v = __ldg(sc + r); sir = v > 0.0 ? 1.0 / v : 0.0; }
        const double* sd = P.scale + 6 * (size_t)j;
#pragma unroll
        for (int m = 0; m < 6; ++m) { const double v = __ldg(sd + m); sj[m] = v > 0.0 ? 1.0 / v : 0.0; }
      }
    }
    // rows 0..2 of A from the lanes that hold them
    double top[3][6];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 6; ++c) top[k][c] = __shfl_sync(full, a[c], g0 + k);
    if (!on) continue;
    // X(k, c') of a node with offset d: [[0, 2dz, -2dy], [-2dz, 0, 2dx], [2dy, -2dx, 0]]
    double y[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) y[c] = a[c] * sir;
    if (r >= 3) {
      const int cc = r - 3;
      const double x0 = cc == 0 ? 0.0 : (cc == 1 ? 2.0 * di[2] : -2.0 * di[1]);     // X_i(0, cc)
      const double x1 = cc == 0 ? -2.0 * di[2] : (cc == 1 ? 0.0 : 2.0 * di[0]);     // X_i(1, cc)
      const double x2 = cc == 0 ? 2.0 * di[1] : (cc == 1 ? -2.0 * di[0] : 0.0);     // X_i(2, cc)
      const double w0 = x0 * si[0], w1 = x1 * si[1], w2 = x2 * si[2];
#pragma unroll
      for (int c = 0; c < 6; ++c) y[c] = fma(w0, top[0][c], fma(w1, top[1][c], fma(w2, top[2][c], y[c])));
    }
    const double u0 = y[0] * sj[0], u1 = y[1] * sj[1], u2 = y[2] * sj[2];
    z[0] += u0; z[1] += u1; z[2] += u2;
    // columns 3..5: y(c) / s_j[c] + sum_{m<3} u_m X_j(m, c - 3)
    z[3] += fma(u1, -2.0 * dj[2], fma(u2, 2.0 * dj[1], y[3] * sj[3]));
    z[4] += fma(u0, 2.0 * dj[2], fma(u2, -2.0 * dj[0], y[4] * sj[4]));
    z[5] += fma(u0, -2.0 * dj[1], fma(u1, 2.0 * dj[0], y[5] * sj[5]));
  }
  if (!valid) return;
  double* out = cb < P.ncomp ? P.Cdiag + 36 * (size_t)(P.c_row0 + cb) : P.Coff + 36 * (size_t)(P.c_slot0 + cb - P.ncomp);
  double2* o2 = reinterpret_cast<double2*>(out) + r;
  o2[0] = make_double2(z[0], z[1]); o2[6] = make_double2(z[2], z[3]); o2[12] = make_double2(z[4], z[5]);
}

// inverse of a symmetric positive definite 6x6 block by Cholesky; false when a pivot is not positive
__device__ __forceinline__ bool spd6_inverse(double (&A)[6][6], double* out /* row-major 36 */) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[j][j];
#pragma unroll
    for (int k = 0; k < 6; ++k) if (k < j) s -= A[j][k] * A[j][k];
    if (!(s > 0.0)) { ok = false; s = 1.0; }
    const double l = sqrt(s);
    A[j][j] = l;
#pragma unroll
    for (int r = 0; r < 6; ++r) if (r > j) {
      double t = A[r][j];
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k < j) t -= A[r][k] * A[j][k];
      A[r][j] = t / l;
    }
  }
  double Li[6][6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r < c) { Li[r][c] = 0.0; continue; }
      double t = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k >= c && k < r) t -= A[r][k] * Li[k][c];
      Li[r][c] = t / A[r][r];
    }
  }
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k >= r && k >= c) t += Li[k][r] * Li[k][c];
      out[r * 6 + c] = t;
    }
  return ok;
}

// block-Jacobi inverses of a coarse level
__global__ void amg_block_inverse_kernel(int n, const double* __restrict__ Adiag, double* __restrict__ Dinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double A[6][6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) A[r][c] = Adiag[36 * (size_t)i + pidx(r, c)];
  double inv[36];
  const bool ok = spd6_inverse(A, inv);
#pragma unroll
  for (int k = 0; k < 36; ++k) Dinv[36 * (size_t)i + k] = ok ? inv[k] : 0.0;
}

// Dense inverse of the coarsest operator (n <= kAmgDenseSmemNodes nodes) by Gauss-Jordan in shared memory, one CTA.
__global__ void __launch_bounds__(kAmgThreads) amg_dense_inverse_kernel(int n, const double* __restrict__ Adiag, const double* __restrict__ Aoff,
                                                                        const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                                                        double* __restrict__ inv) {
  extern __shared__ double M[];
  const int m = 6 * n;
  for (int k = threadIdx.x; k < m * m; k += blockDim.x) M[k] = 0.0;
  __syncthreads();
  for (int t = threadIdx.x; t < n * 36; t += blockDim.x) {
    const int i = t / 36, e = t - i * 36, r = e / 6, c = e - r * 6;
    M[(6 * i + r) * m + 6 * i + c] = Adiag[36 * (size_t)i + pidx(r, c)];
  }
  const int nnz = row_ptr[n];
  for (int t = threadIdx.x; t < nnz * 36; t += blockDim.x) {
    const int p = t / 36, e = t - p * 36, r = e / 6, c = e - r * 6;
    // row of slot p
    int i = 0;
    while (row_ptr[i + 1] <= p) ++i;
    M[(6 * i + r) * m + 6 * col_idx[p] + c] = Aoff[36 * (size_t)p + pidx(r, c)];
  }
  __syncthreads();
  __shared__ double piv_s;
  for (int k = 0; k < m; ++k) {
    if (threadIdx.x == 0) { const double d = M[k * m + k]; piv_s = d != 0.0 ? 1.0 / d : 0.0; }
    __syncthreads();
    const double piv = piv_s;
    // scale the pivot row (except the pivot)
    for (int j = threadIdx.x; j < m; j += blockDim.x) if (j != k) M[k * m + j] *= piv;
    __syncthreads();
    for (int t = threadIdx.x; t < m * m; t += blockDim.x) {
      const int i = t / m, j = t - i * m;
      if (i != k && j != k) M[t] -= M[i * m + k] * M[k * m + j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      if (i != k) M[i * m + k] = -M[i * m + k] * piv;
      else M[k * m + k] = piv;
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < m * m; k += blockDim.x) inv[k] = M[k];
}

// ---- coarsest levels of up to kAmgDenseMaxNodes nodes: explicit inverse by 6x6-block Gauss-Jordan in global memory ----
// (an aggregation V-cycle loses a constant factor per level, and the small levels are pure launch latency: solving the
// first level of <= 512 nodes exactly removes two or three levels from every cycle -- sphere2500 57 -> 26 PCG iterations
// per LM step in tools/amg_prototype.py --dense-below -- for one inversion per LM step and one dense mat-vec per cycle.)
// scatter of the block-CSR operator into the zeroed dense matrix, one CTA per block row
__global__ void __launch_bounds__(kAmgThreads) amg_dense_fill_kernel(int n, const double* __restrict__ Adiag, const double* __restrict__ Aoff,
                                                                     const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                                                     double* __restrict__ M) {
  const int i = blockIdx.x;
  const size_t m = 6 * (size_t)n;
  const int p0 = row_ptr[i], p1 = row_ptr[i + 1];
  for (int t = threadIdx.x; t < (p1 - p0 + 1) * 36; t += blockDim.x) {
    const int q = t / 36, e = t - q * 36, r = e / 6, c = e - r * 6;
    if (q == 0) M[(6 * (size_t)i + r) * m + 6 * (size_t)i + c] = Adiag[36 * (size_t)i + pidx(r, c)];
    else {
      const int p = p0 + q - 1;
      M[(6 * (size_t)i + r) * m + 6 * (size_t)col_idx[p] + c] = Aoff[36 * (size_t)p + pidx(r, c)];
    }
  }
}

struct GjParams {
  int nb;                  // block rows (nodes); the matrix is [6 nb][6 nb], row-major
  double* M;               // in: the SPD operator, out: its inverse
  double* R;               // [2][6][6 nb]: the normalised pivot rows of the current / the next step
  unsigned int* counter;   // grid barrier (zeroed before the launch)
};

// In-place block Gauss-Jordan without pivoting (the operator is SPD).  Every CTA owns a contiguous range of block rows.
// Step k, with R = P_k M[K, :] (P_k the inverse of the pivot block, and R[:, K] := P_k):
//   rows of K:  M[K, :] = R;      other rows i:  M[i, j] = (j in K ? 0 : M[i, j]) - sum_q M_old[i, K_q] R[q, j].
// The owner of block row k + 1 updates those six rows first, inverts the next pivot and publishes the next R (double
// buffered) before it turns to its other rows: ONE grid barrier per step.
__global__ void __launch_bounds__(kGjThreads) amg_dense_gj_kernel(const GjParams P) {
  __shared__ double Cs[kGjMaxOwnRows][6];
  __shared__ double Pk[36];
  const int nb = P.nb;
  const size_t m = 6 * (size_t)nb;
  const int per = (nb + gridDim.x - 1) / gridDim.x;
  const int b0 = min(nb, (int)blockIdx.x * per), b1 = min(nb, b0 + per);
  const int nrows = 6 * (b1 - b0);
  double* const M = P.M;
  unsigned int epoch = 0;

  // R_kn = P M[Kn, :] from the (final) rows of block kn, which this CTA owns
  auto publish = [&](int kn) {
    __syncthreads();                                   // the rows of Kn were written by this CTA's threads
    if (threadIdx.x == 0) {
      double A[6][6];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) A[r][c] = M[(6 * (size_t)kn + r) * m + 6 * (size_t)kn + c];
      double inv[36];
      const bool ok = spd6_inverse(A, inv);
#pragma unroll
      for (int e = 0; e < 36; ++e) Pk[e] = ok ? inv[e] : 0.0;
    }
    __syncthreads();
    double* Rn = P.R + (size_t)(kn & 1) * 6 * m;
    const double* rows = M + 6 * (size_t)kn * m;
    for (int j = threadIdx.x; j < (int)m; j += blockDim.x) {
      const int c = j - 6 * kn;
      if (c >= 0 && c < 6) {
#pragma unroll
        for (int q = 0; q < 6; ++q) Rn[q * m + j] = Pk[q * 6 + c];
      } else {
        double a[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) a[r] = rows[r * m + j];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          double t = 0.0;
#pragma unroll
          for (int r = 0; r < 6; ++r) t = fma(Pk[q * 6 + r], a[r], t);
          Rn[q * m + j] = t;
        }
      }
    }
  };

  if (b0 <= 0 && 0 < b1) publish(0);
  grid_barrier(P.counter, epoch);
  for (int k = 0; k < nb; ++k) {
    const double* R = P.R + (size_t)(k & 1) * 6 * m;
    for (int t = threadIdx.x; t < nrows * 6; t += blockDim.x) Cs[t / 6][t % 6] = M[(6 * (size_t)b0 + t / 6) * m + 6 * (size_t)k + t % 6];
    __syncthreads();
    // rows come in blocks of six: all twelve loads of a (block row, column) item are issued before its 36 FMAs (the
    // matrix lives in L2: a row-at-a-time loop exposed one ~0.7 us round trip per row -- 29 us per step at 2 400 unknowns)
    auto update = [&](int r_lo, int r_hi) {           // local scalar rows [r_lo, r_hi), a multiple of six
      for (int rb = r_lo; rb < r_hi; rb += 6) {
        const int gi0 = 6 * b0 + rb;
        const bool pivot_rows = gi0 == 6 * k;
        for (int j0 = threadIdx.x; j0 < (int)m; j0 += 2 * blockDim.x) {
          // two columns per trip: 24 loads in flight per thread
          double Rq[2][6], a[2][6];
          int jj[2];
          bool on[2], in_k[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            jj[u] = j0 + u * blockDim.x;
            on[u] = jj[u] < (int)m;
            in_k[u] = (jj[u] / 6) == k;
#pragma unroll
            for (int q = 0; q < 6; ++q) Rq[u][q] = on[u] ? __ldcg(R + q * m + jj[u]) : 0.0;
          }
          if (!pivot_rows) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int i = 0; i < 6; ++i) a[u][i] = (on[u] && !in_k[u]) ? M[(size_t)(gi0 + i) * m + jj[u]] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (!on[u]) continue;
            double* col = M + (size_t)gi0 * m + jj[u];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              double t;
              if (pivot_rows) t = Rq[u][i];
              else {
                t = a[u][i];
#pragma unroll
                for (int q = 0; q < 6; ++q) t = fma(-Cs[rb + i][q], Rq[u][q], t);
              }
              col[(size_t)i * m] = t;
            }
          }
        }
      }
    };
    const int kn = k + 1;
    if (kn < nb && kn >= b0 && kn < b1) {
      const int lo = 6 * (kn - b0);
      update(lo, lo + 6);
      publish(kn);
      update(0, lo);
      update(lo + 6, nrows);
    } else {
      update(0, nrows);
    }
    grid_barrier(P.counter, epoch);
  }
}

// x = inv * r on a densely inverted level: one warp per row
__global__ void __launch_bounds__(kAmgThreads) amg_dense_solve_kernel(int m, const double* __restrict__ inv, const double* __restrict__ r,
                                                                      double* __restrict__ x, const int* skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (kAmgThreads / 32) + (threadIdx.x >> 5);
  if (i >= m) return;                           // warp-uniform
  const double* row = inv + (size_t)i * m;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = lane;
  for (; k + 96 < m; k += 128) {
    const double a0 = __ldg(row + k), a1 = __ldg(row + k + 32), a2 = __ldg(row + k + 64), a3 = __ldg(row + k + 96);
    s0 = fma(a0, __ldg(r + k), s0); s1 = fma(a1, __ldg(r + k + 32), s1);
    s2 = fma(a2, __ldg(r + k + 64), s2); s3 = fma(a3, __ldg(r + k + 96), s3);
  }
  for (; k < m; k += 32) s0 = fma(__ldg(row + k), __ldg(r + k), s0);
  const double s = warp_sum((s0 + s1) + (s2 + s3));
  if (lane == 0) x[i] = s;
}

// six lanes per block row, five rows per warp (the lane layout of bsr6_row)
struct RowLane {
  int i, c, g0;
  bool on;
};
__device__ __forceinline__ RowLane amg_row_lane(int n) {
  const int lane = threadIdx.x & 31, grp = lane / 6;
  RowLane L;
  L.c = lane - grp * 6;
  L.g0 = grp * 6;
  L.i = (blockIdx.x * (kAmgThreads / 32) + (threadIdx.x >> 5)) * kRowsPerWarp + grp;
  L.on = grp < kRowsPerWarp && L.i < n;
  return L;
}
// row c of a row-major 6x6 block times the 6-vector spread over the lanes of the group
// (TD = float: the fp32 copy of level 0's block inverses on large graphs -- the same copy in the pre- and the post-sweep)
template <typename TD = double>
__device__ __forceinline__ double amg_block_row_dot(const TD* blk, int c, int g0, double v) {
  const unsigned m = 0xffffffffu;
  const double v0 = __shfl_sync(m, v, g0), v1 = __shfl_sync(m, v, g0 + 1), v2 = __shfl_sync(m, v, g0 + 2);
  const double v3 = __shfl_sync(m, v, g0 + 3), v4 = __shfl_sync(m, v, g0 + 4), v5 = __shfl_sync(m, v, g0 + 5);
  if (blk == nullptr) return 0.0;
  const typename Pair2<TD>::type* b = reinterpret_cast<const typename Pair2<TD>::type*>(blk + 6 * c);
  const double2 m0 = ldg_pair<TD>(b), m1 = ldg_pair<TD>(b + 1), m2 = ldg_pair<TD>(b + 2);
  return m0.x * v0 + m0.y * v1 + m1.x * v2 + m1.y * v3 + m2.x * v4 + m2.y * v5;
}

// x = omega * Dinv r   (first smoothing sweep from a zero guess)
__global__ void __launch_bounds__(kAmgThreads) amg_smooth0_kernel(int n, const double* __restrict__ Dinv, const double* __restrict__ r,
                                                                  double omega, double* __restrict__ x, const int* skip) {
  if (skip && *skip) return;
  const RowLane L = amg_row_lane(n);
  const size_t q = 6 * (size_t)(L.on ? L.i : 0) + L.c;
  const double rv = L.on ? r[q] : 0.0;
  const double t = amg_block_row_dot(L.on ? Dinv + 36 * (size_t)L.i : nullptr, L.c, L.g0, rv);
  if (L.on) x[q] = omega * t;
}

// fp32 copy of an operator array (once per LM step and level)
__global__ void __launch_bounds__(256) amg_to_float_kernel(size_t n, const double* __restrict__ src, float* __restrict__ dst) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) dst[k] = (float)src[k];
}

// second visit of a level: r = t (the new right-hand side), x = omega * Dinv t
__global__ void __launch_bounds__(kAmgThreads) amg_rhs_smooth0_kernel(int n, const double* __restrict__ Dinv, const double* __restrict__ t,
                                                                      double omega, double* __restrict__ r, double* __restrict__ x, const int* skip) {
  if (skip && *skip) return;
  const RowLane L = amg_row_lane(n);
  const size_t q = 6 * (size_t)(L.on ? L.i : 0) + L.c;
  const double tv = L.on ? t[q] : 0.0;
  const double z = amg_block_row_dot(L.on ? Dinv + 36 * (size_t)L.i : nullptr, L.c, L.g0, tv);
  if (L.on) { r[q] = tv; x[q] = omega * z; }
}
// y += x
__global__ void amg_add_kernel(int n6, const double* __restrict__ x, double* __restrict__ y, const int* skip) {
  if (skip && *skip) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n6) y[k] += x[k];
}

// y = x + omega * Dinv (r - A x)
template <typename T, typename TD>
__global__ void __launch_bounds__(kAmgThreads) amg_smooth_kernel(const BsrViewT<T> A, const double* __restrict__ d, const TD* __restrict__ Dinv,
                                                                 const double* __restrict__ r, const double* __restrict__ x, double omega,
                                                                 double* __restrict__ y, const int* skip) {
  if (skip && *skip) return;
  // grid-stride over groups of five rows per warp, like spmv_kernel (the launch is a fixed 8 CTAs per SM)
  const int lane = threadIdx.x & 31, grp = lane / 6, c = lane - grp * 6, g0 = grp * 6;
  const int wpc = kAmgThreads / 32;
  const int gw = blockIdx.x * wpc + (threadIdx.x >> 5), nw = gridDim.x * wpc;
  for (int base = gw * kRowsPerWarp; base < A.n; base += nw * kRowsPerWarp) {
    const int i = base + grp;
    const bool on = grp < kRowsPerWarp && i < A.n;
    const size_t q = 6 * (size_t)(on ? i : 0) + c;
    double t = 0.0;
    if (on) t = r[q] - bsr6_row<false, T>(A.Hdiag, A.Hoff, A.row_ptr, A.col_idx, x, d, i, c);
    const double z = amg_block_row_dot<TD>(on ? Dinv + 36 * (size_t)i : nullptr, c, g0, t);
    if (on) y[q] = x[q] + omega * z;
  }
}

// t = r - A x over the stored rows (six lanes per row): the residual half of residual + restriction on LARGE levels.
// (A fused kernel -- one six-lane group walking its aggregate's member rows -- was measured at 515 us on the 1M-pose level
// against 360 us for a whole smoothing sweep: a long dependent chain per group; this is a streaming SpMV.)
template <typename T>
__global__ void __launch_bounds__(kAmgThreads) amg_residual_kernel(const BsrViewT<T> A, const double* __restrict__ d, const double* __restrict__ r,
                                                                   const double* __restrict__ x, double* __restrict__ t, const int* skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31, grp = lane / 6, c = lane - grp * 6;
  const int wpc = kAmgThreads / 32;
  const int gw = blockIdx.x * wpc + (threadIdx.x >> 5), nw = gridDim.x * wpc;
  for (int base = gw * kRowsPerWarp; base < A.n; base += nw * kRowsPerWarp) {
    const int i = base + grp;
    if (grp < kRowsPerWarp && i < A.n) {
      const size_t q = 6 * (size_t)i + c;
      t[q] = r[q] - bsr6_row<false, T>(A.Hdiag, A.Hoff, A.row_ptr, A.col_idx, x, d, i, c);
    }
  }
}

// x_i += P_i e_{agg(i)} over the stored rows
__global__ void __launch_bounds__(kAmgThreads) amg_prolong_kernel(int n, const int* __restrict__ agg, const double* __restrict__ pos,
                                                                  int pos_stride, const double* __restrict__ cpos,
                                                                  const double* __restrict__ scale, const double* __restrict__ ec,
                                                                  double* __restrict__ x, const int* skip) {
  if (skip && *skip) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t / 6, c = t - i * 6;
  if (i >= n) return;
  const int I = agg[i];
  if (I < 0) return;
  const double* e = ec + 6 * (size_t)I;
  double v = e[c];
  if (c < 3) {
    double dd[3];
    amg_delta(pos, pos_stride, i, cpos, I, dd);
    // X e_r = -2 d x e_r
    const double e3 = e[3], e4 = e[4], e5 = e[5];
    if (c == 0) v -= 2.0 * (dd[1] * e5 - dd[2] * e4);
    else if (c == 1) v -= 2.0 * (dd[2] * e3 - dd[0] * e5);
    else v -= 2.0 * (dd[0] * e4 - dd[1] * e3);
  }
  if (scale) { const double sv = scale[6 * (size_t)i + c]; v = sv > 0.0 ? v / sv : 0.0; }
  x[6 * (size_t)i + c] += v;
}

// ---- whole-warp-per-row forms for the levels that cannot fill the machine (<= kAmgWarpRowMax rows): there the cost of a
// sweep is the length of the longest dependent chain, and coarse Galerkin rows hold up to a few hundred blocks.  Five
// 6-lane groups stride through the row's blocks; a shuffle tree adds the five partial rows.  Result on lanes 0..5.
constexpr int kAmgWarpRowMax = 32768;
template <typename T>
__device__ __forceinline__ double bsr6_row_warp(const T* __restrict__ Hdiag, const T* __restrict__ Hoff,
                                                const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                                const double* __restrict__ x, const double* __restrict__ d, int i, int lane) {
  typedef typename Pair2<T>::type T2;
  const int grp = lane / 6, r = lane - grp * 6;
  double acc = 0.0;
  if (grp < 5) {
    if (grp == 0) {
      const T2* hd = reinterpret_cast<const T2*>(Hdiag + 36 * (size_t)i) + r;
      const double2 h0 = ldg_pair<T>(hd), h1 = ldg_pair<T>(hd + 6), h2 = ldg_pair<T>(hd + 12);
      const double* xi = x + 6 * (size_t)i;
      const double2 x0 = __ldg(reinterpret_cast<const double2*>(xi)), x1 = __ldg(reinterpret_cast<const double2*>(xi + 2)),
                    x2 = __ldg(reinterpret_cast<const double2*>(xi + 4));
      acc = h0.x * x0.x;
      acc = fma(h0.y, x0.y, acc); acc = fma(h1.x, x1.x, acc); acc = fma(h1.y, x1.y, acc);
      acc = fma(h2.x, x2.x, acc); acc = fma(h2.y, x2.y, acc);
      if (d != nullptr) {
        const double xr = (r == 0) ? x0.x : (r == 1) ? x0.y : (r == 2) ? x1.x : (r == 3) ? x1.y : (r == 4) ? x2.x : x2.y;
        acc = fma(__ldg(d + 6 * (size_t)i + r), xr, acc);
      }
    }
    const int p0 = __ldg(row_ptr + i), p1 = __ldg(row_ptr + i + 1);
    // Galerkin rows of the coarse levels hold up to a few hundred blocks: four blocks per trip with all their loads
    // (column ids first, then blocks and x gathers) in flight together, two accumulators
    double acc2 = 0.0;
    int p = p0 + grp;
    for (; p + 15 < p1; p += 20) {
      int j[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) j[q] = __ldg(col_idx + p + 5 * q);
      double2 hb[4][3], xb[4][3];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const T2* ha = reinterpret_cast<const T2*>(Hoff + 36 * (size_t)(p + 5 * q)) + r;
        hb[q][0] = ldg_pair<T>(ha); hb[q][1] = ldg_pair<T>(ha + 6); hb[q][2] = ldg_pair<T>(ha + 12);
        const double2* xa = reinterpret_cast<const double2*>(x + 6 * (size_t)j[q]);
        xb[q][0] = __ldg(xa); xb[q][1] = __ldg(xa + 1); xb[q][2] = __ldg(xa + 2);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double& t = (q & 1) ? acc2 : acc;
        t = fma(hb[q][0].x, xb[q][0].x, t); t = fma(hb[q][0].y, xb[q][0].y, t); t = fma(hb[q][1].x, xb[q][1].x, t);
        t = fma(hb[q][1].y, xb[q][1].y, t); t = fma(hb[q][2].x, xb[q][2].x, t); t = fma(hb[q][2].y, xb[q][2].y, t);
      }
    }
    for (; p < p1; p += 5) {
      const int j = __ldg(col_idx + p);
      const T2* ha = reinterpret_cast<const T2*>(Hoff + 36 * (size_t)p) + r;
      const double2 a0 = ldg_pair<T>(ha), a1 = ldg_pair<T>(ha + 6), a2 = ldg_pair<T>(ha + 12);
      const double* xa = x + 6 * (size_t)j;
      const double2 u0 = __ldg(reinterpret_cast<const double2*>(xa)), u1 = __ldg(reinterpret_cast<const double2*>(xa + 2)),
                    u2 = __ldg(reinterpret_cast<const double2*>(xa + 4));
      acc = fma(a0.x, u0.x, acc); acc = fma(a0.y, u0.y, acc); acc = fma(a1.x, u1.x, acc);
      acc = fma(a1.y, u1.y, acc); acc = fma(a2.x, u2.x, acc); acc = fma(a2.y, u2.y, acc);
    }
    acc += acc2;
  }
  const unsigned full = 0xffffffffu;
  const double a = acc + __shfl_down_sync(full, acc, 12);     // lane c: g0 + g2, lane c + 6: g1 + g3
  const double b = a + __shfl_down_sync(full, a, 6);          // lane c: g0 + g2 + g1 + g3
  return b + __shfl_down_sync(full, acc, 24);                 // + g4
}

// y = x + omega Dinv (r - A x), one warp per row; kResidualOnly: y = r - A x
template <bool kResidualOnly, typename T>
__global__ void __launch_bounds__(kAmgThreads) amg_smooth_warp_kernel(const BsrViewT<T> A, const double* __restrict__ d, const double* __restrict__ Dinv,
                                                                      const double* __restrict__ r, const double* __restrict__ x, double omega,
                                                                      double* __restrict__ y, const int* skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (kAmgThreads / 32) + (threadIdx.x >> 5);
  if (i >= A.n) return;                         // warp-uniform
  const double ax = bsr6_row_warp<T>(A.Hdiag, A.Hoff, A.row_ptr, A.col_idx, x, d, i, lane);
  const bool on = lane < 6;
  const size_t q = 6 * (size_t)i + (on ? lane : 0);
  const double t = on ? r[q] - ax : 0.0;
  if (kResidualOnly) { if (on) y[q] = t; return; }
  const double z = amg_block_row_dot(on ? Dinv + 36 * (size_t)i : nullptr, on ? lane : 0, 0, t);
  if (on) y[q] = x[q] + omega * z;
}

// rc_I = sum_{i in I} P_i^T t_i from a stored residual t (the residual is computed by amg_residual_kernel / amg_smooth_warp_kernel<true>): one warp per
// coarse row, five 6-lane groups stride through the members, shuffle tree, then the fused first sweep of the coarse level
__global__ void __launch_bounds__(kAmgThreads) amg_restrict_kernel(const double* __restrict__ t, int ncomp, int c_row0,
                                                                   const int* __restrict__ mem_ptr, const int* __restrict__ mem_idx,
                                                                   const double* __restrict__ pos, int pos_stride,
                                                                   const double* __restrict__ cpos, const double* __restrict__ scale,
                                                                   double* __restrict__ rc, const double* __restrict__ Dinv_c, double omega,
                                                                   double* __restrict__ xc, const int* skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31, grp = lane / 6, c = lane - grp * 6;
  const int k = blockIdx.x * (kAmgThreads / 32) + (threadIdx.x >> 5);
  if (k >= ncomp) return;                       // warp-uniform
  const int I = c_row0 + k;
  double acc = 0.0;
  if (grp < 5) {
    const int m0 = __ldg(mem_ptr + k), m1 = __ldg(mem_ptr + k + 1);
    for (int m = m0 + grp; m < m1; m += 5) {
      const int i = __ldg(mem_idx + m);
      const double* ti = t + 6 * (size_t)i;
      double s = 1.0, s0 = 1.0, s1 = 1.0, s2 = 1.0;
      if (scale) {
        const double* sc = scale + 6 * (size_t)i;
        s = sc[c] > 0.0 ? 1.0 / sc[c] : 0.0;
        s0 = sc[0] > 0.0 ? 1.0 / sc[0] : 0.0; s1 = sc[1] > 0.0 ? 1.0 / sc[1] : 0.0; s2 = sc[2] > 0.0 ? 1.0 / sc[2] : 0.0;
      }
      double v = s * ti[c];
      if (c >= 3) {
        double dd[3];
        amg_delta(pos, pos_stride, i, cpos, I, dd);
        const double u0 = s0 * ti[0], u1 = s1 * ti[1], u2 = s2 * ti[2];
        if (c == 3) v += 2.0 * (dd[1] * u2 - dd[2] * u1);
        else if (c == 4) v += 2.0 * (dd[2] * u0 - dd[0] * u2);
        else v += 2.0 * (dd[0] * u1 - dd[1] * u0);
      }
      acc += v;
    }
  }
  const unsigned full = 0xffffffffu;
  const double a = acc + __shfl_down_sync(full, acc, 12);
  const double b2 = a + __shfl_down_sync(full, a, 6);
  const double tot = b2 + __shfl_down_sync(full, acc, 24);
  const bool on = lane < 6;
  if (on) rc[6 * (size_t)I + lane] = tot;
  if (Dinv_c != nullptr) {
    const double z = amg_block_row_dot(on ? Dinv_c + 36 * (size_t)I : nullptr, on ? lane : 0, 0, on ? tot : 0.0);
    if (on) xc[6 * (size_t)I + lane] = omega * z;
  }
}

// PGO_AMG_PROFILE=1: device-side stage marks (capturable, one thread): acc[idx] += time since the previous mark
__global__ void amg_mark_kernel(unsigned long long* acc, int idx, const int* skip) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  if (idx >= 0 && !(skip && *skip)) acc[idx] += t - acc[63];
  acc[63] = t;
}

// ---- PCG pieces (Chronopoulos-Gear with a general preconditioner) ----
// init: x = 0, r = b, p = s = 0, x0 = omega Minv r (first pre-smoothing sweep of the first V-cycle)
template <typename TD>
__global__ void __launch_bounds__(kAmgThreads) amg_pcg_init_kernel(int n, const double* __restrict__ b, const TD* __restrict__ Minv,
                                                                   double omega, double* x, double* r, double* p, double* s, double* x0) {
  const RowLane L = amg_row_lane(n);
  const size_t q = 6 * (size_t)(L.on ? L.i : 0) + L.c;
  const double rv = L.on ? b[q] : 0.0;
  const double t = amg_block_row_dot<TD>(L.on ? Minv + 36 * (size_t)L.i : nullptr, L.c, L.g0, rv);
  if (L.on) { x[q] = 0.0; r[q] = rv; p[q] = 0.0; s[q] = 0.0; x0[q] = omega * t; }
}

// w = (A + D) u over the owned rows with per-CTA partials of w.u and r.u in a fixed order: a whole warp per row (small
// level 0: the sweep is latency bound).  Large level 0: the plain spmv_kernel with w.u in its epilogue + amg_dot_kernel.
__global__ void __launch_bounds__(kAmgThreads) amg_spmv_dots_warp_kernel(const BsrView A, const double* __restrict__ d, const double* __restrict__ u,
                                                                         const double* __restrict__ r, double* __restrict__ w,
                                                                         double* __restrict__ part /* [2][gridDim.x] */, const int* skip) {
  __shared__ double red0[kAmgThreads / 32], red1[kAmgThreads / 32];
  const bool idle = skip && *skip;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (kAmgThreads / 32) + (threadIdx.x >> 5);
  double a0 = 0.0, a1 = 0.0;
  if (!idle && i < A.n) {
    const double v = bsr6_row_warp<double>(A.Hdiag, A.Hoff, A.row_ptr, A.col_idx, u, d, i, lane);
    if (lane < 6) {
      const size_t q = 6 * (size_t)i + lane;
      w[q] = v;
      const double uv = __ldg(u + q);
      a0 = v * uv;
      a1 = __ldg(r + q) * uv;
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1);
  if (lane == 0) { red0[threadIdx.x >> 5] = a0; red1[threadIdx.x >> 5] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < kAmgThreads / 32; ++k) { t0 += red0[k]; t1 += red1[k]; }
    part[blockIdx.x] = t0;
    part[gridDim.x + blockIdx.x] = t1;
  }
}

// per-CTA partials of a . b in a fixed order (the r.u of the large-graph path: there the CG product is the plain spmv_kernel
// with w.u in its epilogue -- 0.27 ms on the 1M-pose level against 0.35 ms for the kernel that carries both sums)
__global__ void __launch_bounds__(kAmgThreads) amg_dot_kernel(int n6, const double* __restrict__ a, const double* __restrict__ b,
                                                              double* __restrict__ part, const int* skip) {
  __shared__ double red[kAmgThreads / 32];
  double acc = 0.0;
  if (!(skip && *skip))
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n6; k += gridDim.x * blockDim.x) acc = fma(a[k], b[k], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kAmgThreads / 32; ++k) t += red[k];
    part[blockIdx.x] = t;
  }
}

// sums nq quantities of per-CTA partials (part[q * nparts + k]) into out[q] in a fixed order (one CTA)
__global__ void __launch_bounds__(kAmgThreads) amg_reduce_kernel(const double* __restrict__ part, int nparts, int nq, double* __restrict__ out) {
  __shared__ double red[kAmgThreads / 32];
  for (int q = 0; q < nq; ++q) {
    double t = 0.0;
    for (int k = threadIdx.x; k < nparts; k += kAmgThreads) t += part[(size_t)q * nparts + k];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < kAmgThreads / 32; ++k) s += red[k];
      out[q] = s;
    }
    __syncthreads();
  }
}

// scalar step: delta = w.u, gamma = r.u
__device__ __forceinline__ void amg_pcg_scalar_step(PcgMultiState* st, double delta, double gamma, int max_iterations, double tol) {
  if (st->done) return;
  if (st->iter == 0) { st->gamma0 = gamma; st->gamma = gamma; }
  else { st->gamma_old = st->gamma; st->gamma = gamma; }
  if (!(st->gamma0 > 0.0)) { st->flag = 0; st->done = 1; return; }
  if (st->iter > 0 && gamma <= tol * tol * st->gamma0) { st->flag = 0; st->done = 1; return; }
  if (st->iter >= max_iterations) { st->flag = 1; st->done = 1; return; }
  st->delta = delta;
  if (st->iter == 0) { st->beta = 0.0; st->alpha = gamma / delta; }
  else { st->beta = gamma / st->gamma_old; st->alpha = gamma / (delta - st->beta * gamma / st->alpha); }
  if (!(st->alpha > 0.0) || !isfinite(st->alpha)) { st->flag = 2; st->done = 1; return; }
  st->iter++;
}
// multi-GPU: red[0], red[1] were all-reduced
__global__ void amg_pcg_scalar_kernel(PcgMultiState* st, const double* __restrict__ red, int max_iterations, double tol) {
  if (threadIdx.x == 0) amg_pcg_scalar_step(st, red[0], red[1], max_iterations, tol);
}
// multi-GPU over peer memory: wait for every rank's partial sums (pushed into slot [rank] of my window), add them in
// rank order -- bit-identical on every rank -- and take the scalar step (one warp)
__global__ void amg_pcg_scalar_peer_kernel(const PeerWaitArgs A, int world, int me, const double* __restrict__ mine, PcgMultiState* st,
                                           int max_iterations, double tol) {
  if (st->done) return;
  const unsigned int s = *A.seq;
  if (threadIdx.x < A.n_sources) {
    if (!peer_spin(A.flags + A.src_rank[threadIdx.x], s) && A.timeout_flag) *A.timeout_flag = 1;
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const double* stage = A.stage[s & 1];
    double d = 0.0, gm = 0.0;
    for (int r = 0; r < world; ++r) {
      d += r == me ? mine[0] : __ldcg(stage + 6 * (size_t)r);
      gm += r == me ? mine[1] : __ldcg(stage + 6 * (size_t)r + 1);
    }
    amg_pcg_scalar_step(st, d, gm, max_iterations, tol);
  }
}
// one GPU: sum the per-CTA partials (fixed order) and take the scalar step in the same one-CTA kernel
__global__ void __launch_bounds__(kAmgThreads) amg_pcg_reduce_scalar_kernel(PcgMultiState* st, const double* __restrict__ part, int nparts,
                                                                            int max_iterations, double tol) {
  __shared__ double red[2][kAmgThreads / 32];
  if (st->done) return;
  double t0 = 0.0, t1 = 0.0;
  for (int k = threadIdx.x; k < nparts; k += kAmgThreads) { t0 += part[k]; t1 += part[nparts + k]; }
  t0 = warp_sum(t0); t1 = warp_sum(t1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = t0; red[1][threadIdx.x >> 5] = t1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < kAmgThreads / 32; ++k) { s0 += red[0][k]; s1 += red[1][k]; }
    amg_pcg_scalar_step(st, s0, s1, max_iterations, tol);
  }
}

// p = u + beta p, s = w + beta s, x += alpha p, r -= alpha s, x0 = omega Minv r
template <typename TD>
__global__ void __launch_bounds__(kAmgThreads) amg_pcg_update_kernel(int n, const TD* __restrict__ Minv, const double* __restrict__ u,
                                                                     const double* __restrict__ w, double omega, double* x, double* r,
                                                                     double* p, double* s, double* x0, const PcgMultiState* st) {
  if (st->done) return;
  const RowLane L = amg_row_lane(n);
  const size_t q = 6 * (size_t)(L.on ? L.i : 0) + L.c;
  double rv = 0.0;
  if (L.on) {
    const double alpha = st->alpha, beta = st->beta;
    const double pv = u[q] + beta * p[q];
    const double sv = w[q] + beta * s[q];
    p[q] = pv; s[q] = sv;
    x[q] += alpha * pv;
    rv = r[q] - alpha * sv;
    r[q] = rv;
  }
  const double t = amg_block_row_dot<TD>(L.on ? Minv + 36 * (size_t)L.i : nullptr, L.c, L.g0, rv);
  if (L.on) x0[q] = omega * t;
}

// epilogue partials over the owned rows: x.b, x.w (w = A x), x.D x
__global__ void __launch_bounds__(kAmgThreads) amg_final_kernel(int n6, const double* __restrict__ x, const double* __restrict__ b,
                                                                const double* __restrict__ w, const double* __restrict__ d,
                                                                double* __restrict__ part /* [3][gridDim.x] */) {
  __shared__ double red[3][kAmgThreads / 32];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n6; k += gridDim.x * blockDim.x) {
    const double xv = x[k];
    a0 = fma(xv, b[k], a0); a1 = fma(xv, w[k], a1); a2 = fma(xv * xv, d[k], a2);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; red[2][threadIdx.x >> 5] = a2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kAmgThreads / 32; ++k) t += red[threadIdx.x][k];
    part[threadIdx.x * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void amg_final_store_kernel(const double* __restrict__ red, const PcgMultiState* st, DeviceScalars* sc) {
  if (threadIdx.x != 0) return;
  sc->xtb = red[0]; sc->xtAx = red[1]; sc->xtDx = red[2];
  sc->pcg_gamma0 = st->gamma0; sc->pcg_gamma = st->gamma; sc->pcg_iterations = st->iter; sc->pcg_flag = st->flag;
}

}  // namespace pgo

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static void amg_destroy(pgo::Amg* M, int device) {
  if (!M) return;
  pgo::pool_pinned_release(device, M->state_h);
  pgo::pool_event_release(device, M->ev[0]);
  pgo::pool_event_release(device, M->ev[1]);
  for (auto& e : M->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
  if (M->tail_graph) cudaGraphExecDestroy(M->tail_graph);
  if (M->peer) peer_destroy(M->owner, M->peer);
  delete M;   // device blocks were borrowed through dev_alloc and go back with the graph's
}

template <typename Tp>
static int amg_upload(pgo_graph* g, Tp** dst, const std::vector<Tp>& src) {
  PGO_TRY(dev_alloc(g, dst, src.size()));
  if (!src.empty()) CUDA_TRY(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(Tp), cudaMemcpyHostToDevice, g->stream));
  return PGO_OK;
}

static int amg_rows_grid(int n) { return std::max(1, (n + (pgo::kAmgThreads / 32) * pgo::kRowsPerWarp - 1) / ((pgo::kAmgThreads / 32) * pgo::kRowsPerWarp)); }

// Halo exchange of a per-node vector on a distributed level (see halo_exchange in pgo_b200.cu).
static int amg_exchange(pgo_graph* g, pgo::Amg* M, const pgo::AmgLevelDev& L, double* v, int width, int stride, const int* skip) {
  (void)M;
  if (g->world <= 1 || L.replicated || L.nbr.empty()) return PGO_OK;
  if (stride != width) return set_error(PGO_ERR_INVALID_ARGUMENT, "amg_exchange: strided vectors are not supported");
  if (M && M->peer && L.peer_ch >= 0 && width == 6) {
    const bool prof = M->prof && skip;   // PGO_AMG_PROFILE: [40] time before the exchange since the last mark, [41] push, [42] wait
    if (prof) pgo::amg_mark_kernel<<<1, 1, 0, g->stream>>>(M->prof, 40, skip);
    if (L.push.n_targets > 0) pgo::peer_push_kernel<<<dim3(L.push_ctas, L.push.n_targets), pgo::kPeerThreads, 0, g->stream>>>(L.push, L.send_idx, v, skip);
    if (prof) pgo::amg_mark_kernel<<<1, 1, 0, g->stream>>>(M->prof, 41, skip);
    if (L.wait.n_sources > 0) pgo::peer_wait_kernel<<<L.wait_ctas, pgo::kPeerThreads, 0, g->stream>>>(L.wait, v, skip);
    if (prof) pgo::amg_mark_kernel<<<1, 1, 0, g->stream>>>(M->prof, 42, skip);
    g->launches += 2;
    M->peer->pushes++; M->peer->push_bytes += (long long)L.send_ptr.back() * 48;
    return PGO_OK;
  }
  return halo_exchange(g, L.nbr, L.send_ptr, L.recv_ptr, L.send_idx, L.n_own, v, width, skip);
}

// Per-iteration all-gather of the residual of the first replicated level (rows computed rank by rank).
static int amg_gather(pgo_graph* g, double* v, const std::vector<int>& off, int unit);
static int amg_gather_residual(pgo_graph* g, pgo::Amg* M, const pgo::AmgLevelDev& C, const int* skip) {
  if (g->world <= 1 || C.gather_off.empty()) return PGO_OK;
  if (M->peer && C.gather_ch >= 0) {
    if (C.gpush.n_targets > 0) pgo::peer_push_kernel<<<dim3(C.gpush_ctas, C.gpush.n_targets), pgo::kPeerThreads, 0, g->stream>>>(C.gpush, nullptr, C.r, skip);
    if (C.gwait.n_sources > 0) pgo::peer_wait_kernel<<<C.gwait_ctas, pgo::kPeerThreads, 0, g->stream>>>(C.gwait, C.r, skip);
    g->launches += 2;
    M->peer->pushes++; M->peer->push_bytes += (long long)(C.gather_off[g->rank + 1] - C.gather_off[g->rank]) * 48 * (g->world - 1);
    return PGO_OK;
  }
  return amg_gather(g, C.r, C.gather_off, 6);
}

// All-gather of rank-owned ranges of a replicated array (`unit` doubles per item, item ranges off[r]..off[r+1]).
static int amg_gather(pgo_graph* g, double* v, const std::vector<int>& off, int unit) {
  if (g->world <= 1 || off.empty()) return PGO_OK;
  NCCL_TRY(ncclGroupStart());
  for (int r = 0; r < g->world; ++r) {
    const size_t cnt = (size_t)(off[r + 1] - off[r]) * unit;
    if (cnt == 0) continue;
    double* p = v + (size_t)off[r] * unit;
    NCCL_TRY(ncclBroadcast(p, p, cnt, ncclDouble, r, g->comm, g->stream));
  }
  NCCL_TRY(ncclGroupEnd());
  g->comm_calls++;
  g->comm_bytes += (long long)(off[g->rank + 1] - off[g->rank]) * unit * 8;
  return PGO_OK;
}

static int amg_allreduce(pgo_graph* g, double* buf, int count) { return allreduce_sum(g, buf, (size_t)count); }

// Peer-memory channels of the per-iteration exchanges: the halo of every distributed level, the residual gather of the
// first replicated level, the scalar all-reduce.  Collective; leaves M->peer == nullptr when windows cannot be mapped.
static int amg_peer_setup(pgo_graph* g, pgo::Amg* M) {
  using namespace pgo;
  const int nl = M->num_levels, W = g->world, me = g->rank;
  std::vector<PeerChannelSpec> spec;
  bool symmetric = true;
  for (int l = 0; l < nl; ++l) {
    AmgLevelDev& D = M->lv[l];
    if (!D.replicated) {
      PeerChannelSpec c;
      c.stage_items = (size_t)D.n_halo;
      c.recv_item_off.assign(W, -1);
      for (size_t k = 0; k < D.nbr.size(); ++k)
        if (D.recv_ptr[k + 1] > D.recv_ptr[k]) c.recv_item_off[D.nbr[k]] = D.recv_ptr[k];
      // (the argument that two staging buffers are enough needs symmetric partners; a rank that sees otherwise vetoes
      // inside peer_create -- the decision must be collective, every rank is in its all-gather)
      for (size_t k = 0; k < D.nbr.size(); ++k)
        if ((D.recv_ptr[k + 1] > D.recv_ptr[k]) != (D.send_ptr[k + 1] > D.send_ptr[k])) symmetric = false;
      D.peer_ch = (int)spec.size();
      spec.push_back(c);
    }
    if (!D.gather_off.empty()) {
      PeerChannelSpec c;
      c.stage_items = (size_t)D.n_own;
      c.recv_item_off.assign(W, -1);
      for (int r = 0; r < W; ++r) if (r != me && D.gather_off[r + 1] > D.gather_off[r]) c.recv_item_off[r] = D.gather_off[r];
      D.gather_ch = (int)spec.size();
      spec.push_back(c);
    }
  }
  PeerChannelSpec red;
  red.stage_items = (size_t)W;
  red.recv_item_off.assign(W, -1);
  for (int r = 0; r < W; ++r) if (r != me) red.recv_item_off[r] = r;
  const int ch_red = (int)spec.size();
  spec.push_back(red);
  PGO_TRY(peer_create(g, spec, &M->peer, symmetric));
  if (!M->peer) {
    for (auto& D : M->lv) { D.peer_ch = -1; D.gather_ch = -1; }
    return PGO_OK;
  }
  const PeerCtx* P = M->peer;
  auto ctas_for = [](long long doubles) { return (int)std::max<long long>(1, std::min<long long>(64, (doubles + 4 * kPeerThreads - 1) / (4 * kPeerThreads))); };
  int* timeout = &M->state->pad;
  for (int l = 0; l < nl; ++l) {
    AmgLevelDev& D = M->lv[l];
    if (D.peer_ch >= 0) {
      D.push = peer_push_args(P, D.peer_ch, D.nbr, std::vector<int>(D.send_ptr.begin(), D.send_ptr.end() - 1),
                              std::vector<int>(D.send_ptr.begin() + 1, D.send_ptr.end()));
      std::vector<int> src;
      for (size_t k = 0; k < D.nbr.size(); ++k) if (D.recv_ptr[k + 1] > D.recv_ptr[k]) src.push_back(D.nbr[k]);
      D.wait = peer_wait_args(P, D.peer_ch, src);
      D.wait.timeout_flag = timeout;
      D.wait.n_copies = 1; D.wait.copy_src[0] = 0; D.wait.copy_dst[0] = 6 * D.n_own; D.wait.copy_n[0] = 6 * D.n_halo;
      {
        long long most = 0;
        for (size_t k = 0; k + 1 < D.send_ptr.size(); ++k) most = std::max<long long>(most, D.send_ptr[k + 1] - D.send_ptr[k]);
        D.push_ctas = (int)std::max<long long>(1, std::min<long long>(16, (6 * most + 8 * kPeerThreads - 1) / (8 * kPeerThreads)));
      }
      D.wait_ctas = ctas_for(6ll * D.n_halo);
    }
    if (D.gather_ch >= 0) {
      std::vector<int> others, src;
      for (int r = 0; r < W; ++r) {
        if (r == me) continue;
        others.push_back(r);
        if (D.gather_off[r + 1] > D.gather_off[r]) src.push_back(r);
      }
      D.gpush = peer_push_args(P, D.gather_ch, others, std::vector<int>(others.size(), D.gather_off[me]),
                               std::vector<int>(others.size(), D.gather_off[me + 1]));
      D.gwait = peer_wait_args(P, D.gather_ch, src);
      D.gwait.timeout_flag = timeout;
      D.gwait.n_copies = 2;
      D.gwait.copy_src[0] = 0; D.gwait.copy_dst[0] = 0; D.gwait.copy_n[0] = 6 * D.gather_off[me];
      D.gwait.copy_src[1] = 6 * D.gather_off[me + 1]; D.gwait.copy_dst[1] = 6 * D.gather_off[me + 1];
      D.gwait.copy_n[1] = 6 * (D.n_own - D.gather_off[me + 1]);
      D.gpush_ctas = (int)std::max<long long>(1, std::min<long long>(16, (6ll * (D.gather_off[me + 1] - D.gather_off[me]) + 8 * kPeerThreads - 1) / (8 * kPeerThreads)));
      D.gwait_ctas = ctas_for(6ll * D.n_own);
    }
  }
  {
    std::vector<int> others;
    for (int r = 0; r < W; ++r) if (r != me) others.push_back(r);
    M->rpush = peer_push_args(P, ch_red, others, std::vector<int>(others.size(), 0), std::vector<int>(others.size(), 1));
    M->rwait = peer_wait_args(P, ch_red, others);
    M->rwait.timeout_flag = timeout;
  }
  return PGO_OK;
}

// Build the hierarchy for this graph (host analysis + uploads).  pos0: [N_global][3] setup-time positions.
static int amg_create(pgo_graph* g, pgo::Amg** out) {
  using namespace pgo;
  Amg* M = new Amg();
  *out = M;
  const AmgHostParams prm = amg_host_params(g->N_global);
  if (const char* e = getenv("PGO_AMG_OMEGA")) M->omega = atof(e);
  if (const char* e = getenv("PGO_AMG_NU")) M->nu = std::max(1, atoi(e));
  if (const char* e = getenv("PGO_AMG_COARSE_SWEEPS")) M->coarse_sweeps = std::max(1, atoi(e));
  if (const char* e = getenv("PGO_AMG_FP32")) M->fp32_ops = atoi(e) != 0;
  // cycle shape: a second visit of the first two coarse levels pays when level 0 is a long streaming sweep (>= 400 k rows
  // here: 1M-pose grid 208 -> 120 PCG iterations per LM step, 5.2 -> 4.4 s); on smaller slices the extra coarse visits are
  // pure launch latency and the V-cycle is faster (100 k torus: 354 ms V, 452 ms W) -- profiles/r2q_w_cycle_ab_1xB200.txt
  if (g->n_own >= 400000) { M->gamma = 2; M->gamma_depth = 2; }
  if (const char* e = getenv("PGO_AMG_GAMMA")) { M->gamma = std::max(1, atoi(e)); M->gamma_depth = M->gamma >= 2 ? 99 : 0; }
  if (const char* e = getenv("PGO_AMG_GAMMA_DEPTH")) M->gamma_depth = std::max(0, atoi(e));
  std::vector<AmgGlobalLevel> G;
  std::vector<AmgLocalLevel> Lh;
  const double t0 = wall_s();
  PGO_TRY(graph_amg_hierarchy(g, prm, &G, &Lh));
  const double t1 = wall_s();
  const int nl = (int)Lh.size();
  M->num_levels = nl;
  M->lv.assign(nl, AmgLevelDev());
  size_t max_send = 0;
  for (int l = 0; l < nl; ++l) {
    const AmgLocalLevel& H = Lh[l];
    AmgLevelDev& D = M->lv[l];
    D.replicated = H.replicated;
    D.n_own = H.n_own; D.n_halo = H.n_halo;
    const size_t n_loc = (size_t)H.n_own + H.n_halo;
    if (l == 0) {
      D.row_ptr = g->row_ptr; D.col_idx = g->col_idx; D.nnz = g->nnz_off;
      D.pos_stride = 8;
    } else {
      D.nnz = (long long)H.col_idx.size();
      PGO_TRY(amg_upload(g, &D.row_ptr, H.row_ptr));
      PGO_TRY(amg_upload(g, &D.col_idx, H.col_idx));
      PGO_TRY(dev_alloc(g, &D.Adiag, (size_t)H.n_own * 36));
      PGO_TRY(dev_alloc(g, &D.Aoff, (size_t)D.nnz * 36));
      PGO_TRY(dev_alloc(g, &D.Dinv, (size_t)H.n_own * 36));
      PGO_TRY(dev_alloc(g, &D.r, n_loc * 6));
      PGO_TRY(dev_alloc(g, &D.pos, n_loc * 3));
      D.pos_stride = 3;
    }
    if (M->fp32_ops) {
      PGO_TRY(dev_alloc(g, &D.Adiag_f, (size_t)H.n_own * 36));
      PGO_TRY(dev_alloc(g, &D.Aoff_f, (size_t)D.nnz * 36));
    }
    PGO_TRY(dev_alloc(g, &D.x, n_loc * 6));
    PGO_TRY(dev_alloc(g, &D.y, n_loc * 6));
    if (l > 0) {
      PGO_TRY(dev_alloc(g, &D.z, n_loc * 6));
      CUDA_TRY(cudaMemsetAsync(D.z, 0, std::max<size_t>(n_loc * 6, 1) * sizeof(double), g->stream));
    }
    CUDA_TRY(cudaMemsetAsync(D.x, 0, std::max<size_t>(n_loc * 6, 1) * sizeof(double), g->stream));
    CUDA_TRY(cudaMemsetAsync(D.y, 0, std::max<size_t>(n_loc * 6, 1) * sizeof(double), g->stream));
    if (l + 1 < nl) {
      PGO_TRY(amg_upload(g, &D.agg, H.agg));
      PGO_TRY(amg_upload(g, &D.mem_ptr, H.mem_ptr));
      PGO_TRY(amg_upload(g, &D.mem_idx, H.mem_idx));
      PGO_TRY(amg_upload(g, &D.gal_ptr, H.gal_ptr));
      PGO_TRY(amg_upload(g, &D.gal_row, H.gal_row));
      PGO_TRY(amg_upload(g, &D.gal_slot, H.gal_slot));
      D.c_row0 = H.c_row0; D.c_row1 = H.c_row1;
      D.c_slot0 = Lh[l + 1].row_ptr[H.c_row0];
      D.n_cblk = (int)H.gal_ptr.size() - 1;
    }
    D.nbr = H.nbr; D.send_ptr = H.send_ptr; D.recv_ptr = H.recv_ptr;
    D.gather_off = H.gather_off; D.gather_slot_off = H.gather_slot_off;
    if (l == 0) D.send_idx = g->send_idx;
    else if (!H.send_idx.empty()) PGO_TRY(amg_upload(g, &D.send_idx, H.send_idx));
    if (l > 0) max_send = std::max(max_send, H.send_idx.size());
    M->blocks_all_levels += (long long)H.n_own + (l == 0 ? g->nnz_off : (long long)H.col_idx.size());
  }
  // the graph's send buffer (level-0 boundary x 8 doubles) serves every level: a coarse boundary node holds at least one
  // fine boundary node, so the lists only shrink -- checked, not assumed
  if (max_send > g->send_idx_h.size()) return set_error(PGO_ERR_NUMERICAL, "amg: a coarse level sends more nodes than level 0");
  const AmgLevelDev& last = M->lv[nl - 1];
  if (nl > 1 && (g->world == 1 || last.replicated) && last.n_own <= kAmgDenseMaxNodes) {
    PGO_TRY(dev_alloc(g, &M->dense_inv, (size_t)36 * last.n_own * last.n_own));
    if (last.n_own > kAmgDenseSmemNodes) {
      PGO_TRY(dev_alloc(g, &M->gj_rows, (size_t)2 * 36 * last.n_own));
      PGO_TRY(dev_alloc(g, &M->gj_counter, 1));
    }
  }
  PGO_TRY(dev_alloc(g, &M->state, 1));
  CUDA_TRY(pool_pinned(g->device, reinterpret_cast<void**>(&M->state_h)));
  static_assert(2 * sizeof(PcgMultiState) <= kPinnedBytes, "pinned PCG state slots");
  CUDA_TRY(pool_event(g->device, &M->ev[0]));
  CUDA_TRY(pool_event(g->device, &M->ev[1]));
  M->owner = g;
  if (M->fp32_ops && nl > 1 && M->lv[0].n_own > kAmgWarpRowMax) PGO_TRY(dev_alloc(g, &M->Minv_f, (size_t)M->lv[0].n_own * 36));
  if (getenv("PGO_AMG_PROFILE")) {
    PGO_TRY(dev_alloc(g, &M->prof, 64));
    CUDA_TRY(cudaMemsetAsync(M->prof, 0, 64 * sizeof(unsigned long long), g->stream));
  }
  if (g->world > 1) PGO_TRY(amg_peer_setup(g, M));
  M->part_cap = 3 * std::max(8 * g->num_sms, 1);
  PGO_TRY(dev_alloc(g, &M->part, (size_t)M->part_cap));
  PGO_TRY(dev_alloc(g, &M->red, 8));
  // communication per PCG iteration (payload sent by this rank): exchanges of the V-cycle + the SpMV + gathers + scalars
  if (g->world > 1) {
    long long bytes = 16; int calls = 1;
    for (int l = 0; l < nl; ++l) {
      const AmgLevelDev& D = M->lv[l];
      if (!D.nbr.empty()) {
        const int per_cycle = (l + 1 < nl ? 2 * M->nu : M->coarse_sweeps - 1) + (l == 0 ? 1 : 0);
        bytes += (long long)per_cycle * D.send_ptr.back() * 48; calls += per_cycle;
      }
      if (!D.gather_off.empty()) { bytes += (long long)(D.gather_off[g->rank + 1] - D.gather_off[g->rank]) * 48; calls += 1; }
    }
    M->comm_bytes_per_iteration = bytes; M->comm_calls_per_iteration = calls;
  }
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if (getenv("PGO_PROFILE_HOST")) {
    fprintf(stderr, "[pgo amg] rank %d: hierarchy %.1f ms, upload %.1f ms, levels:", g->rank, 1e3 * (t1 - t0), 1e3 * (wall_s() - t1));
    for (int l = 0; l < nl; ++l) fprintf(stderr, " %d%s(+%d halo, %lld blk)", M->lv[l].n_own, M->lv[l].replicated ? "R" : "", M->lv[l].n_halo, M->lv[l].nnz);
    fprintf(stderr, "%s\n", M->dense_inv ? " dense coarsest" : " smoothed coarsest");
  }
  return PGO_OK;
}

static pgo::BsrViewT<float> amg_view_f(const pgo::AmgLevelDev& D) {
  pgo::BsrViewT<float> A;
  A.n = D.n_own; A.Hdiag = D.Adiag_f; A.Hoff = D.Aoff_f; A.row_ptr = D.row_ptr; A.col_idx = D.col_idx;
  return A;
}
static pgo::BsrViewT<double> amg_view_d(const pgo::AmgLevelDev& D) {
  pgo::BsrViewT<double> A;
  A.n = D.n_own; A.Hdiag = D.Adiag; A.Hoff = D.Aoff; A.row_ptr = D.row_ptr; A.col_idx = D.col_idx;
  return A;
}
static long long amg_peer_pushes(const pgo_graph* g) { return g->amg && g->amg->peer ? g->amg->peer->pushes : 0; }
static long long amg_peer_bytes(const pgo_graph* g) { return g->amg && g->amg->peer ? g->amg->peer->push_bytes : 0; }

static pgo::BsrView amg_view(const pgo::AmgLevelDev& D) {
  pgo::BsrView A;
  A.n = D.n_own; A.Hdiag = D.Adiag; A.Hoff = D.Aoff; A.row_ptr = D.row_ptr; A.col_idx = D.col_idx;
  return A;
}

// Numeric setup for the current H, D and poses: centroids, Galerkin operators, block inverses, coarsest inverse.
static int amg_setup_numeric(pgo_graph* g, pgo::Amg* M) {
  using namespace pgo;
  const int nl = M->num_levels;
  AmgLevelDev& L0 = M->lv[0];
  L0.Adiag = g->Hdiag; L0.Aoff = g->Hoff; L0.Dinv = g->Minv; L0.pos = g->poses; L0.r = g->vr;
  // fp32 copies of the operators the cycle sweeps over (the CG product and the Galerkin products keep reading fp64)
  auto to_float = [&](const AmgLevelDev& D) {
    if (!M->fp32_ops) return;
    const size_t nd = (size_t)D.n_own * 36, no = (size_t)D.nnz * 36;
    if (nd) amg_to_float_kernel<<<(int)std::min<size_t>((nd + 255) / 256, (size_t)8 * g->num_sms), 256, 0, g->stream>>>(nd, D.Adiag, D.Adiag_f);
    if (no) amg_to_float_kernel<<<(int)std::min<size_t>((no + 255) / 256, (size_t)8 * g->num_sms), 256, 0, g->stream>>>(no, D.Aoff, D.Aoff_f);
    g->launches += 2;
  };
  if (nl > 1) to_float(L0);
  if (M->Minv_f) {
    const size_t nd = (size_t)L0.n_own * 36;
    amg_to_float_kernel<<<(int)std::min<size_t>((nd + 255) / 256, (size_t)8 * g->num_sms), 256, 0, g->stream>>>(nd, g->Minv, M->Minv_f);
    g->launches++;
  }
  for (int l = 0; l + 1 < nl; ++l) {
    AmgLevelDev& F = M->lv[l];
    AmgLevelDev& C = M->lv[l + 1];
    const int ncomp = F.c_row1 - F.c_row0;
    if (ncomp > 0) {
      amg_centroid_kernel<<<(ncomp + 255) / 256, 256, 0, g->stream>>>(ncomp, F.c_row0, F.mem_ptr, F.mem_idx, F.pos, F.pos_stride, C.pos);
      g->launches++;
    }
    if (!C.gather_off.empty()) PGO_TRY(amg_gather(g, C.pos, C.gather_off, 3));
    else PGO_TRY(amg_exchange(g, M, C, C.pos, 3, 3, nullptr));
    if (F.n_cblk > 0) {
      GalerkinParams P;
      P.n_cblk = F.n_cblk; P.ncomp = ncomp; P.c_row0 = F.c_row0; P.c_slot0 = F.c_slot0;
      P.gal_ptr = F.gal_ptr; P.gal_row = F.gal_row; P.gal_slot = F.gal_slot;
      P.Adiag = F.Adiag; P.Aoff = F.Aoff; P.col_idx = F.col_idx;
      P.dlm = l == 0 ? g->dlm : nullptr; P.scale = l == 0 ? g->scale : nullptr;
      P.agg = F.agg; P.pos = F.pos; P.pos_stride = F.pos_stride; P.cpos = C.pos;
      P.Cdiag = C.Adiag; P.Coff = C.Aoff;
      amg_galerkin_kernel<<<amg_rows_grid(F.n_cblk), kAmgThreads, 0, g->stream>>>(P);
      g->launches++;
    }
    if (!C.gather_off.empty()) {
      PGO_TRY(amg_gather(g, C.Adiag, C.gather_off, 36));
      PGO_TRY(amg_gather(g, C.Aoff, C.gather_slot_off, 36));
    }
    if (!(l + 2 == nl && M->dense_inv)) to_float(C);
    if (l + 2 == nl && M->dense_inv && C.n_own > kAmgDenseSmemNodes) {
      const size_t m = 6 * (size_t)C.n_own;
      CUDA_TRY(cudaMemsetAsync(M->dense_inv, 0, m * m * sizeof(double), g->stream));
      CUDA_TRY(cudaMemsetAsync(M->gj_counter, 0, sizeof(unsigned int), g->stream));
      amg_dense_fill_kernel<<<C.n_own, kAmgThreads, 0, g->stream>>>(C.n_own, C.Adiag, C.Aoff, C.row_ptr, C.col_idx, M->dense_inv);
      GjParams P;
      P.nb = C.n_own; P.M = M->dense_inv; P.R = M->gj_rows; P.counter = M->gj_counter;
      const int per = std::max((C.n_own + g->num_sms - 1) / g->num_sms, 1);
      if (6 * per > kGjMaxOwnRows) return set_error(PGO_ERR_NUMERICAL, "amg: dense coarsest level of %d nodes on %d SMs", C.n_own, g->num_sms);
      const int grid = (C.n_own + per - 1) / per;
      void* args[] = {&P};
      CUDA_TRY(cudaLaunchCooperativeKernel((void*)amg_dense_gj_kernel, dim3(grid), dim3(kGjThreads), args, 0, g->stream));
      g->launches++;
    } else if (l + 2 == nl && M->dense_inv) {
      const int m = 6 * C.n_own;
      const int smem = m * m * (int)sizeof(double);
      int have = 0;
      if (!pool_cache_get(g->device, kCacheAmgDenseAttr, &have)) {
        CUDA_TRY(cudaFuncSetAttribute(amg_dense_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      36 * kAmgDenseSmemNodes * kAmgDenseSmemNodes * (int)sizeof(double)));
        pool_cache_set(g->device, kCacheAmgDenseAttr, 1);
      }
      amg_dense_inverse_kernel<<<1, kAmgThreads, smem, g->stream>>>(C.n_own, C.Adiag, C.Aoff, C.row_ptr, C.col_idx, M->dense_inv);
    } else {
      amg_block_inverse_kernel<<<(C.n_own + 127) / 128, 128, 0, g->stream>>>(C.n_own, C.Adiag, C.Dinv);
    }
    g->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return PGO_OK;
}

static inline void amg_mark(pgo_graph* g, pgo::Amg* M, int idx) {
  if (M->prof) pgo::amg_mark_kernel<<<1, 1, 0, g->stream>>>(M->prof, idx, &M->state->done);
}

// y = x + omega Dinv (r - A x) / t = r - A x on level l: streaming six-lanes-per-row form on large levels, a warp per row on
// small ones; the operator is read from its fp32 copy when the cycle runs in mixed precision (vectors and arithmetic: fp64)
template <typename T>
static void amg_launch_smooth_t(pgo_graph* g, pgo::Amg* M, int l, const pgo::BsrViewT<T>& A, const double* x, double* y, const int* skip) {
  using namespace pgo;
  const AmgLevelDev& D = M->lv[l];
  if (D.n_own <= kAmgWarpRowMax)
    amg_smooth_warp_kernel<false, T><<<(D.n_own + 7) / 8, kAmgThreads, 0, g->stream>>>(A, l == 0 ? g->dlm : nullptr, D.Dinv, D.r, x, M->omega, y, skip);
  else if (l == 0 && M->Minv_f)
    amg_smooth_kernel<T, float><<<std::min(amg_rows_grid(D.n_own), 8 * g->num_sms), kAmgThreads, 0, g->stream>>>(A, g->dlm, M->Minv_f, D.r, x, M->omega, y, skip);
  else
    amg_smooth_kernel<T, double><<<std::min(amg_rows_grid(D.n_own), 8 * g->num_sms), kAmgThreads, 0, g->stream>>>(A, l == 0 ? g->dlm : nullptr, D.Dinv, D.r, x, M->omega, y, skip);
  g->launches++;
}
template <typename T>
static void amg_launch_residual_t(pgo_graph* g, pgo::Amg* M, int l, const pgo::BsrViewT<T>& A, const double* x, double* t, const int* skip) {
  using namespace pgo;
  const AmgLevelDev& D = M->lv[l];
  if (D.n_own <= kAmgWarpRowMax)
    amg_smooth_warp_kernel<true, T><<<(D.n_own + 7) / 8, kAmgThreads, 0, g->stream>>>(A, l == 0 ? g->dlm : nullptr, nullptr, D.r, x, 0.0, t, skip);
  else
    amg_residual_kernel<T><<<std::min(amg_rows_grid(D.n_own), 8 * g->num_sms), kAmgThreads, 0, g->stream>>>(A, l == 0 ? g->dlm : nullptr, D.r, x, t, skip);
  g->launches++;
}
static void amg_launch_smooth(pgo_graph* g, pgo::Amg* M, int l, const double* x, double* y, const int* skip) {
  if (M->fp32_ops) amg_launch_smooth_t<float>(g, M, l, amg_view_f(M->lv[l]), x, y, skip);
  else amg_launch_smooth_t<double>(g, M, l, amg_view_d(M->lv[l]), x, y, skip);
}
static void amg_launch_residual(pgo_graph* g, pgo::Amg* M, int l, const double* x, double* t, const int* skip) {
  if (M->fp32_ops) amg_launch_residual_t<float>(g, M, l, amg_view_f(M->lv[l]), x, t, skip);
  else amg_launch_residual_t<double>(g, M, l, amg_view_d(M->lv[l]), x, t, skip);
}

// One smoothing sweep y = x + omega Dinv (r - A x) on level l (halo of x exchanged first).
static int amg_sweep(pgo_graph* g, pgo::Amg* M, int l, double* x, double* y, const int* skip) {
  using namespace pgo;
  const AmgLevelDev& D = M->lv[l];
  PGO_TRY(amg_exchange(g, M, D, x, 6, 6, skip));
  amg_launch_smooth(g, M, l, x, y, skip);
  return PGO_OK;
}

// u = M^-1 r: one multigrid cycle.  Level 0: r = the CG residual (g->vr) and lv[0].x already holds omega Minv r.
// The result is a level-0 vector (owned rows valid); *out points at it.
//
// Cycle shape.  gamma == 1: V(nu, nu).  gamma == 2: levels 1 .. gamma_depth are visited TWICE from their parent (a
// W-cycle truncated at gamma_depth): after the first coarse correction e = M r_c the parent asks for a second one on
// r_c - A_c e and adds it -- the coarse operator applied is 2M - M A_c M, symmetric like M.  An aggregation cycle with
// piecewise-rigid (unsmoothed) coarse spaces loses a constant factor per level (the pose-graph Hessian bends like a
// plate: translations follow the integral of the rotations), and the second visit buys most of it back:
// tools/amg_prototype.py --gamma 2, 400 x 400 grid 53 -> 34 iterations, and the 1M-pose grid below.
static int amg_vcycle(pgo_graph* g, pgo::Amg* M, double** out) {
  using namespace pgo;
  const int nl = M->num_levels;
  const int* skip = &M->state->done;
  std::vector<double*> cur(nl), oth(nl);
  for (int l = 0; l < nl; ++l) { cur[l] = M->lv[l].x; oth[l] = M->lv[l].y; }
  // going down from level l: remaining pre-smoothing sweeps, residual, restriction (+ first sweep of level l + 1)
  auto down = [&](int l) -> int {
    const AmgLevelDev& D = M->lv[l];
    const AmgLevelDev& C = M->lv[l + 1];
    // (x_l = omega Dinv r_l arrives with r_l: from the PCG update on level 0, from the restriction below otherwise)
    if (l > 0 && !M->lv[l].gather_off.empty()) {
      amg_smooth0_kernel<<<amg_rows_grid(D.n_own), kAmgThreads, 0, g->stream>>>(D.n_own, D.Dinv, D.r, M->omega, cur[l], skip);
      g->launches++;
    }
    for (int s = 1; s < M->nu; ++s) { PGO_TRY(amg_sweep(g, M, l, cur[l], oth[l], skip)); std::swap(cur[l], oth[l]); }
    PGO_TRY(amg_exchange(g, M, D, cur[l], 6, 6, skip));
    const int ncomp = D.c_row1 - D.c_row0;
    // the next level's first sweep is fused unless its residual still has to be gathered or it is solved densely
    const bool coarsest_next = l + 2 == nl;
    const bool fuse_next = C.gather_off.empty() && !(coarsest_next && M->dense_inv);
    cur[l + 1] = C.x; oth[l + 1] = C.y;
    if (ncomp > 0) {
      // residual into the spare buffer (small level: a whole warp per row; large level: streaming, six lanes per row),
      // then the restriction (a warp per coarse row)
      amg_launch_residual(g, M, l, cur[l], oth[l], skip);
      amg_restrict_kernel<<<(ncomp + 7) / 8, kAmgThreads, 0, g->stream>>>(oth[l], ncomp, D.c_row0, D.mem_ptr, D.mem_idx, D.pos, D.pos_stride, C.pos,
                                                                          l == 0 ? g->scale : nullptr, C.r, fuse_next ? C.Dinv : nullptr,
                                                                          M->omega, cur[l + 1], skip);
      g->launches += 1;
    }
    PGO_TRY(amg_gather_residual(g, M, C, skip));
    return PGO_OK;
  };
  auto coarsest = [&]() -> int {
    const int l = nl - 1;
    const AmgLevelDev& D = M->lv[l];
    if (nl > 1 && M->dense_inv) {
      amg_dense_solve_kernel<<<(6 * D.n_own + 7) / 8, kAmgThreads, 0, g->stream>>>(6 * D.n_own, M->dense_inv, D.r, cur[l], skip);
      g->launches++;
    } else if (nl > 1) {
      if (!D.gather_off.empty()) {
        amg_smooth0_kernel<<<amg_rows_grid(D.n_own), kAmgThreads, 0, g->stream>>>(D.n_own, D.Dinv, D.r, M->omega, cur[l], skip);
        g->launches++;
      }
      for (int s = 1; s < M->coarse_sweeps; ++s) { PGO_TRY(amg_sweep(g, M, l, cur[l], oth[l], skip)); std::swap(cur[l], oth[l]); }
    }
    // nl == 1: plain block-Jacobi, lv[0].x = omega Minv r is the result (omega only rescales the preconditioner)
    return PGO_OK;
  };
  // coming back up to level l: prolongation, post-smoothing
  auto up = [&](int l) -> int {
    const AmgLevelDev& D = M->lv[l];
    const AmgLevelDev& C = M->lv[l + 1];
    amg_prolong_kernel<<<(D.n_own * 6 + 255) / 256, 256, 0, g->stream>>>(D.n_own, D.agg, D.pos, D.pos_stride, C.pos, l == 0 ? g->scale : nullptr,
                                                                         cur[l + 1], cur[l], skip);
    g->launches++;
    for (int s = 0; s < M->nu; ++s) { PGO_TRY(amg_sweep(g, M, l, cur[l], oth[l], skip)); std::swap(cur[l], oth[l]); }
    return PGO_OK;
  };
  // second visit of level l (its first correction e = cur[l] stays where it is): r_l <- r_l - A_l e, a fresh cycle on
  // it in the level's two other buffers, e += its result
  std::function<int(int)> cycle;
  auto revisit = [&](int l) -> int {
    const AmgLevelDev& D = M->lv[l];
    double* e = cur[l];
    double* b1 = e == D.x ? D.y : D.x;
    double* b2 = e == D.z ? D.y : D.z;
    PGO_TRY(amg_exchange(g, M, D, e, 6, 6, skip));
    amg_launch_residual(g, M, l, e, b2, skip);
    amg_rhs_smooth0_kernel<<<amg_rows_grid(D.n_own), kAmgThreads, 0, g->stream>>>(D.n_own, D.Dinv, b2, M->omega, D.r, b1, skip);
    g->launches += 1;
    cur[l] = b1; oth[l] = b2;
    PGO_TRY(cycle(l));
    amg_add_kernel<<<(6 * D.n_own + 255) / 256, 256, 0, g->stream>>>(6 * D.n_own, e, cur[l], skip);
    g->launches++;
    return PGO_OK;
  };
  // one cycle from level l down and back: on entry lv[l].r holds the right-hand side and cur[l] = omega Dinv r
  cycle = [&](int l) -> int {
    if (l == nl - 1) { PGO_TRY(coarsest()); amg_mark(g, M, nl - 1); return PGO_OK; }
    PGO_TRY(down(l));
    amg_mark(g, M, l);
    PGO_TRY(cycle(l + 1));
    if (M->gamma >= 2 && l + 1 < nl - 1 && l + 1 <= M->gamma_depth) PGO_TRY(revisit(l + 1));
    PGO_TRY(up(l));
    amg_mark(g, M, nl + l);
    return PGO_OK;
  };
  // Levels [Lc, nl) are the replicated tail of a multi-GPU hierarchy (one GPU: just the coarsest): no communication in
  // there, the same launches with the same pointers in every cycle -- and each of them a few microseconds long, i.e.
  // launch bound.  Multi-GPU runs whose exchanges go through NCCL replay that tail as a CUDA graph (NCCL stays outside
  // of captures; with peer-memory exchanges and on one GPU the whole PCG iteration is one graph anyway).
  int Lr = nl;
  for (int l = nl - 1; l >= 0; --l) if (g->world > 1 && M->lv[l].replicated) Lr = l;
  const int Lc = std::min(Lr, nl - 1);
  static const bool tail_graph_off = getenv("PGO_AMG_TAIL_GRAPH") && atoi(getenv("PGO_AMG_TAIL_GRAPH")) == 0;
  const bool want_tail_graph = g->world > 1 && Lc < nl - 1 && Lc > 0 && !tail_graph_off && !M->whole_iteration_graph && M->gamma < 2;
  if (!want_tail_graph) {
    PGO_TRY(cycle(0));
    *out = cur[0];
    return PGO_OK;
  }
  for (int l = 0; l < Lc; ++l) { PGO_TRY(down(l)); amg_mark(g, M, l); }
  auto tail = [&]() -> int { return cycle(Lc); };
  if (M->tail_graph) {
    CUDA_TRY(cudaGraphLaunch(M->tail_graph, g->stream));
    g->launches += M->tail_graph_kernels;
    for (int l = Lc; l < nl; ++l) cur[l] = M->tail_cur[l];
  } else if (M->tail_calls >= 1 && !M->tail_graph_failed) {
    // (the first cycle ran as plain launches: every lazy initialisation is behind us)
    const std::vector<double*> cur0 = cur, oth0 = oth;
    const long long l0 = g->launches;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(g->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = tail();
    const cudaError_t ce = cudaStreamEndCapture(g->stream, &graph);
    M->tail_graph_kernels = (int)(g->launches - l0);
    g->launches = l0;
    bool ok = rc == PGO_OK && ce == cudaSuccess && graph != nullptr;
    if (ok && cudaGraphInstantiate(&M->tail_graph, graph, 0) != cudaSuccess) { M->tail_graph = nullptr; ok = false; }
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      M->tail_graph_failed = true;
      cur = cur0; oth = oth0;
      PGO_TRY(tail());
    } else {
      M->tail_cur = cur;                       // where the captured tail leaves its results
      CUDA_TRY(cudaGraphLaunch(M->tail_graph, g->stream));
      g->launches += M->tail_graph_kernels;
    }
  } else {
    PGO_TRY(tail());
    M->tail_calls++;
  }
  for (int l = Lc - 1; l >= 0; --l) { PGO_TRY(up(l)); amg_mark(g, M, nl + l); }
  *out = cur[0];
  return PGO_OK;
}

// One PCG iteration, enqueued on g->stream: V-cycle, halo of u, SpMV + partial dots, (all-)reduce, scalar step, update.
static int amg_enqueue_iteration(pgo_graph* g, pgo::Amg* M, const pgo_solver_options* o, int sp_ctas, int rows_grid) {
  using namespace pgo;
  AmgLevelDev& L0 = M->lv[0];
  PcgMultiState* st = M->state;
  const int n = L0.n_own;
  double* u = nullptr;
  const int nl = M->num_levels;
  PGO_TRY(amg_vcycle(g, M, &u));
  PGO_TRY(amg_exchange(g, M, L0, u, 6, 6, &st->done));
  amg_mark(g, M, 2 * nl);
  if (M->warp_spmv) {
    amg_spmv_dots_warp_kernel<<<sp_ctas, kAmgThreads, 0, g->stream>>>(amg_view(L0), g->dlm, u, g->vr, g->vw, M->part, &st->done);
  } else {
    spmv_kernel<true><<<sp_ctas, 256, 0, g->stream>>>(amg_view(L0), u, g->dlm, g->vw, true, M->part, &st->done);
    amg_dot_kernel<<<sp_ctas, kAmgThreads, 0, g->stream>>>(6 * n, g->vr, u, M->part + sp_ctas, &st->done);
    g->launches++;
  }
  amg_mark(g, M, 2 * nl + 1);
  if (g->world > 1) {
    amg_reduce_kernel<<<1, kAmgThreads, 0, g->stream>>>(M->part, sp_ctas, 2, M->red);
    if (M->peer) {
      peer_push_kernel<<<dim3(1, M->rpush.n_targets), 32, 0, g->stream>>>(M->rpush, nullptr, M->red, &st->done);
      amg_pcg_scalar_peer_kernel<<<1, 32, 0, g->stream>>>(M->rwait, g->world, g->rank, M->red, st, o->pcg_max_iterations, o->pcg_tolerance);
      g->launches += 1;
      M->peer->pushes++; M->peer->push_bytes += 48ll * (g->world - 1);
    } else {
      PGO_TRY(amg_allreduce(g, M->red, 2));
      amg_pcg_scalar_kernel<<<1, 32, 0, g->stream>>>(st, M->red, o->pcg_max_iterations, o->pcg_tolerance);
    }
    g->launches++;
  } else {
    amg_pcg_reduce_scalar_kernel<<<1, kAmgThreads, 0, g->stream>>>(st, M->part, sp_ctas, o->pcg_max_iterations, o->pcg_tolerance);
  }
  amg_mark(g, M, 2 * nl + 2);
  if (M->Minv_f) amg_pcg_update_kernel<float><<<rows_grid, kAmgThreads, 0, g->stream>>>(n, M->Minv_f, u, g->vw, M->omega, g->vx, g->vr, g->vp, g->vs, L0.x, st);
  else amg_pcg_update_kernel<double><<<rows_grid, kAmgThreads, 0, g->stream>>>(n, g->Minv, u, g->vw, M->omega, g->vx, g->vr, g->vp, g->vs, L0.x, st);
  amg_mark(g, M, 2 * nl + 3);
  g->launches += 3;
  return PGO_OK;
}

// (H + diag(dlm)) x = b by AMG-preconditioned CG.  x -> g->vx; statistics -> g->scalars.  Stream-ordered; the host polls
// the `done` flag one batch behind the GPU (kernels after convergence are no-ops), and every rank takes the same exit
// decision because the flag derives from all-reduced, bit-identical scalars.  One GPU: the iteration is captured once
// per (H, poses) buffer pair into a CUDA graph and replayed -- a small graph's iteration is two dozen dependent
// few-microsecond kernels, i.e. launch bound.
static int amg_pcg_solve(pgo_graph* g, const pgo_solver_options* o, const double* b) {
  using namespace pgo;
  Amg* M = g->amg;
  const int n = M->lv[0].n_own, n6 = 6 * n;
  PGO_TRY(amg_setup_numeric(g, M));
  PcgMultiState* st = M->state;
  const int rows_grid = amg_rows_grid(n);
  // small level 0: one warp per row (its grid is exact, 8 rows per CTA), else the grid-stride 6-lanes-per-row form
  M->warp_spmv = (n + 7) / 8 <= M->part_cap / 3;
  const int sp_ctas = M->warp_spmv ? std::max(1, (n + 7) / 8) : std::max(1, std::min(rows_grid, std::min(8 * g->num_sms, M->part_cap / 3)));
  const int dot_ctas = std::max(1, std::min((n6 + kAmgThreads - 1) / kAmgThreads, std::min(4 * g->num_sms, M->part_cap / 3)));
  AmgLevelDev& L0 = M->lv[0];
  CUDA_TRY(cudaMemsetAsync(st, 0, sizeof(PcgMultiState), g->stream));
  if (M->Minv_f) amg_pcg_init_kernel<float><<<rows_grid, kAmgThreads, 0, g->stream>>>(n, b, M->Minv_f, M->omega, g->vx, g->vr, g->vp, g->vs, L0.x);
  else amg_pcg_init_kernel<double><<<rows_grid, kAmgThreads, 0, g->stream>>>(n, b, g->Minv, M->omega, g->vx, g->vr, g->vp, g->vs, L0.x);
  g->launches++;
  if (M->prof) amg_mark_kernel<<<1, 1, 0, g->stream>>>(M->prof, -1, nullptr);
  static const int graph_env = getenv("PGO_AMG_GRAPH") ? atoi(getenv("PGO_AMG_GRAPH")) : -1;
  // (multi-GPU: only when every per-iteration exchange runs over peer memory -- NCCL calls stay out of captures)
  const bool use_graph = (graph_env >= 0 ? graph_env != 0 : true) && (g->world == 1 || M->peer != nullptr);
  M->whole_iteration_graph = use_graph;
  Amg::IterGraph* ig = nullptr;
  if (use_graph) {
    const void* key[4] = {g->Hdiag, g->Hoff, g->poses, g->scale};
    for (auto& e : M->graphs)
      if (!std::memcmp(e.key, key, sizeof key) && e.max_it == o->pcg_max_iterations && e.tol == o->pcg_tolerance) ig = &e;
    if (!ig) {
      if (M->graphs.size() >= 8) { for (auto& e : M->graphs) cudaGraphExecDestroy(e.exec); M->graphs.clear(); }
      const long long l0 = g->launches;
      const long long px0 = M->peer ? M->peer->pushes : 0, pb0 = M->peer ? M->peer->push_bytes : 0;
      cudaGraph_t graph = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(g->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = amg_enqueue_iteration(g, M, o, sp_ctas, rows_grid);
      const cudaError_t ce = cudaStreamEndCapture(g->stream, &graph);
      if (rc != PGO_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) return set_error(PGO_ERR_CUDA, "capturing the PCG iteration failed: %s", cudaGetErrorString(ce));
      Amg::IterGraph e;
      std::memcpy(e.key, key, sizeof key);
      e.max_it = o->pcg_max_iterations; e.tol = o->pcg_tolerance; e.kernels = (int)(g->launches - l0); e.exec = nullptr;
      e.pushes = M->peer ? M->peer->pushes - px0 : 0; e.push_bytes = M->peer ? M->peer->push_bytes - pb0 : 0;
      if (M->peer) { M->peer->pushes = px0; M->peer->push_bytes = pb0; }     // counted per replay below
      g->launches = l0;
      const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) return set_error(PGO_ERR_CUDA, "instantiating the PCG iteration graph failed: %s", cudaGetErrorString(ie));
      M->graphs.push_back(e);
      ig = &M->graphs.back();
    }
  }
  const int batch = n >= 100000 ? 4 : 8;
  int enqueued = 0;          // batches
  const int max_batches = (o->pcg_max_iterations + batch - 1) / batch + 2;
  bool finished = false;
  while (!finished) {
    for (int k = 0; k < batch; ++k) {
      if (ig) {
        CUDA_TRY(cudaGraphLaunch(ig->exec, g->stream));
        g->launches += ig->kernels;
        if (M->peer) { M->peer->pushes += ig->pushes; M->peer->push_bytes += ig->push_bytes; }
      }
      else PGO_TRY(amg_enqueue_iteration(g, M, o, sp_ctas, rows_grid));
    }
    const int slot = enqueued & 1;
    CUDA_TRY(cudaMemcpyAsync(M->state_h + slot, st, sizeof(PcgMultiState), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaEventRecord(M->ev[slot], g->stream));
    ++enqueued;
    if (enqueued >= 2) {
      const int prev = (enqueued - 2) & 1;
      CUDA_TRY(cudaEventSynchronize(M->ev[prev]));
      if (M->state_h[prev].pad) return set_error(PGO_ERR_NCCL, "multilevel PCG: a peer-memory exchange timed out (a rank of the solve is gone?)");
      if (M->state_h[prev].done) finished = true;
    }
    if (enqueued >= max_batches) finished = true;
  }
  if (M->prof && g->rank == 0 && getenv("PGO_AMG_PROFILE") && atoi(getenv("PGO_AMG_PROFILE")) > 1) {
    // cumulative stage times so far (all solves of this graph): printed after every solve at level 2
    unsigned long long h[64];
    CUDA_TRY(cudaMemcpyAsync(h, M->prof, sizeof h, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    const int nl = M->num_levels;
    fprintf(stderr, "[pgo amg profile] cumulative ms:");
    for (int l = 0; l + 1 < nl; ++l) fprintf(stderr, " down%d %.2f", l, 1e-6 * (double)h[l]);
    fprintf(stderr, " coarsest %.2f", 1e-6 * (double)h[nl - 1]);
    for (int l = nl - 2; l >= 0; --l) fprintf(stderr, " up%d %.2f", l, 1e-6 * (double)h[nl + l]);
    fprintf(stderr, " | exchange(u) %.2f spmv+dots %.2f reduce+scalar %.2f update %.2f | halo exchanges: compute before %.2f push %.2f wait %.2f\n",
            1e-6 * (double)h[2 * nl], 1e-6 * (double)h[2 * nl + 1], 1e-6 * (double)h[2 * nl + 2], 1e-6 * (double)h[2 * nl + 3],
            1e-6 * (double)h[40], 1e-6 * (double)h[41], 1e-6 * (double)h[42]);
  }
  // epilogue: w = A x for the model cost change
  PGO_TRY(amg_exchange(g, M, L0, g->vx, 6, 6, nullptr));
  spmv_kernel<false><<<sp_ctas, 256, 0, g->stream>>>(amg_view(L0), g->vx, g->dlm, g->vw, true);
  amg_final_kernel<<<dot_ctas, kAmgThreads, 0, g->stream>>>(n6, g->vx, b, g->vw, g->dlm, M->part);
  amg_reduce_kernel<<<1, kAmgThreads, 0, g->stream>>>(M->part, dot_ctas, 3, M->red);
  PGO_TRY(amg_allreduce(g, M->red, 3));
  amg_final_store_kernel<<<1, 32, 0, g->stream>>>(M->red, st, g->scalars);
  g->launches += 4;
  CUDA_TRY(cudaGetLastError());
  return PGO_OK;
}
