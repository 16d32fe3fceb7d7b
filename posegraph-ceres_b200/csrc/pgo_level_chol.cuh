// pgo_level_chol.cuh -- level-scheduled sparse block Cholesky used as the PCG preconditioner.
//
// Why: the reference solves the damped normal equations exactly (SPARSE_NORMAL_CHOLESKY,
// REF/test/pose_graph_ceres_plus_finial.cpp:536).  Odometry-chain pose graphs such as KITTI-00 are
// beam-like: at late LM radii block-Jacobi PCG needs 10^5 iterations to follow that exact path.
// Here M = L L^T is the exact factor of H + D on a parallel elimination order, so the first PCG
// iterate x1 = alpha M^-1 b already is the direct solution; further PCG iterations (whose SpMV is
// still bsr6_row) act as iterative refinement and only run when ||b - A x1|| > tol ||b||.
//
// Host (once per graph): rounds of independent-set minimum-degree elimination -> levels; nodes of
// a level are mutually non-adjacent, so their columns factor in parallel.  Structure of L, the
// L<-A gather map, per-node Schur update tasks are precomputed.
// Device (ONE persistent kernel per LM step):
//   S   LM diagonal D = clamp(diag H) / radius (LevenbergMarquardtStrategy::ComputeStep), gather
//       A + D into the factor storage, working rhs t = b
//   F   level by level: 6x6 Cholesky of the pivot block, y_v = L_vv^-1 t_v, column scaling, Schur
//       updates and the rhs fan-out t_u -= L_uv y_v (the forward substitution rides on the factor)
//   B   level by level, descending: x_v = L_vv^-T (y_v - sum_u L_uv^T x_u)   (gather, deterministic)
//   CG  q = A p, alpha, x/r/Ax update, ||r|| check -> done, else z = M^-1 r (F without factoring + B) ...
// Two launch shapes share the kernel body: small graphs run as ONE thread-block cluster (16 CTAs on
// 16 SMs of a GPC) whose level barrier is the hardware cluster barrier (~0.2 us); large graphs
// run as a cooperative grid with an atomic-counter barrier.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pgo_kernels.cuh"

namespace pgo {

constexpr int kCholThreads = 512;   // per CTA, both launch shapes
constexpr int kCholClusterMaxNodes = 60000;
constexpr int kCholBlockMaxNodes = 256;       // at most this many nodes: one CTA with block barriers (measured on Manhattan-100)
constexpr long long kCholClusterMaxTasks = 200000;
constexpr int kCholWideLevelNodes = 384;       // cluster shape: leading levels with at least this many nodes run grid-wide   // graphs up to this many variable poses use the cluster shape

struct __align__(16) CholTask { int p; int q; int target; int pad; };   // target >= 0: L slot ; < 0: diagonal of node (-target-1)

struct LevelChol {
  bool usable = false;
  int n_nodes = 0;            // active poses
  int N = 0;
  long long factor_blocks = 0;
  int num_levels = 0;
  int max_degree = 0;
  long long n_slots = 0, n_tasks = 0;
  // device
  int* level_ptr = nullptr;     // [L+1] into nodes[]
  int* level_split = nullptr;   // [L] level mode: 8 / 16 / 32 lanes per node (staged), 1 = high-degree two-phase level
  int4* nodes = nullptr;        // [n_nodes+1] by elimination position: {pose id, col0, col1, task0}
  int* col_row = nullptr;       // [n_slots] row pose id of each slot
  int* l2a = nullptr;           // [n_slots] BSR off-diagonal entry that seeds the slot, or -1 (pure fill)
  CholTask* tasks = nullptr;    // [n_tasks]
  double* Lblk = nullptr;       // [n_slots][36] row-major (rows: row pose, cols: column pose)
  double* Ldiag = nullptr;      // [N][36] W_vv, then inverse of its lower Cholesky factor
  double* vt = nullptr;         // [N][6] working rhs / y of the sweeps
  double* partials = nullptr;
  unsigned int* barrier = nullptr;
  std::vector<int> level_ptr_h, level_mode_h;   // host copies for the wide-level launches
  int max_ctas = 0;             // cooperative-grid shape
  int cluster_ctas = 0;         // cluster shape: CTAs per cluster (0 = unavailable)
  std::vector<std::pair<void*, size_t>> blocks;   // device memory borrowed from the per-device pool
};

static void level_chol_destroy(LevelChol* c, int device) {
  if (!c) return;
  for (auto& blk : c->blocks) pool_free(device, blk.first, blk.second);
  delete c;
}

template <typename Tp>
static int chol_alloc(LevelChol* C, int device, Tp** dst, size_t count) {
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(Tp);
  CUDA_TRY(pool_alloc(device, reinterpret_cast<void**>(dst), bytes));
  C->blocks.emplace_back(static_cast<void*>(*dst), bytes);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
struct CholParams {
  BsrView A;
  // LM diagonal: mode 0 = new diagonal from H, 1 = reuse `diagonal`, 2 = dlm given
  int lm_mode; double min_diag, max_diag, radius;
  const LmState* lm;              // device-resident LM loop: radius / reuse-diagonal / done come from here (else nullptr)
  double* diagonal; double* dlm;
  const unsigned char* active;
  const double* b;
  double *x, *r, *z, *q, *p, *ax;   // PCG vectors [N][6]
  int num_levels, n_nodes;
  const int *level_ptr, *level_split, *col_row, *l2a;
  const int4* nodes;
  const CholTask* tasks;
  double *Lblk, *Ldiag, *vt;
  long long n_slots;
  double* partials;          // [8 regions][4 values][G]
  unsigned int* barrier;     // [0] grid barrier counter, [1] factor-failure flag
  DeviceScalars* scalars;
  int max_iterations;
  double tolerance;
  double accept;                  // 2-norm relative residual at which the direct solve is accepted without refinement
  int first_level;                // levels < first_level (and the S phase) were done by chol_wide_kernel launches
  int setup_done;                 // the S phase ran as a chol_wide_kernel launch
  unsigned long long* timeline;   // debug (PGO_TIMELINE=1): %globaltimer marks of CTA 0 / thread 0, [0] = count
};

__device__ __forceinline__ void chol_mark(const CholParams& P, int tag) {
  if (P.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long k = P.timeline[0];
    if (k < 1000) { P.timeline[1 + 2 * k] = t; P.timeline[2 + 2 * k] = (unsigned long long)tag; P.timeline[0] = k + 1; }
  }
}

#ifdef PGO_CHOL_FINE_MARKS   // debug build only: stage marks inside the staged factor step
#define CHOL_FINE(tag) chol_mark(P, tag)
#else
#define CHOL_FINE(tag) ((void)0)
#endif

// launch shapes of the solver kernel
enum CholShape { kShapeGrid = 0, kShapeCluster = 1, kShapeBlock = 3 };

template <int kShape>
__device__ __forceinline__ void chol_sync(unsigned int* counter, unsigned int& epoch) {
  if constexpr (kShape == kShapeCluster) {
    // hardware barrier over every thread of the cluster; release/acquire orders the global-memory traffic
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else if constexpr (kShape == kShapeBlock) {
    __syncthreads();   // one CTA: the block barrier orders its global-memory traffic (all mutable operands are read past L1)
  } else {
    grid_barrier(counter, epoch);
  }
}

// One warp, node at elimination position k.
//   kFactor: Cholesky of the 6x6 pivot block (every lane redundantly), store the inverse of the lower factor,
//            scale the column L_uv = W_uv Linv^T.
//   always : y_v = Linv t_v (t = working rhs), store y_v, fan out t_u -= L_uv y_v to the later rows.
template <bool kFactor>
__device__ __forceinline__ bool chol_forward_node(const CholParams& P, int v, int p0, int p1, int lane) {
  double Li[6][6];   // inverse of the lower factor (lower triangular)
  double* dv = P.Ldiag + 36 * (size_t)v;
  bool ok = true;
  if (kFactor) {
    double A[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) if (c <= r) A[r][c] = __ldcg(dv + r * 6 + c);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double s = A[j][j];
#pragma unroll
      for (int t = 0; t < 6; ++t) if (t < j) s -= A[j][t] * A[j][t];
      if (!(s > 0.0)) { ok = false; s = 1.0; }
      const double il = rsqrt(s);
      A[j][j] = il;            // the diagonal keeps 1 / L_jj
#pragma unroll
      for (int r = 0; r < 6; ++r) if (r > j) {
        double t2 = A[r][j];
#pragma unroll
        for (int t = 0; t < 6; ++t) if (t < j) t2 -= A[r][t] * A[j][t];
        A[r][j] = t2 * il;
      }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        if (r < c) { Li[r][c] = 0.0; continue; }
        double t = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int t3 = 0; t3 < 6; ++t3) if (t3 >= c && t3 < r) t -= A[r][t3] * Li[t3][c];
        Li[r][c] = t * A[r][r];
      }
    __syncwarp();
    if (lane < 6) {   // store Linv (row-major): lane = row
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double val = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) if (r == lane) val = Li[r][c];
        dv[lane * 6 + c] = val;
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) Li[r][c] = (c <= r) ? __ldcg(dv + r * 6 + c) : 0.0;
  }
  // y_v = Linv t_v
  double y[6];
  {
    double t[6];
    double* tv = P.vt + 6 * (size_t)v;
#pragma unroll
    for (int c = 0; c < 6; ++c) t[c] = __ldcg(tv + c);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) if (c <= r) s = fma(Li[r][c], t[c], s);
      y[r] = s;
    }
    __syncwarp();
    if (lane < 6) {
      double val = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) if (r == lane) val = y[r];
      tv[lane] = val;
    }
  }
  // column items: (block, row)
  const int items = (p1 - p0) * 6;
  for (int it = lane; it < items; it += 32) {
    const int blk = it / 6, r = it - blk * 6;
    double* w = P.Lblk + 36 * (size_t)(p0 + blk) + r * 6;
    double o[6];
    if (kFactor) {
      double wr[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) wr[c] = __ldcg(w + c);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < 6; ++t) if (t <= c) s = fma(wr[t], Li[c][t], s);
        o[c] = s;
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) w[c] = o[c];
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) o[c] = __ldcg(w + c);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) s = fma(o[c], y[c], s);
    atomicAdd(P.vt + 6 * (size_t)__ldg(P.col_row + p0 + blk) + r, -s);
  }
  return ok;
}

// Schur update item: row r of  target -= L_p * L_q^T
__device__ __forceinline__ void chol_update_item(const CholParams& P, const CholTask t, int r) {
  const double* lp = P.Lblk + 36 * (size_t)t.p + r * 6;
  const double* lq = P.Lblk + 36 * (size_t)t.q;
  double a[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) a[c] = __ldcg(lp + c);
  double* out = (t.target >= 0) ? (P.Lblk + 36 * (size_t)t.target + r * 6) : (P.Ldiag + 36 * (size_t)(-t.target - 1) + r * 6);
  const int cmax = (t.target >= 0) ? 5 : r;        // pivot blocks are symmetric: only their lower triangle is kept
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s = fma(a[k], __ldcg(lq + c * 6 + k), s);
    if (c <= cmax) atomicAdd(out + c, -s);
  }
}

// ---- shared-memory staging of one node's column (blocks, row ids, Schur tasks) per lane group ----
constexpr int kStashBlocks = 16;    // per warp: 4 nodes x 4, 2 x 8 or 1 x 16 blocks
constexpr int kStashTasks = 144;    // per warp: 4 x 36, 2 x 72 or 1 x 144 tasks (deg (deg + 1) / 2 <= 10 / 36 / 136)
struct __align__(16) WarpStash {
  double L[kStashBlocks][36];
  CholTask task[kStashTasks];
  int row[kStashBlocks];
};
constexpr int kCholSmemBytes = (kCholThreads / 32) * (int)sizeof(WarpStash);

__device__ __forceinline__ void cp_async_cg16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_ca4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Factor step for 32 / kLanes nodes per warp (positions kk .. kk + 32 / kLanes - 1 of one level, degree <= 128 / kLanes):
// everything a node needs is fetched in ONE latency epoch (cp.async into the stash + broadcast loads of the pivot
// block and rhs), the Cholesky, column scaling and Schur products then run out of registers / shared memory, and
// results leave as plain stores and fire-and-forget fp64 RED atomics.
template <int kLanes, bool kFactor = true>
__device__ __forceinline__ bool chol_factor_staged(const CholParams& P, WarpStash& st, int kk, int k1, int lane,
                                                   bool pre = false, int4 pre0 = int4(), int4 pre1 = int4()) {   // pre0/1: this lane's node records k, k+1 (prefetched)
  constexpr int kNpw = 32 / kLanes;
  constexpr int kBlk = kStashBlocks / kNpw;
  constexpr int kTsk = kStashTasks / kNpw;
  const int sub = lane % kLanes, gi = lane / kLanes;
  const int k = kk + gi;
  const bool valid = k < k1;
  int v = 0, p0 = 0, p1 = 0, t0 = 0, t1 = 0;
  if (valid) {
    const int4 nm = pre ? pre0 : __ldg(P.nodes + k);
    v = nm.x; p0 = nm.y; p1 = nm.z; t0 = nm.w;
    t1 = pre ? pre1.w : __ldg(&P.nodes[k + 1].w);
  }
  if (kFactor) CHOL_FINE(500);
  const int deg = p1 - p0, ntask = t1 - t0;
  double* SL = &st.L[gi * kBlk][0];
  CholTask* ST = st.task + gi * kTsk;
  int* SR = st.row + gi * kBlk;
  {
    // static structure first (it does not depend on other nodes), then the node's column; L1-bypassing copies
    // (cp.async.cg): the blocks were updated by other SMs' RED atomics at L2
    for (int j = sub; j < deg; j += kLanes) SR[j] = __ldg(P.col_row + p0 + j);
    if (kFactor) {
      for (int j = sub; j < ntask; j += kLanes) cp_async_cg16(ST + j, P.tasks + t0 + j);
    }
    const double* src = P.Lblk + 36 * (size_t)p0;
    for (int c = sub; c < deg * 18; c += kLanes) cp_async_cg16(SL + 2 * c, src + 2 * c);
  }
  if (kFactor) CHOL_FINE(501);
  double* dv = P.Ldiag + 36 * (size_t)v;
  double* tv = P.vt + 6 * (size_t)v;
  double A[6][6], t[6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) if (c <= r) A[r][c] = valid ? __ldcg(dv + r * 6 + c) : (r == c ? 1.0 : 0.0);
#pragma unroll
  for (int c = 0; c < 6; ++c) t[c] = valid ? __ldcg(tv + c) : 0.0;
  bool ok = true;
  double Li[6][6];
  if (kFactor) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double s = A[j][j];
#pragma unroll
      for (int q = 0; q < 6; ++q) if (q < j) s -= A[j][q] * A[j][q];
      if (!(s > 0.0)) { ok = false; s = 1.0; }
      const double il = rsqrt(s);
      A[j][j] = il;            // the diagonal keeps 1 / L_jj
#pragma unroll
      for (int r = 0; r < 6; ++r) if (r > j) {
        double t2 = A[r][j];
#pragma unroll
        for (int q = 0; q < 6; ++q) if (q < j) t2 -= A[r][q] * A[j][q];
        A[r][j] = t2 * il;
      }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        if (r < c) { Li[r][c] = 0.0; continue; }
        double x = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) if (q >= c && q < r) x -= A[r][q] * Li[q][c];
        Li[r][c] = x * A[r][r];
      }
  } else {
    // refinement sweep: the pivot block already holds the inverse of its lower factor
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) Li[r][c] = (c <= r) ? A[r][c] : 0.0;
  }
  double y[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) if (c <= r) s = fma(Li[r][c], t[c], s);
    y[r] = s;
  }
  if (valid && sub < 6) {   // store Linv row `sub` and y[sub]
    double yv = 0.0;
    if (kFactor) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double val = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) if (r == sub) val = Li[r][c];
        dv[sub * 6 + c] = val;
      }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) if (r == sub) yv = y[r];
    tv[sub] = yv;
  }
  if (kFactor) CHOL_FINE(502);
  cp_async_wait_all();
  __syncwarp();
  if (kFactor) CHOL_FINE(503);
  // column scaling L_uv = W_uv Linv^T in the stash (+ write-through to global), rhs fan-out t_u -= L_uv y_v
  for (int it = sub; it < deg * 6; it += kLanes) {
    const int blk = it / 6, r = it - blk * 6;
    double* w = SL + 36 * blk + 6 * r;
    double wr[6], o[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) wr[c] = w[c];
    if (kFactor) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) if (q <= c) s = fma(wr[q], Li[c][q], s);
        o[c] = s;
      }
      double* gw_ = P.Lblk + 36 * (size_t)(p0 + blk) + 6 * r;
#pragma unroll
      for (int c = 0; c < 6; ++c) { w[c] = o[c]; gw_[c] = o[c]; }
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) o[c] = wr[c];
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) s = fma(o[c], y[c], s);
    atomicAdd(P.vt + 6 * (size_t)SR[blk] + r, -s);
  }
  __syncwarp();
  if (!kFactor) return true;
  CHOL_FINE(504);
  // Schur updates: row r of target -= L_p L_q^T, operands from the stash
  for (int it = sub; it < ntask * 6; it += kLanes) {
    const int ti = it / 6, r = it - ti * 6;
    const CholTask tk = ST[ti];
    const double* lp = SL + 36 * (tk.p - p0) + 6 * r;
    const double* lq = SL + 36 * (tk.q - p0);
    double a[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) a[c] = lp[c];
    double* out = (tk.target >= 0) ? (P.Lblk + 36 * (size_t)tk.target + r * 6) : (P.Ldiag + 36 * (size_t)(-tk.target - 1) + r * 6);
    const int cmax = (tk.target >= 0) ? 5 : r;     // pivot blocks are symmetric: only their lower triangle is kept
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 6; ++q) s = fma(a[q], lq[c * 6 + q], s);
      if (c <= cmax) atomicAdd(out + c, -s);
    }
  }
  CHOL_FINE(505);
  __syncwarp();
  return ok;
}

// Backward step for 32 / kLanes nodes per warp: x_v = Linv_v^T (y_v - sum_{u in col(v)} L_uv^T x_u).
template <int kLanes>
__device__ __forceinline__ void chol_backward_staged(const CholParams& P, int kk, int k1, int lane, double* dst, double* dst2,
                                                     bool pre = false, int4 pre0 = int4()) {
  constexpr int kNpw = 32 / kLanes;
  constexpr int kSub = kLanes / 6;                         // 6-lane column groups per node: 1, 2, 5
  constexpr int kIter = (kStashBlocks / kNpw + kSub - 1) / kSub;   // blocks per column group: 4, 4, 4
  const int sub = lane % kLanes, gi = lane / kLanes, base = lane - sub;
  const int k = kk + gi;
  const bool valid = k < k1;
  int v = 0, p0 = 0, p1 = 0;
  if (valid) { const int4 nm = pre ? pre0 : __ldg(P.nodes + k); v = nm.x; p0 = nm.y; p1 = nm.z; }
  const int sg = sub / 6, c = sub - 6 * sg;
  const bool on = valid && sg < kSub;
  int rows[kIter];
#pragma unroll
  for (int j = 0; j < kIter; ++j) {
    const int p = p0 + sg + j * kSub;
    rows[j] = (on && p < p1) ? __ldg(P.col_row + p) : -1;
  }
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < kIter; ++j) {
    if (rows[j] >= 0) {
      const double* L = P.Lblk + 36 * (size_t)(p0 + sg + j * kSub) + c;
      const double* xu = dst + 6 * (size_t)rows[j];
#pragma unroll
      for (int rr = 0; rr < 6; ++rr) acc = fma(__ldcg(L + rr * 6), __ldcg(xu + rr), acc);
    }
  }
  const int c6 = sub % 6;
  double tot = 0.0;
#pragma unroll
  for (int g = 0; g < kSub; ++g) tot += __shfl_sync(0xffffffffu, acc, base + g * 6 + c6);
  const double sv = __ldcg(P.vt + 6 * (size_t)v + c6) - tot;
  double xv = 0.0;
#pragma unroll
  for (int rr = 0; rr < 6; ++rr) {
    const double sr = __shfl_sync(0xffffffffu, sv, base + rr);
    if (sub < 6 && rr >= sub) xv = fma(__ldcg(P.Ldiag + 36 * (size_t)v + rr * 6 + sub), sr, xv);
  }
  if (valid && sub < 6) { dst[6 * (size_t)v + sub] = xv; if (dst2) dst2[6 * (size_t)v + sub] = xv; }
}

__device__ __forceinline__ double cta_sum_n(double v, double* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  const int nwarp = blockDim.x >> 5;
  for (int k = 0; k < nwarp; ++k) t += red[k];
  return t;
}

// S phase: LM diagonal D = clamp(diag H) / radius, factor storage <- A + D, working rhs t = b, PCG vectors.
__device__ __forceinline__ void chol_setup_phase(const CholParams& P, int gtid, int gthreads) {
  const int n6 = 6 * P.A.n;
  const int lm_mode = P.lm ? (P.lm->reuse_diagonal ? 1 : 0) : P.lm_mode;
  const double radius = P.lm ? P.lm->radius : P.radius;
  for (int k = gtid; k < n6; k += gthreads) {
    const int i = k / 6, c = k - 6 * i;
    double dd;
    if (lm_mode == 2) {
      dd = P.dlm[k];
    } else {
      double dg;
      if (lm_mode == 1) dg = P.diagonal[k];
      else { dg = fmin(fmax(P.A.Hdiag[36 * (size_t)i + pidx(c, c)], P.min_diag), P.max_diag); P.diagonal[k] = dg; }
      dd = dg / radius;
      P.dlm[k] = dd;
    }
    P.vt[k] = P.b[k];
    P.x[k] = 0.0; P.ax[k] = 0.0; P.r[k] = P.b[k]; P.z[k] = 0.0; P.p[k] = 0.0;   // inactive poses keep z = p = 0
    // diagonal block row c of pose i
    const double* hd = P.A.Hdiag + 36 * (size_t)i;
    double* ld = P.Ldiag + 36 * (size_t)i + 6 * c;
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) ld[cc] = hd[pidx(c, cc)] + (cc == c ? dd : 0.0);
  }
  {
    // factor storage <- A: four (slot, row) items per thread and pass so that their loads are in flight together
    const double* __restrict__ hoff = P.A.Hoff;
    const int* __restrict__ l2a = P.l2a;
    double* __restrict__ lblk = P.Lblk;
    const long long total = P.n_slots * 6;
    for (long long e0 = gtid; e0 < total; e0 += 4LL * gthreads) {
      int src[4];
      double vals[4][6];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long e = e0 + (long long)j * gthreads;
        src[j] = (e < total) ? __ldg(l2a + e / 6) : -2;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long e = e0 + (long long)j * gthreads;
        const int rr = (int)(e % 6);
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) vals[j][cc] = (src[j] >= 0) ? __ldg(hoff + 36 * (size_t)src[j] + pidx(rr, cc)) : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long e = e0 + (long long)j * gthreads;
        if (src[j] != -2) {
          double* dst = lblk + 6 * e;
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) dst[cc] = vals[j][cc];
        }
      }
    }
  }
}

// Wide levels and the S phase as ordinary grid-wide launches: RED-atomic throughput is per SM (LSU bound), so the
// levels that hold most of the nodes run on all SMs; the kernel boundary is their barrier.  phase 0 = S,
// phase 1 = factor the nodes [k0, k1) of one level in `mode` (8 / 16 / 32 lanes per node).
__global__ void __launch_bounds__(kCholThreads, 1) chol_wide_kernel(const CholParams P, int phase, int mode, int k0, int k1) {
  extern __shared__ __align__(16) unsigned char chol_smem[];
  if (P.lm && P.lm->done) return;   // the LM loop has terminated: iterations enqueued ahead of the host's check are no-ops
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * (kCholThreads / 32) + warp, nw = gridDim.x * (kCholThreads / 32);
  if (phase == 0) { chol_setup_phase(P, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); return; }
  WarpStash& stash = reinterpret_cast<WarpStash*>(chol_smem)[warp];
  bool ok = true;
  if (mode == 8) { for (int kk = k0 + gw * 4; kk < k1; kk += nw * 4) ok &= chol_factor_staged<8>(P, stash, kk, k1, lane); }
  else if (mode == 16) { for (int kk = k0 + gw * 2; kk < k1; kk += nw * 2) ok &= chol_factor_staged<16>(P, stash, kk, k1, lane); }
  else { for (int kk = k0 + gw; kk < k1; kk += nw) ok &= chol_factor_staged<32>(P, stash, kk, k1, lane); }
  if (!ok) atomicExch(P.barrier + 1, 1u);
}

template <int kShape>
__global__ void __launch_bounds__(kCholThreads, 1) level_chol_pcg_kernel(const CholParams P) {
  __shared__ double red[kCholThreads / 32];
  __shared__ double bcast;
  extern __shared__ __align__(16) unsigned char chol_smem[];
  if (P.lm && P.lm->done) return;   // uniform over the whole launch (nobody writes `done` while a solver kernel runs)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpStash& stash = reinterpret_cast<WarpStash*>(chol_smem)[warp];
  const int warps_per_cta = kCholThreads / 32;
  const int gw = blockIdx.x * warps_per_cta + warp;
  const int nw = gridDim.x * warps_per_cta;
  // level work is dealt CTA-minor (node j -> CTA j mod G): a narrow level then spreads over all SMs instead of filling
  // the first CTAs -- the fp64 RED throughput that bounds a node's Schur updates is per SM
  const int gwx = warp * (int)gridDim.x + (int)blockIdx.x;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gthreads = gridDim.x * blockDim.x;
  const int grp = lane / 6, r6 = lane - grp * 6;
  const bool lane_on = grp < kRowsPerWarp;
  const int n = P.A.n;
  const int n6 = 6 * n;
  const int G = gridDim.x;
  unsigned int epoch = 0;
  int region = 0;
  auto next_region = [&]() -> double* { region = (region + 1) & 7; return P.partials + (size_t)region * 4 * G; };
  // level table in shared memory (one L2 round trip less per level) and node-record prefetch across the level barrier
  constexpr int kMaxLv = 1024;
  __shared__ int s_lptr[kMaxLv + 1];
  __shared__ int s_lmode[kMaxLv];
  const bool lv_smem = P.num_levels <= kMaxLv;
  if (lv_smem) {
    for (int l = threadIdx.x; l <= P.num_levels; l += blockDim.x) { s_lptr[l] = __ldg(P.level_ptr + l); if (l < P.num_levels) s_lmode[l] = __ldg(P.level_split + l); }
    __syncthreads();
  }
  auto lv_ptr = [&](int l) -> int { return lv_smem ? s_lptr[l] : __ldg(P.level_ptr + l); };
  auto lv_mode = [&](int l) -> int { return lv_smem ? s_lmode[l] : __ldg(P.level_split + l); };
  int4 pre0 = make_int4(0, 0, 0, 0), pre1 = make_int4(0, 0, 0, 0);
  bool have_pre = false;
  // fetch this lane's node records for the first group this warp will own in level l
  auto prefetch_level = [&](int l, bool bwd = false) {
    have_pre = false;
    if (l < 0 || l >= P.num_levels) return;
    const int raw = lv_mode(l);
    if (raw == 1) return;
    const int m = bwd ? (raw >> 8) : (raw & 0xff);
    const int k = lv_ptr(l) + gwx * (32 / m) + lane / m;
    if (k < lv_ptr(l + 1)) { pre0 = __ldg(P.nodes + k); pre1 = __ldg(P.nodes + k + 1); }
    have_pre = true;
  };

  // ---- S: LM diagonal, gather A + D into the factor storage, working rhs ----
  chol_mark(P, 0);
  if (!P.setup_done) chol_setup_phase(P, gtid, gthreads);
  chol_sync<kShape>(P.barrier, epoch);
  chol_mark(P, 1);

  // forward sweep over the levels; kFactor also factors.  level_split[l] is the level's mode: 8 / 16 / 32 = lanes
  // per node of the staged path (degree <= 4 / 8 / 16), 1 = high-degree level (global operands, tasks spread over
  // the whole launch after an extra barrier).
  auto forward = [&](auto factor_tag) {
    constexpr bool kFactor = decltype(factor_tag)::value;
    const int l0 = kFactor ? P.first_level : 0;
    prefetch_level(l0);
    for (int l = l0; l < P.num_levels; ++l) {
      const int k0 = lv_ptr(l), k1 = lv_ptr(l + 1);
      const int mode = lv_mode(l) & 0xff;
      const bool split = kFactor && mode == 1;
      if (mode != 1) {
        bool ok = true;
        bool pp = have_pre;   // only the first group of the level was prefetched
        if (mode == 8) { for (int kk = k0 + gwx * 4; kk < k1; kk += nw * 4) { ok &= chol_factor_staged<8, kFactor>(P, stash, kk, k1, lane, pp, pre0, pre1); pp = false; } }
        else if (mode == 16) { for (int kk = k0 + gwx * 2; kk < k1; kk += nw * 2) { ok &= chol_factor_staged<16, kFactor>(P, stash, kk, k1, lane, pp, pre0, pre1); pp = false; } }
        else { for (int kk = k0 + gwx; kk < k1; kk += nw) { ok &= chol_factor_staged<32, kFactor>(P, stash, kk, k1, lane, pp, pre0, pre1); pp = false; } }
        if (!ok) atomicExch(P.barrier + 1, 1u);
      } else {
        for (int k = k0 + gwx; k < k1; k += nw) {
          const int4 nm = __ldg(P.nodes + k);
          const bool ok = chol_forward_node<kFactor>(P, nm.x, nm.y, nm.z, lane);
          if (kFactor && !ok && lane == 0) atomicExch(P.barrier + 1, 1u);
        }
      }
      if (split) {
        chol_sync<kShape>(P.barrier, epoch);
        const long long t0 = __ldg(&P.nodes[k0].w), t1 = __ldg(&P.nodes[k1].w);
        for (long long it = (long long)gtid; it < (t1 - t0) * 6; it += gthreads) chol_update_item(P, P.tasks[t0 + it / 6], (int)(it % 6));
      }
      prefetch_level(l + 1);       // static structure: its latency hides behind the barrier
      chol_mark(P, 100 + l);
      chol_sync<kShape>(P.barrier, epoch);
      chol_mark(P, 200 + l);
    }
  };
  // backward: x_v = Linv_v^T (y_v - sum_{u later} L_uv^T x_u), levels descending; writes dst (and dst2)
  auto backward = [&](double* dst, double* dst2) {
    prefetch_level(P.num_levels - 1, true);
    for (int l = P.num_levels - 1; l >= 0; --l) {
      const int k0 = lv_ptr(l), k1 = lv_ptr(l + 1);
      const int raw_mode = lv_mode(l);
      const int mode = raw_mode == 1 ? 1 : (raw_mode >> 8);
      bool pp = have_pre;
      if (mode == 8) { for (int kk = k0 + gwx * 4; kk < k1; kk += nw * 4) { chol_backward_staged<8>(P, kk, k1, lane, dst, dst2, pp, pre0); pp = false; } }
      else if (mode == 16) { for (int kk = k0 + gwx * 2; kk < k1; kk += nw * 2) { chol_backward_staged<16>(P, kk, k1, lane, dst, dst2, pp, pre0); pp = false; } }
      else if (mode == 32) { for (int kk = k0 + gwx; kk < k1; kk += nw) { chol_backward_staged<32>(P, kk, k1, lane, dst, dst2, pp, pre0); pp = false; } }
      else {
        for (int k = k0 + gwx; k < k1; k += nw) {
          const int4 nm = __ldg(P.nodes + k);
          const int v = nm.x, p0 = nm.y, p1 = nm.z;
          double acc = 0.0;   // lane (grp, c = r6): sum_u sum_r L_uv[r][c] x_u[r]
          if (lane_on) {
            for (int p = p0 + grp; p < p1; p += kRowsPerWarp) {
              const double* L = P.Lblk + 36 * (size_t)p + r6;
              const double* xu = dst + 6 * (size_t)__ldg(P.col_row + p);
#pragma unroll
              for (int rr = 0; rr < 6; ++rr) acc = fma(__ldcg(L + rr * 6), __ldcg(xu + rr), acc);
            }
          }
          double tot = 0.0;
#pragma unroll
          for (int gq = 0; gq < kRowsPerWarp; ++gq) tot += __shfl_sync(0xffffffffu, acc, gq * 6 + (lane % 6));
          const double sv = __ldcg(P.vt + 6 * (size_t)v + (lane % 6)) - tot;
          double xv = 0.0;
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) {
            const double sr = __shfl_sync(0xffffffffu, sv, rr);
            if (lane < 6 && rr >= lane) xv = fma(__ldcg(P.Ldiag + 36 * (size_t)v + rr * 6 + lane), sr, xv);
          }
          if (lane < 6) { dst[6 * (size_t)v + lane] = xv; if (dst2) dst2[6 * (size_t)v + lane] = xv; }
        }
      }
      prefetch_level(l - 1, true);
      chol_mark(P, 300 + l);
      chol_sync<kShape>(P.barrier, epoch);
      chol_mark(P, 400 + l);
    }
  };

  // ---- F + B: z = p = M^-1 b ----
  forward(std::true_type{});
  backward(P.z, P.p);

  // ---- CG ----
  double rho = 0.0, rho0 = 0.0, bb = 0.0, rr = 0.0;
  double xtb = 0.0, xtAx = 0.0, xtDx = 0.0;
  int iter = 0, flag = 0, norm_kind = 0;
  for (;;) {
    // q = A p ; pq = p.q ; (first pass also rho = r.z and bb = b.b)
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int base = gw * kRowsPerWarp; base < n; base += nw * kRowsPerWarp) {
      const int i = base + grp;
      if (lane_on && i < n) {
        const double qv = bsr6_row<true>(P.A.Hdiag, P.A.Hoff, P.A.row_ptr, P.A.col_idx, P.p, P.dlm, i, r6);
        const size_t k = 6 * (size_t)i + r6;
        P.q[k] = qv;
        a0 = fma(qv, __ldcg(P.p + k), a0);
        if (iter == 0) { const double bv = P.b[k]; a1 = fma(bv, __ldcg(P.z + k), a1); a2 = fma(bv, bv, a2); }
      }
    }
    a0 = cta_sum_n(a0, red);
    if (iter == 0) { a1 = cta_sum_n(a1, red); a2 = cta_sum_n(a2, red); }
    double* part = next_region();
    if (threadIdx.x == 0) { part[blockIdx.x] = a0; if (iter == 0) { part[G + blockIdx.x] = a1; part[2 * G + blockIdx.x] = a2; } }
    chol_sync<kShape>(P.barrier, epoch);
    chol_mark(P, 2);
    const double pq = reduce_partials(part, G, &bcast);
    if (iter == 0) {
      rho0 = rho = reduce_partials(part + G, G, &bcast);
      bb = reduce_partials(part + 2 * G, G, &bcast);
      if (!(rho0 > 0.0) || !isfinite(rho0)) { flag = (rho0 == 0.0) ? 0 : 2; break; }
    }
    if (!(pq > 0.0) || !isfinite(pq)) { flag = 2; break; }
    const double alpha = rho / pq;
    ++iter;
    // x += alpha p ; r -= alpha q ; Ax += alpha q ; rr = r.r ; x.b ; x.Ax ; x.D x
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int k = gtid; k < n6; k += gthreads) {
      const double qv = __ldcg(P.q + k);
      const double xv = P.x[k] + alpha * __ldcg(P.p + k);
      const double rv = P.r[k] - alpha * qv;
      const double av = P.ax[k] + alpha * qv;
      P.x[k] = xv; P.r[k] = rv; P.ax[k] = av;
      s0 = fma(rv, rv, s0); s1 = fma(xv, P.b[k], s1); s2 = fma(xv, av, s2); s3 = fma(xv * xv, P.dlm[k], s3);
    }
    s0 = cta_sum_n(s0, red); s1 = cta_sum_n(s1, red); s2 = cta_sum_n(s2, red); s3 = cta_sum_n(s3, red);
    part = next_region();
    if (threadIdx.x == 0) { part[blockIdx.x] = s0; part[G + blockIdx.x] = s1; part[2 * G + blockIdx.x] = s2; part[3 * G + blockIdx.x] = s3; }
    chol_sync<kShape>(P.barrier, epoch);
    chol_mark(P, 3);
    rr = reduce_partials(part, G, &bcast);
    xtb = reduce_partials(part + G, G, &bcast);
    xtAx = reduce_partials(part + 2 * G, G, &bcast);
    xtDx = reduce_partials(part + 3 * G, G, &bcast);
    if (rr <= P.accept * P.accept * bb) { norm_kind = 1; break; }               // ||b - A x|| <= accept ||b||
    // z = M^-1 r
    for (int k = gtid; k < n6; k += gthreads) { P.vt[k] = P.r[k]; }
    chol_sync<kShape>(P.barrier, epoch);
    forward(std::false_type{});
    backward(P.z, nullptr);
    double t0 = 0.0;
    for (int k = gtid; k < n6; k += gthreads) t0 = fma(P.r[k], __ldcg(P.z + k), t0);
    t0 = cta_sum_n(t0, red);
    part = next_region();
    if (threadIdx.x == 0) part[blockIdx.x] = t0;
    chol_sync<kShape>(P.barrier, epoch);
    const double rho_new = reduce_partials(part, G, &bcast);
    const double beta = rho_new / rho;
    rho = rho_new;
    if (fabs(rho) <= P.tolerance * P.tolerance * rho0) break;                      // sqrt(r.M^-1 r) <= tol sqrt(b.M^-1 b)
    if (iter >= P.max_iterations) { flag = 1; break; }
    for (int k = gtid; k < n6; k += gthreads) P.p[k] = __ldcg(P.z + k) + beta * P.p[k];
    chol_sync<kShape>(P.barrier, epoch);
  }
  chol_mark(P, 4);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    P.scalars->xtb = xtb; P.scalars->xtAx = xtAx; P.scalars->xtDx = xtDx;
    if (norm_kind == 1) { P.scalars->pcg_gamma0 = bb; P.scalars->pcg_gamma = rr; }
    else { P.scalars->pcg_gamma0 = rho0; P.scalars->pcg_gamma = fabs(rho); }
    unsigned int failed;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(failed) : "l"(P.barrier + 1) : "memory");
    P.scalars->pcg_iterations = iter; P.scalars->pcg_flag = failed ? 3 : flag;
  }
}

// ---------------------------------------------------------------------------------------------
// host side: symbolic analysis
// ---------------------------------------------------------------------------------------------
template <typename Tp>
static int chol_upload(LevelChol* C, int device, Tp** dst, const std::vector<Tp>& src, cudaStream_t stream) {
  PGO_TRY(chol_alloc(C, device, dst, src.size()));
  if (!src.empty()) CUDA_TRY(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(Tp), cudaMemcpyHostToDevice, stream));
  return 0;
}

// Pure host part: elimination levels, structure of L, gather map, Schur tasks.
struct LevelCholSymbolic {
  std::string error;            // message of a failed analysis (the analysis may run on a helper thread; pgo_last_error() is per thread)
  bool usable = false;
  int n_nodes = 0, num_levels = 0, max_degree = 0;
  long long n_slots = 0;
  std::vector<int> level_ptr, level_split, col_row, l2a;
  std::vector<int4> nodes;
  std::vector<CholTask> tasks;
};

// max_levels / max_node_degree: AUTO only accepts a "cheap" factor (few levels, every level a low-degree staged level);
// the analysis gives up as soon as a limit is exceeded.
static int level_chol_symbolic(LevelCholSymbolic* S, int N, const unsigned char* active, const int* a_row_ptr,
                               const int* a_col_idx, double max_fill_ratio, int max_levels = 8192, int max_node_degree = 1 << 30) {
  const bool prof = getenv("PGO_PROFILE_HOST") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tp = now();
  auto lap = [&](const char* what) { if (prof) { const double t = now(); fprintf(stderr, "[pgo symbolic] %-22s %8.1f us\n", what, 1e6 * (t - tp)); tp = t; } };
  // adjacency lists live in one arena (offset / size / capacity per pose): no per-node allocations, and the column
  // structure of an eliminated node is simply its frozen list
  struct AdjList { long long off; int n, cap; };
  std::vector<AdjList> adj(N, AdjList{0, 0, 0});
  std::vector<int> pool;
  int n_nodes = 0;
  long long a_off = 0;
  {
    long long need = 0;
    for (int i = 0; i < N; ++i) if (active[i]) need += std::max(8, 2 * (a_row_ptr[i + 1] - a_row_ptr[i]));
    pool.reserve((size_t)(need + need / 2 + 1024));
    for (int i = 0; i < N; ++i) {
      if (!active[i]) continue;
      ++n_nodes;
      const int deg = a_row_ptr[i + 1] - a_row_ptr[i];
      adj[i] = AdjList{(long long)pool.size(), deg, std::max(8, 2 * deg)};
      pool.resize(pool.size() + (size_t)adj[i].cap);
      std::copy(a_col_idx + a_row_ptr[i], a_col_idx + a_row_ptr[i + 1], pool.begin() + adj[i].off);   // sorted, symmetric, active only
      a_off += deg;
    }
  }
  S->n_nodes = n_nodes;
  if (n_nodes == 0) return 0;
  const long long fill_cap = (long long)std::min(max_fill_ratio * (double)(a_off / 2 + n_nodes) + 64.0, 4.0e9);

  std::vector<int> pos(N, -1), order;
  order.reserve(n_nodes);
  std::vector<int>& level_ptr = S->level_ptr;
  std::vector<int>& level_split = S->level_split;
  level_ptr.assign(1, 0);
  // structure of L's column v (row pose ids, sorted by id) = adj[v] at the moment v is eliminated
  std::vector<unsigned char> alive(N, 0), blocked(N, 0);
  std::vector<int> alive_list;
  for (int i = 0; i < N; ++i) if (active[i]) { alive[i] = 1; alive_list.push_back(i); }
  std::vector<unsigned long long> cand;   // (degree << 32 | pose)
  std::vector<int> tmp, sel;
  std::vector<std::vector<int>> buckets;
  alive_list.reserve(n_nodes); cand.reserve(n_nodes); sel.reserve(n_nodes); tmp.reserve(64);
  long long slots = 0, work = 0;
  const long long work_cap = 100LL * (a_off + n_nodes) + 32000000LL;   // symbolic effort bound (merged adjacency entries)
  while (!alive_list.empty()) {
    int dmin = 1 << 30;
    long long dsum = 0;
    for (int v : alive_list) { dmin = std::min(dmin, adj[v].n); dsum += (long long)adj[v].n; }
    // mesh-like graphs (2-D grids, dense random loops) fill in quickly: give up as soon as the remaining
    // graph is both large and dense instead of grinding through a factor that would not pay off
    if (prof && getenv("PGO_PROFILE_LEVELS")) fprintf(stderr, "[pgo round] %3zu alive %8zu mean degree %.2f min %d\n", level_ptr.size() - 1, alive_list.size(), (double)dsum / (double)alive_list.size(), dmin);
    if (max_fill_ratio < 1e29 && alive_list.size() > 20000 && dsum > 16LL * (long long)alive_list.size()) return 0;
    // "cheap factor only" callers (AUTO: a node-degree limit is set): a chain-like graph keeps its mean degree near 2-3
    // while thousands of nodes are alive (KITTI-00: <= 3.1 above 1 000 alive nodes); a mesh is past 8 after one or two
    // rounds and will break the degree limit later anyway -- stop now instead of after seconds of merging
    if (max_node_degree < (1 << 30) && alive_list.size() > 1024 && dsum > (long long)(max_node_degree / 2) * (long long)alive_list.size()) return 0;
    const int thr = 2 * dmin + 2;
    // greedy independent set in order of increasing degree (then pose id): degree buckets instead of a sort
    sel.clear();
    cand.clear();
    if ((int)buckets.size() < thr - dmin + 1) buckets.resize((size_t)(thr - dmin + 1));
    for (int d = dmin; d <= thr; ++d) buckets[d - dmin].clear();
    for (int v : alive_list) if (adj[v].n <= thr) buckets[adj[v].n - dmin].push_back(v);
    for (int d = dmin; d <= thr; ++d) {
      for (int v : buckets[d - dmin]) {
        cand.push_back((unsigned long long)(unsigned)v);
        if (blocked[v]) continue;
        sel.push_back(v);
        blocked[v] = 1;
        const int* av = pool.data() + adj[v].off;
        for (int j = 0; j < adj[v].n; ++j) blocked[av[j]] = 1;
      }
    }
    int lvl_maxdeg = 0;
    for (int v : sel) {
      pos[v] = (int)order.size();
      order.push_back(v);
      slots += (long long)adj[v].n;
      lvl_maxdeg = std::max(lvl_maxdeg, adj[v].n);
    }
    if (slots > fill_cap || (int)level_ptr.size() > max_levels || lvl_maxdeg > max_node_degree || work > work_cap) return 0;   // not usable
    for (int v : sel) {
      const int nbn = adj[v].n;
      for (int jn = 0; jn < nbn; ++jn) {
        // adj[u] = (adj[u] U nb) \ {u, v}
        const int* nb = pool.data() + adj[v].off;      // re-read: the arena may have grown
        const int u = nb[jn];
        AdjList& au_l = adj[u];
        const int* au = pool.data() + au_l.off;
        const int na = au_l.n;
        work += (long long)(na + nbn);
        tmp.clear();
        int ia = 0, ib = 0;
        while (ia < na || ib < nbn) {
          int x;
          if (ib >= nbn || (ia < na && au[ia] < nb[ib])) x = au[ia++];
          else if (ia >= na || nb[ib] < au[ia]) x = nb[ib++];
          else { x = au[ia]; ++ia; ++ib; }
          if (x != u && x != v) tmp.push_back(x);
        }
        if ((int)tmp.size() > au_l.cap) {               // move the list to the end of the arena
          au_l.cap = 2 * (int)tmp.size();
          au_l.off = (long long)pool.size();
          pool.resize(pool.size() + (size_t)au_l.cap);
        }
        au_l.n = (int)tmp.size();
        std::copy(tmp.begin(), tmp.end(), pool.begin() + au_l.off);
      }
      alive[v] = 0;
    }
    for (unsigned long long dv : cand) blocked[(int)(dv & 0xffffffffu)] = 0;
    for (int v : sel) { const int* cv = pool.data() + adj[v].off; for (int j = 0; j < adj[v].n; ++j) blocked[cv[j]] = 0; }
    size_t w = 0;
    for (size_t k = 0; k < alive_list.size(); ++k) if (alive[alive_list[k]]) alive_list[w++] = alive_list[k];
    alive_list.resize(w);
    level_ptr.push_back((int)order.size());
    static const bool force32 = getenv("PGO_CHOL_MODE32") != nullptr;   // debug
    static const bool no_widen = getenv("PGO_CHOL_NO_WIDEN") != nullptr;   // debug: lanes per node from the degree only
    // lanes per node: what the degree needs (stash capacity), widened on narrow levels -- a level with fewer nodes than
    // the launch has warps is bound by one warp's walk through the node's Schur products, so give the node all 32 lanes
    // warps of the launch that will run this level: one CTA (block shape), the 16-CTA cluster, or one CTA per SM
    // (grid shape, and the wide levels of a cluster-shape factorisation that go out as chol_wide_kernel launches)
    const int lvl_nodes = (int)sel.size();
    const int launch_ctas = n_nodes <= kCholBlockMaxNodes ? 1 : (n_nodes > kCholClusterMaxNodes || lvl_nodes >= kCholWideLevelNodes) ? 148 : 16;
    const int launch_warps = launch_ctas * (kCholThreads / 32);
    int lanes = lvl_maxdeg <= 4 ? 8 : lvl_maxdeg <= 8 ? 16 : 32;
    if (!no_widen) lanes = std::max(lanes, lvl_nodes <= launch_warps ? 32 : lvl_nodes <= 2 * launch_warps ? 16 : 8);
    // the backward sweep of every level runs inside the solver kernel (no wide launches): its lanes are widened against
    // that launch only.  Packed: bits 0-7 forward lanes, bits 8-15 backward lanes; 1 = high-degree level.
    static const bool no_bwd_split = getenv("PGO_CHOL_NO_BWD_SPLIT") != nullptr;   // debug
    const int bwd_warps = (n_nodes <= kCholBlockMaxNodes ? 1 : n_nodes > kCholClusterMaxNodes ? 148 : 16) * (kCholThreads / 32);
    int bwd_lanes = lvl_maxdeg <= 4 ? 8 : lvl_maxdeg <= 8 ? 16 : 32;
    if (!no_widen) bwd_lanes = std::max(bwd_lanes, lvl_nodes <= bwd_warps ? 32 : lvl_nodes <= 2 * bwd_warps ? 16 : 8);
    if (force32) lanes = bwd_lanes = 32;
    if (no_bwd_split) bwd_lanes = lanes;
    level_split.push_back(lvl_maxdeg > 16 ? 1 : (lanes | (bwd_lanes << 8)));
    S->max_degree = std::max(S->max_degree, lvl_maxdeg);
  }
  S->num_levels = (int)level_split.size();
  S->n_slots = slots;
  lap("elimination rounds");

  // column-major slots by elimination position
  std::vector<int> col_ptr(n_nodes + 1, 0);
  std::vector<int>& col_row = S->col_row;
  col_row.resize((size_t)slots);
  for (int k = 0; k < n_nodes; ++k) col_ptr[k + 1] = col_ptr[k] + adj[order[k]].n;
  for (int k = 0; k < n_nodes; ++k) std::copy(pool.begin() + adj[order[k]].off, pool.begin() + adj[order[k]].off + adj[order[k]].n, col_row.begin() + col_ptr[k]);
  auto slot_of = [&](int row, int col) -> int {   // block (row, col), col eliminated first
    const int k = pos[col];
    const int* b = col_row.data() + col_ptr[k];
    const int* e = col_row.data() + col_ptr[k + 1];
    const int* it = std::lower_bound(b, e, row);
    return (it != e && *it == row) ? (int)(it - col_row.data()) : -1;
  };
  lap("column structure");
  // Schur update tasks per node
  S->nodes.resize((size_t)n_nodes + 1);
  std::vector<CholTask>& tasks = S->tasks;
  {
    size_t est = 0;
    for (int k = 0; k < n_nodes; ++k) { const size_t d = (size_t)adj[order[k]].n; est += d * (d + 1) / 2; }
    tasks.reserve(est);
  }
  for (int k = 0; k < n_nodes; ++k) {
    const int p0 = col_ptr[k], p1 = col_ptr[k + 1];
    S->nodes[k] = make_int4(order[k], p0, p1, (int)tasks.size());
    for (int p = p0; p < p1; ++p) {
      tasks.push_back({p, p, -col_row[p] - 1, 0});                  // diagonal of row(p)
      for (int q = p0; q < p1; ++q) {
        if (q == p) continue;
        const int u = col_row[p], w = col_row[q];
        if (pos[u] > pos[w]) {                                   // block (u, w): rows u, cols w = L_p L_q^T
          const int t = slot_of(u, w);
          if (t < 0) return set_error(PGO_ERR_NUMERICAL, "level Cholesky: missing fill slot");
          tasks.push_back({p, q, t, 0});
        }
      }
    }
    if (tasks.size() > 2000000000ull) return 0;
  }
  S->nodes[n_nodes] = make_int4(-1, (int)slots, (int)slots, (int)tasks.size());
  lap("tasks");
  // L slot <- A (BSR off-diagonal) entry
  S->l2a.assign((size_t)slots, -1);
  for (int i = 0; i < N; ++i)
    for (int p = a_row_ptr[i]; p < a_row_ptr[i + 1]; ++p) {
      const int j = a_col_idx[p];
      if (active[i] && active[j] && pos[j] < pos[i]) {
        const int s = slot_of(i, j);
        if (s < 0) return set_error(PGO_ERR_NUMERICAL, "level Cholesky: missing slot for a Hessian block");
        S->l2a[s] = p;
      }
    }
  lap("l2a");
  if (prof && getenv("PGO_PROFILE_LEVELS")) {
    for (int l = 0; l < S->num_levels; ++l) {
      int mx = 0; long long sum = 0;
      for (int k = level_ptr[l]; k < level_ptr[l + 1]; ++k) { const int nt = S->nodes[k + 1].w - S->nodes[k].w; mx = std::max(mx, nt); sum += nt; }
      fprintf(stderr, "[pgo level] %3d nodes %5d lanes fwd %2d bwd %2d tasks max %4d mean %.1f\n", l, level_ptr[l + 1] - level_ptr[l],
              level_split[l] & 0xff, level_split[l] >> 8, mx, (double)sum / std::max(1, level_ptr[l + 1] - level_ptr[l]));
    }
  }
  if (prof) {   // fingerprint of the whole symbolic result (regression aid for changes to this function)
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const void* ptr, size_t bytes) { const unsigned char* b = (const unsigned char*)ptr; for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    mix(S->level_ptr.data(), S->level_ptr.size() * sizeof(int)); mix(S->level_split.data(), S->level_split.size() * sizeof(int));
    mix(S->col_row.data(), S->col_row.size() * sizeof(int)); mix(S->l2a.data(), S->l2a.size() * sizeof(int));
    mix(S->nodes.data(), S->nodes.size() * sizeof(int4)); mix(S->tasks.data(), S->tasks.size() * sizeof(CholTask));
    fprintf(stderr, "[pgo symbolic] fingerprint %016llx (%d levels, %lld slots, %zu tasks)\n", h, S->num_levels, S->n_slots, S->tasks.size());
  }
  S->usable = true;
  return 0;
}

// Device half of the analysis: upload the symbolic structure `S` (computed by level_chol_symbolic, possibly on a helper
// thread while the graph was being uploaded) and size the launch shapes.
static int level_chol_analyze(LevelChol** out, int device, int N, const LevelCholSymbolic& S, cudaStream_t stream) {
  LevelChol* C = new LevelChol();
  *out = C;
  C->N = N;
  C->n_nodes = S.n_nodes;
  if (!S.usable) return 0;
  C->num_levels = S.num_levels;
  C->max_degree = S.max_degree;
  C->n_slots = S.n_slots;
  C->factor_blocks = S.n_slots + S.n_nodes;
  C->n_tasks = (long long)S.tasks.size();
  C->level_ptr_h = S.level_ptr;
  C->level_mode_h = S.level_split;
  PGO_TRY(chol_upload(C, device, &C->level_ptr, S.level_ptr, stream));
  PGO_TRY(chol_upload(C, device, &C->level_split, S.level_split, stream));
  PGO_TRY(chol_upload(C, device, &C->nodes, S.nodes, stream));
  PGO_TRY(chol_upload(C, device, &C->col_row, S.col_row, stream));
  PGO_TRY(chol_upload(C, device, &C->l2a, S.l2a, stream));
  PGO_TRY(chol_upload(C, device, &C->tasks, S.tasks, stream));
  PGO_TRY(chol_alloc(C, device, &C->Lblk, (size_t)S.n_slots * 36));
  PGO_TRY(chol_alloc(C, device, &C->Ldiag, (size_t)N * 36));
  PGO_TRY(chol_alloc(C, device, &C->vt, (size_t)N * 6));
  PGO_TRY(chol_alloc(C, device, &C->barrier, 4));
  // function attributes and occupancy are facts about (kernel, DEVICE): cached per device (pgo_pool.cuh), not per process
  int per_sm = -1, cluster_max = -1;
  const int sms = pool_num_sms(device);
  if (!pool_cache_get(device, kCacheCholPerSm, &per_sm)) {
    CUDA_TRY(cudaFuncSetAttribute(level_chol_pcg_kernel<kShapeGrid>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(level_chol_pcg_kernel<kShapeCluster>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(level_chol_pcg_kernel<kShapeBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(chol_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholSmemBytes));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, level_chol_pcg_kernel<kShapeGrid>, kCholThreads, kCholSmemBytes));
    pool_cache_set(device, kCacheCholPerSm, per_sm);
  }
  C->max_ctas = std::max(1, std::min(per_sm, 1) * sms);
  // cluster shape: the largest cluster (16, then 8) the device can co-schedule
  C->cluster_ctas = 0;
  // the cluster shape only pays when the whole factorisation is tiny (latency bound); otherwise all SMs are needed
  const bool want_cluster = S.n_nodes <= kCholClusterMaxNodes && (long long)S.tasks.size() <= kCholClusterMaxTasks;
  if (want_cluster && !pool_cache_get(device, kCacheCholCluster, &cluster_max)) {
    cluster_max = 0;
    cudaFuncSetAttribute(level_chol_pcg_kernel<kShapeCluster>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaGetLastError();
    for (int cs : {16, 8}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kCholThreads); cfg.dynamicSmemBytes = kCholSmemBytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, level_chol_pcg_kernel<kShapeCluster>, &cfg) == cudaSuccess && nclusters >= 1) {
        cluster_max = cs;
        break;
      }
      cudaGetLastError();
    }
    pool_cache_set(device, kCacheCholCluster, cluster_max);
  }
  if (want_cluster) C->cluster_ctas = std::max(cluster_max, 0);
  PGO_TRY(chol_alloc(C, device, &C->partials, (size_t)8 * 4 * std::max(C->max_ctas, 16)));
  CUDA_TRY(cudaStreamSynchronize(stream));
  C->usable = true;
  return 0;
}

struct LmDiagonal {           // LevenbergMarquardtStrategy::ComputeStep's D = sqrt(diagonal / radius), squared
  int mode;                   // 0 new diagonal from H, 1 reuse, 2 dlm given
  double min_diag, max_diag, radius;
  double* diagonal; double* dlm;
};

// Factor (H + D) and solve (H + D) x = b by PCG preconditioned with the factor, one launch.
static int level_chol_solve(LevelChol* C, BsrView A, const LmDiagonal& lm, const LmState* lm_state, const unsigned char* active, const double* b,
                            double* x, double* r, double* z, double* q, double* p, double* ax, int max_iterations,
                            double tolerance, double accept, int num_ctas, DeviceScalars* scalars, cudaStream_t stream,
                            long long* launches) {
  CholParams P;
  P.lm = lm_state;
  P.A = A; P.lm_mode = lm.mode; P.min_diag = lm.min_diag; P.max_diag = lm.max_diag; P.radius = lm.radius;
  P.diagonal = lm.diagonal; P.dlm = lm.dlm; P.active = active;
  P.b = b; P.x = x; P.r = r; P.z = z; P.q = q; P.p = p; P.ax = ax;
  P.num_levels = C->num_levels; P.n_nodes = C->n_nodes;
  P.level_ptr = C->level_ptr; P.level_split = C->level_split; P.col_row = C->col_row; P.l2a = C->l2a; P.nodes = C->nodes;
  P.tasks = C->tasks; P.Lblk = C->Lblk; P.Ldiag = C->Ldiag; P.vt = C->vt;
  P.n_slots = C->n_slots; P.partials = C->partials; P.barrier = C->barrier; P.scalars = scalars;
  P.max_iterations = max_iterations; P.tolerance = tolerance; P.accept = accept;
  static const bool want_timeline = getenv("PGO_TIMELINE") != nullptr;
  static unsigned long long* timeline_d = nullptr;
  P.timeline = nullptr;
  if (want_timeline) {
    if (!timeline_d) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&timeline_d), 2001 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(timeline_d, 0, 2001 * sizeof(unsigned long long), stream));

    P.timeline = timeline_d;
  }
  CUDA_TRY(cudaMemsetAsync(C->barrier, 0, 4 * sizeof(unsigned int), stream));
  // launch shape: one CTA for tiny factorisations, one cluster for small ones, else the cooperative grid.
  // PGO_CHOL_SHAPE=grid|cluster|block overrides (measurement aid).
  static const char* shape_env = getenv("PGO_CHOL_SHAPE");
  static const int block_max_nodes = getenv("PGO_CHOL_BLOCK_MAX") ? atoi(getenv("PGO_CHOL_BLOCK_MAX")) : kCholBlockMaxNodes;
  int shape = C->cluster_ctas > 0 ? kShapeCluster : kShapeGrid;
  // a factorisation of a few hundred nodes cannot use more than one SM's warps per level: one CTA, block barriers
  if (C->n_nodes <= block_max_nodes) shape = kShapeBlock;
  if (shape_env) {
    if (!strcmp(shape_env, "grid")) shape = kShapeGrid;
    else if (!strcmp(shape_env, "block")) shape = kShapeBlock;
    else if (!strcmp(shape_env, "cluster") && C->cluster_ctas > 0) shape = kShapeCluster;
  }
  if (num_ctas > 0 && (shape == kShapeCluster || shape == kShapeBlock)) shape = kShapeGrid;
  P.first_level = 0; P.setup_done = 0;
  const int sms = C->max_ctas;   // one CTA per SM
  auto launch_setup = [&]() {
    const int s_items = (int)std::min<long long>(std::max<long long>((long long)A.n * 6, C->n_slots * 6 / 4), 1LL << 30);
    const int s_ctas = std::max(1, std::min((s_items + kCholThreads - 1) / kCholThreads, 2 * sms));
    chol_wide_kernel<<<s_ctas, kCholThreads, kCholSmemBytes, stream>>>(P, 0, 0, 0, 0);
    if (launches) (*launches)++;
    P.setup_done = 1;
  };
  if (shape == kShapeBlock) {
    level_chol_pcg_kernel<kShapeBlock><<<1, kCholThreads, kCholSmemBytes, stream>>>(P);
    CUDA_TRY(cudaGetLastError());
  } else if (shape == kShapeCluster) {
    // S phase and the leading wide levels as grid-wide launches (see chol_wide_kernel)
    int first = 0;
    while (first < C->num_levels && C->level_mode_h[first] != 1 &&
           C->level_ptr_h[first + 1] - C->level_ptr_h[first] >= kCholWideLevelNodes) ++first;
    if (first > 0) {
      launch_setup();
      for (int l = 0; l < first; ++l) {
        const int k0 = C->level_ptr_h[l], k1 = C->level_ptr_h[l + 1], mode = C->level_mode_h[l] & 0xff;
        const int per_cta = (kCholThreads / 32) * (32 / mode);
        const int ctas = std::max(1, std::min((k1 - k0 + per_cta - 1) / per_cta, sms));
        chol_wide_kernel<<<ctas, kCholThreads, kCholSmemBytes, stream>>>(P, 1, mode, k0, k1);
      }
      CUDA_TRY(cudaGetLastError());
      if (launches) (*launches) += first;
      P.first_level = first;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C->cluster_ctas); cfg.blockDim = dim3(kCholThreads); cfg.dynamicSmemBytes = kCholSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C->cluster_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, level_chol_pcg_kernel<kShapeCluster>, P));
  } else {
    int grid = num_ctas > 0 ? num_ctas : std::max(1, (A.n + 79) / 80);
    grid = std::max(1, std::min(grid, C->max_ctas));
    void* args[] = {&P};
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)level_chol_pcg_kernel<kShapeGrid>, dim3(grid), dim3(kCholThreads), args, kCholSmemBytes, stream));
  }
  if (launches) (*launches)++;
  if (want_timeline) {
    std::vector<unsigned long long> h(2001);
    CUDA_TRY(cudaMemcpyAsync(h.data(), timeline_d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    fprintf(stderr, "[pgo timeline] shape=%d levels=%d marks=%llu :", shape, C->num_levels, h[0]);
    for (unsigned long long k = 0; k < h[0] && k < 1000; ++k)
      fprintf(stderr, " %llu@%.2f", h[2 + 2 * k], (double)(h[1 + 2 * k] - h[1]) * 1e-3);
    fprintf(stderr, "\n");
  }
  return 0;
}

}  // namespace pgo
