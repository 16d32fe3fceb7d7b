// pgo_level_chol.cuh -- level-scheduled sparse block Cholesky used as the PCG preconditioner.
//
// Why: the reference solves the damped normal equations exactly (SPARSE_NORMAL_CHOLESKY,
// REF/test/pose_graph_ceres_plus_finial.cpp:505).  Odometry-chain pose graphs such as KITTI-00 are
// beam-like: at late LM radii block-Jacobi PCG needs 10^5 iterations to follow that exact path.
// Here M = L L^T is the exact factor of H + D on a parallel elimination order, so PCG (whose
// SpMV is still bsr6_row) converges in 1-3 iterations and acts as iterative refinement.
//
// Host (once per graph): rounds of independent-set minimum-degree elimination -> levels; nodes of
// a level are mutually non-adjacent, so their columns factor in parallel.  Structure of L, the
// A->L scatter map, per-node Schur update tasks and per-node row lists are precomputed.
// Device (one persistent cooperative kernel per solve): scatter A, factor level by level
// (1 grid barrier per level when every column is short, 2 otherwise), then PCG whose M^-1 is a
// forward + backward sweep over the levels (gather form, deterministic).
#pragma once

#include <algorithm>
#include <vector>

#include "pgo_kernels.cuh"

namespace pgo {

struct CholTask { int p; int q; int target; };   // target >= 0: L slot ; < 0: diagonal of node (-target-1)

struct LevelChol {
  bool usable = false;
  int n_nodes = 0;            // active poses
  int N = 0;
  long long factor_blocks = 0;
  int num_levels = 0;
  int max_degree = 0;
  long long n_slots = 0, n_tasks = 0;
  // device
  int* level_ptr = nullptr;     // [L+1] into order[]
  int* level_split = nullptr;   // [L] 1 = two-phase level
  int* order = nullptr;         // [n_nodes] pose ids in elimination order
  int* col_ptr = nullptr;       // [n_nodes+1] by elimination position -> slots
  int* col_row = nullptr;       // [n_slots] row pose id of each slot
  int* row_ptr = nullptr;       // [N+1] by pose id -> row entries
  int* row_slot = nullptr;      // [n_slots] slot
  int* row_col = nullptr;       // [n_slots] column pose id
  int* task_ptr = nullptr;      // [n_nodes+1] by elimination position
  CholTask* tasks = nullptr;    // [n_tasks]
  int* a2l = nullptr;           // [nnz_off] BSR off-diagonal entry -> slot or -1
  double* Lblk = nullptr;       // [n_slots][36] row-major (rows: row pose, cols: column pose)
  double* Ldiag = nullptr;      // [N][36] W_vv, then inverse of its lower Cholesky factor
  double* vt = nullptr;         // [N][6] sweep workspace
  double* partials = nullptr;
  unsigned int* barrier = nullptr;
  int max_ctas = 0;
};

static void level_chol_destroy(LevelChol* c) {
  if (!c) return;
  void* ptrs[] = {c->level_ptr, c->level_split, c->order, c->col_ptr, c->col_row, c->row_ptr, c->row_slot, c->row_col,
                  c->task_ptr, c->tasks, c->a2l, c->Lblk, c->Ldiag, c->vt, c->partials, c->barrier};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete c;
}

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
struct CholParams {
  BsrView A;
  const double* dlm;
  const double* b;
  double *x, *r, *z, *q, *p;   // PCG vectors [N][6]
  int num_levels, n_nodes;
  const int *level_ptr, *level_split, *order, *col_ptr, *col_row, *row_ptr, *row_slot, *row_col, *task_ptr, *a2l;
  const CholTask* tasks;
  double *Lblk, *Ldiag, *vt;
  long long n_slots;
  double* partials;
  unsigned int* barrier;
  DeviceScalars* scalars;
  int max_iterations;
  double tolerance;
  int do_factor;
};

// One warp: Cholesky of the 6x6 diagonal block of node v (every lane redundantly), store the inverse
// of the lower factor, then scale the column: L_uv = W_uv * Linv^T (one lane per block row).
__device__ __forceinline__ bool chol_node_factor(const CholParams& P, int k, int v, int lane) {
  double A[6][6];
  double* dv = P.Ldiag + 36 * (size_t)v;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) A[r][c] = __ldcg(dv + r * 6 + c);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[j][j];
#pragma unroll
    for (int t = 0; t < 6; ++t) if (t < j) s -= A[j][t] * A[j][t];
    if (!(s > 0.0)) { ok = false; s = 1.0; }
    const double l = sqrt(s), il = 1.0 / l;
    A[j][j] = l;
#pragma unroll
    for (int r = 0; r < 6; ++r) if (r > j) {
      double t2 = A[r][j];
#pragma unroll
      for (int t = 0; t < 6; ++t) if (t < j) t2 -= A[r][t] * A[j][t];
      A[r][j] = t2 * il;
    }
  }
  double Li[6][6];   // inverse of the lower factor
#pragma unroll
  for (int c = 0; c < 6; ++c)
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r < c) { Li[r][c] = 0.0; continue; }
      double t = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int t3 = 0; t3 < 6; ++t3) if (t3 >= c && t3 < r) t -= A[r][t3] * Li[t3][c];
      Li[r][c] = t / A[r][r];
    }
  __syncwarp();
  // store Linv (row-major) : lanes 0..5 write one row each
  if (lane < 6) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double val = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) if (r == lane) val = Li[r][c];
      dv[lane * 6 + c] = val;
    }
  }
  // column scaling: item = (block, row)
  const int p0 = P.col_ptr[k], p1 = P.col_ptr[k + 1];
  const int items = (p1 - p0) * 6;
  for (int it = lane; it < items; it += 32) {
    const int blk = it / 6, r = it - blk * 6;
    double* w = P.Lblk + 36 * (size_t)(p0 + blk) + r * 6;
    double wr[6], o[6];
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
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s = fma(a[k], __ldcg(lq + c * 6 + k), s);
    atomicAdd(out + c, -s);
  }
}

__global__ void __launch_bounds__(kPcgThreads) level_chol_pcg_kernel(const CholParams P) {
  __shared__ double red[kPcgThreads / 32];
  __shared__ double bcast;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps_per_cta = kPcgThreads / 32;
  const int gw = blockIdx.x * warps_per_cta + warp;
  const int nw = gridDim.x * warps_per_cta;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gthreads = gridDim.x * blockDim.x;
  const int grp = lane / 6, r6 = lane - grp * 6;
  const bool lane_on = grp < kRowsPerWarp;
  const int n = P.A.n;
  const int G = gridDim.x;
  unsigned int epoch = 0;
  int fail = 0;

  if (P.do_factor) {
    // ---- scatter A + D into the factor storage ----
    for (long long i = gtid; i < P.n_slots * 36; i += gthreads) P.Lblk[i] = 0.0;
    grid_barrier(P.barrier, epoch);
    for (long long e = gtid; e < (long long)P.A.row_ptr[n] * 36; e += gthreads) {
      const int pe = (int)(e / 36), k = (int)(e - 36LL * pe);
      const int slot = P.a2l[pe];
      if (slot >= 0) P.Lblk[36 * (size_t)slot + k] = P.A.Hoff[36 * (size_t)pe + pidx(k / 6, k % 6)];
    }
    for (long long e = gtid; e < (long long)n * 36; e += gthreads) {
      const int i = (int)(e / 36), k = (int)(e - 36LL * i);
      const int rr = k / 6, cc = k % 6;
      double v = P.A.Hdiag[36 * (size_t)i + pidx(rr, cc)];
      if (rr == cc) v += P.dlm[6 * (size_t)i + rr];
      P.Ldiag[e] = v;
    }
    grid_barrier(P.barrier, epoch);
    // ---- numeric factorisation, level by level ----
    for (int l = 0; l < P.num_levels; ++l) {
      const int k0 = P.level_ptr[l], k1 = P.level_ptr[l + 1];
      const bool split = P.level_split[l] != 0;
      for (int k = k0 + gw; k < k1; k += nw) {
        const int v = P.order[k];
        if (!chol_node_factor(P, k, v, lane)) { fail = 1; if (lane == 0) atomicExch(P.barrier + 1, 1u); }
        if (!split) {
          __syncwarp();
          const int t0 = P.task_ptr[k], t1 = P.task_ptr[k + 1];
          for (int it = lane; it < (t1 - t0) * 6; it += 32) chol_update_item(P, P.tasks[t0 + it / 6], it % 6);
        }
      }
      if (split) {
        grid_barrier(P.barrier, epoch);
        const long long t0 = P.task_ptr[k0], t1 = P.task_ptr[k1];
        for (long long it = (long long)gtid; it < (t1 - t0) * 6; it += gthreads) chol_update_item(P, P.tasks[t0 + it / 6], (int)(it % 6));
      }
      grid_barrier(P.barrier, epoch);
    }
  }

  // z = (L L^T)^-1 src  (z also used as the backward-sweep output); inactive poses get 0.
  auto apply_minv = [&](const double* src, double* dst) {
    // forward: y_v = Linv_v (src_v - sum_{w earlier} L_vw y_w), levels ascending; y kept in vt
    for (int l = 0; l < P.num_levels; ++l) {
      const int k0 = P.level_ptr[l], k1 = P.level_ptr[l + 1];
      for (int k = k0 + gw; k < k1; k += nw) {
        const int v = P.order[k];
        const int e0 = P.row_ptr[v], e1 = P.row_ptr[v + 1];
        double acc = 0.0;   // lane (grp, r6): partial of row r6 over this group's blocks
        if (lane_on) {
          for (int e = e0 + grp; e < e1; e += kRowsPerWarp) {
            const double* L = P.Lblk + 36 * (size_t)P.row_slot[e] + r6 * 6;
            const double* y = P.vt + 6 * (size_t)P.row_col[e];
#pragma unroll
            for (int c = 0; c < 6; ++c) acc = fma(__ldcg(L + c), __ldcg(y + c), acc);
          }
        }
        // sum the 5 groups (fixed order)
        double tot = 0.0;
#pragma unroll
        for (int gq = 0; gq < kRowsPerWarp; ++gq) tot += __shfl_sync(0xffffffffu, acc, gq * 6 + (lane % 6));
        const double tv = __ldcg(src + 6 * (size_t)v + (lane % 6)) - tot;     // lanes 0..5 hold t_v[0..5]
        double yv = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double tc = __shfl_sync(0xffffffffu, tv, c);
          if (lane < 6 && c <= lane) yv = fma(__ldcg(P.Ldiag + 36 * (size_t)v + lane * 6 + c), tc, yv);
        }
        if (lane < 6) P.vt[6 * (size_t)v + lane] = yv;
      }
      grid_barrier(P.barrier, epoch);
    }
    // backward: x_v = Linv_v^T (y_v - sum_{u later} L_uv^T x_u), levels descending
    for (int l = P.num_levels - 1; l >= 0; --l) {
      const int k0 = P.level_ptr[l], k1 = P.level_ptr[l + 1];
      for (int k = k0 + gw; k < k1; k += nw) {
        const int v = P.order[k];
        const int p0 = P.col_ptr[k], p1 = P.col_ptr[k + 1];
        double acc = 0.0;   // lane (grp, c = r6): sum_u sum_r L_uv[r][c] x_u[r]
        if (lane_on) {
          for (int p = p0 + grp; p < p1; p += kRowsPerWarp) {
            const double* L = P.Lblk + 36 * (size_t)p + r6;
            const double* xu = dst + 6 * (size_t)P.col_row[p];
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
        if (lane < 6) dst[6 * (size_t)v + lane] = xv;
      }
      grid_barrier(P.barrier, epoch);
    }
  };

  // ---- PCG: x = 0, r = b ----
  const int n6 = 6 * n;
  for (int k = gtid; k < n6; k += gthreads) { P.x[k] = 0.0; P.r[k] = P.b[k]; P.z[k] = 0.0; P.vt[k] = 0.0; }
  grid_barrier(P.barrier, epoch);
  apply_minv(P.r, P.z);
  double acc = 0.0;
  for (int k = gtid; k < n6; k += gthreads) { const double zv = __ldcg(P.z + k); P.p[k] = zv; acc = fma(P.r[k], zv, acc); }
  acc = cta_sum(acc, red);
  int slot = 0;
  if (threadIdx.x == 0) P.partials[(size_t)slot * G + blockIdx.x] = acc;
  grid_barrier(P.barrier, epoch);
  const double rho0 = reduce_partials(P.partials + (size_t)slot * G, G, &bcast);
  double rho = rho0;
  int iter = 0, flag = 0;
  (void)fail;
  const double stop = P.tolerance * P.tolerance * rho0;
  if (rho0 > 0.0 && isfinite(rho0)) {
    for (;;) {
      // q = A p ; pq = p.q
      acc = 0.0;
      for (int base = gw * kRowsPerWarp; base < n; base += nw * kRowsPerWarp) {
        const int i = base + grp;
        if (lane_on && i < n) {
          const double qv = bsr6_row<true>(P.A.Hdiag, P.A.Hoff, P.A.row_ptr, P.A.col_idx, P.p, P.dlm, i, r6);
          const size_t k = 6 * (size_t)i + r6;
          P.q[k] = qv;
          acc = fma(qv, __ldcg(P.p + k), acc);
        }
      }
      acc = cta_sum(acc, red);
      slot = (slot + 1) % 4;
      if (threadIdx.x == 0) P.partials[(size_t)slot * G + blockIdx.x] = acc;
      grid_barrier(P.barrier, epoch);
      const double pq = reduce_partials(P.partials + (size_t)slot * G, G, &bcast);
      if (!(pq > 0.0) || !isfinite(pq)) { flag = 2; break; }
      const double alpha = rho / pq;
      ++iter;
      for (int k = gtid; k < n6; k += gthreads) {
        P.x[k] += alpha * __ldcg(P.p + k);
        P.r[k] -= alpha * __ldcg(P.q + k);
      }
      grid_barrier(P.barrier, epoch);
      apply_minv(P.r, P.z);
      acc = 0.0;
      for (int k = gtid; k < n6; k += gthreads) acc = fma(P.r[k], __ldcg(P.z + k), acc);
      acc = cta_sum(acc, red);
      slot = (slot + 1) % 4;
      if (threadIdx.x == 0) P.partials[(size_t)slot * G + blockIdx.x] = acc;
      grid_barrier(P.barrier, epoch);
      const double rho_new = reduce_partials(P.partials + (size_t)slot * G, G, &bcast);
      const double beta = rho_new / rho;
      rho = rho_new;
      if (fabs(rho) <= stop) break;
      if (iter >= P.max_iterations) { flag = flag ? flag : 1; break; }
      for (int k = gtid; k < n6; k += gthreads) P.p[k] = __ldcg(P.z + k) + beta * P.p[k];
      grid_barrier(P.barrier, epoch);
    }
  }
  // ---- epilogue: x^T b, x^T (H + D) x, x^T D x ----
  grid_barrier(P.barrier, epoch);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int base = gw * kRowsPerWarp; base < n; base += nw * kRowsPerWarp) {
    const int i = base + grp;
    if (lane_on && i < n) {
      const double ax = bsr6_row<true>(P.A.Hdiag, P.A.Hoff, P.A.row_ptr, P.A.col_idx, P.x, P.dlm, i, r6);
      const size_t k = 6 * (size_t)i + r6;
      const double xv = __ldcg(P.x + k);
      a0 = fma(xv, P.b[k], a0); a1 = fma(xv, ax, a1); a2 = fma(xv * xv, P.dlm[k], a2);
    }
  }
  a0 = cta_sum(a0, red); a1 = cta_sum(a1, red); a2 = cta_sum(a2, red);
  double* pe = P.partials + (size_t)4 * G;
  if (threadIdx.x == 0) { pe[blockIdx.x] = a0; pe[G + blockIdx.x] = a1; pe[2 * G + blockIdx.x] = a2; }
  grid_barrier(P.barrier, epoch);
  const double s0 = reduce_partials(pe, G, &bcast);
  const double s1 = reduce_partials(pe + G, G, &bcast);
  const double s2 = reduce_partials(pe + 2 * G, G, &bcast);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    P.scalars->xtb = s0; P.scalars->xtAx = s1; P.scalars->xtDx = s2;
    P.scalars->pcg_gamma0 = rho0; P.scalars->pcg_gamma = fabs(rho);
    unsigned int failed;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(failed) : "l"(P.barrier + 1) : "memory");
    P.scalars->pcg_iterations = iter; P.scalars->pcg_flag = failed ? 3 : flag;
  }
}

// ---------------------------------------------------------------------------------------------
// host side: symbolic analysis
// ---------------------------------------------------------------------------------------------
template <typename Tp>
static int chol_upload(Tp** dst, const std::vector<Tp>& src, cudaStream_t stream) {
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(dst), std::max<size_t>(src.size(), 1) * sizeof(Tp)));
  if (!src.empty()) CUDA_TRY(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(Tp), cudaMemcpyHostToDevice, stream));
  return 0;
}

static int level_chol_analyze(LevelChol** out, int N, const unsigned char* active, const int* a_row_ptr,
                              const int* a_col_idx, double max_fill_ratio, cudaStream_t stream) {
  LevelChol* C = new LevelChol();
  *out = C;
  C->N = N;
  std::vector<std::vector<int>> adj(N);
  int n_nodes = 0;
  long long a_off = 0;
  for (int i = 0; i < N; ++i) {
    if (!active[i]) continue;
    ++n_nodes;
    adj[i].assign(a_col_idx + a_row_ptr[i], a_col_idx + a_row_ptr[i + 1]);   // sorted, symmetric, active only
    a_off += (long long)adj[i].size();
  }
  C->n_nodes = n_nodes;
  if (n_nodes == 0) return 0;
  const long long fill_cap = (long long)std::min(max_fill_ratio * (double)(a_off / 2 + n_nodes) + 64.0, 4.0e9);

  std::vector<int> pos(N, -1), order;
  order.reserve(n_nodes);
  std::vector<int> level_ptr(1, 0), level_split;
  std::vector<std::vector<int>> col_rows(N);       // structure of L's column v (row pose ids, sorted by id)
  std::vector<unsigned char> alive(N, 0), blocked(N, 0);
  std::vector<int> alive_list;
  for (int i = 0; i < N; ++i) if (active[i]) { alive[i] = 1; alive_list.push_back(i); }
  std::vector<std::pair<int, int>> cand;
  std::vector<int> tmp, sel;
  long long slots = 0;
  while (!alive_list.empty()) {
    int dmin = 1 << 30;
    for (int v : alive_list) dmin = std::min(dmin, (int)adj[v].size());
    const int thr = 2 * dmin + 2;
    cand.clear();
    for (int v : alive_list) if ((int)adj[v].size() <= thr) cand.emplace_back((int)adj[v].size(), v);
    std::sort(cand.begin(), cand.end());
    sel.clear();
    for (auto& dv : cand) {
      const int v = dv.second;
      if (blocked[v]) continue;
      sel.push_back(v);
      blocked[v] = 1;
      for (int u : adj[v]) blocked[u] = 1;
    }
    int lvl_maxdeg = 0;
    for (int v : sel) {
      pos[v] = (int)order.size();
      order.push_back(v);
      col_rows[v] = adj[v];
      slots += (long long)adj[v].size();
      lvl_maxdeg = std::max(lvl_maxdeg, (int)adj[v].size());
    }
    if (slots > fill_cap || (int)level_ptr.size() > 8192) return 0;   // not usable (too much fill / too deep)
    for (int v : sel) {
      const std::vector<int>& nb = col_rows[v];
      for (int u : nb) {
        // adj[u] = (adj[u] U nb) \ {u, v}
        tmp.clear();
        std::set_union(adj[u].begin(), adj[u].end(), nb.begin(), nb.end(), std::back_inserter(tmp));
        adj[u].clear();
        for (int x : tmp) if (x != u && x != v) adj[u].push_back(x);
      }
      alive[v] = 0;
      std::vector<int>().swap(adj[v]);
    }
    // unblock
    for (auto& dv : cand) { blocked[dv.second] = 0; }
    for (int v : sel) for (int u : col_rows[v]) blocked[u] = 0;
    size_t w = 0;
    for (size_t k = 0; k < alive_list.size(); ++k) if (alive[alive_list[k]]) alive_list[w++] = alive_list[k];
    alive_list.resize(w);
    level_ptr.push_back((int)order.size());
    level_split.push_back(lvl_maxdeg > 10 ? 1 : 0);
    C->max_degree = std::max(C->max_degree, lvl_maxdeg);
  }
  C->num_levels = (int)level_split.size();
  C->n_slots = slots;
  C->factor_blocks = slots + n_nodes;

  // column-major slots by elimination position
  std::vector<int> col_ptr(n_nodes + 1, 0), col_row((size_t)slots);
  for (int k = 0; k < n_nodes; ++k) col_ptr[k + 1] = col_ptr[k] + (int)col_rows[order[k]].size();
  for (int k = 0; k < n_nodes; ++k) std::copy(col_rows[order[k]].begin(), col_rows[order[k]].end(), col_row.begin() + col_ptr[k]);
  auto slot_of = [&](int row, int col) -> int {   // block (row, col), col eliminated first
    const int k = pos[col];
    const int* b = col_row.data() + col_ptr[k];
    const int* e = col_row.data() + col_ptr[k + 1];
    const int* it = std::lower_bound(b, e, row);
    return (it != e && *it == row) ? (int)(it - col_row.data()) : -1;
  };
  // row lists (by pose id)
  std::vector<int> row_ptr(N + 1, 0), row_slot((size_t)slots), row_col((size_t)slots);
  for (long long s = 0; s < slots; ++s) row_ptr[col_row[s] + 1]++;
  for (int i = 0; i < N; ++i) row_ptr[i + 1] += row_ptr[i];
  {
    std::vector<int> fillp(row_ptr.begin(), row_ptr.end() - 1);
    for (int k = 0; k < n_nodes; ++k)
      for (int s = col_ptr[k]; s < col_ptr[k + 1]; ++s) { const int u = col_row[s]; row_slot[fillp[u]] = s; row_col[fillp[u]] = order[k]; fillp[u]++; }
  }
  // Schur update tasks per node
  std::vector<int> task_ptr(n_nodes + 1, 0);
  std::vector<CholTask> tasks;
  for (int k = 0; k < n_nodes; ++k) {
    const int p0 = col_ptr[k], p1 = col_ptr[k + 1];
    for (int p = p0; p < p1; ++p) {
      tasks.push_back({p, p, -col_row[p] - 1});                  // diagonal of row(p)
      for (int q = p0; q < p1; ++q) {
        if (q == p) continue;
        const int u = col_row[p], w = col_row[q];
        if (pos[u] > pos[w]) {                                   // block (u, w): rows u, cols w = L_p L_q^T
          const int t = slot_of(u, w);
          if (t < 0) { return set_error(PGO_ERR_NUMERICAL, "level Cholesky: missing fill slot"); }
          tasks.push_back({p, q, t});
        }
      }
    }
    task_ptr[k + 1] = (int)tasks.size();
  }
  C->n_tasks = (long long)tasks.size();
  // A (BSR off-diagonal) -> L slot
  std::vector<int> a2l((size_t)a_row_ptr[N], -1);
  for (int i = 0; i < N; ++i)
    for (int p = a_row_ptr[i]; p < a_row_ptr[i + 1]; ++p) {
      const int j = a_col_idx[p];
      if (active[i] && active[j] && pos[j] < pos[i]) a2l[p] = slot_of(i, j);
    }

  PGO_TRY(chol_upload(&C->level_ptr, level_ptr, stream));
  PGO_TRY(chol_upload(&C->level_split, level_split, stream));
  PGO_TRY(chol_upload(&C->order, order, stream));
  PGO_TRY(chol_upload(&C->col_ptr, col_ptr, stream));
  PGO_TRY(chol_upload(&C->col_row, col_row, stream));
  PGO_TRY(chol_upload(&C->row_ptr, row_ptr, stream));
  PGO_TRY(chol_upload(&C->row_slot, row_slot, stream));
  PGO_TRY(chol_upload(&C->row_col, row_col, stream));
  PGO_TRY(chol_upload(&C->task_ptr, task_ptr, stream));
  PGO_TRY(chol_upload(&C->tasks, tasks, stream));
  PGO_TRY(chol_upload(&C->a2l, a2l, stream));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&C->Lblk), std::max<size_t>((size_t)slots * 36, 1) * sizeof(double)));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&C->Ldiag), (size_t)N * 36 * sizeof(double)));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&C->vt), (size_t)N * 6 * sizeof(double)));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&C->barrier), 4 * sizeof(unsigned int)));
  int dev = 0, sms = 0, per_sm = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, level_chol_pcg_kernel, kPcgThreads, 0));
  C->max_ctas = std::max(1, std::min(per_sm, 2) * sms);
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&C->partials), (size_t)8 * C->max_ctas * sizeof(double)));
  CUDA_TRY(cudaStreamSynchronize(stream));
  C->usable = true;
  return 0;
}

// Factor (H + D) and solve (H + D) x = b by PCG preconditioned with the factor, one launch.
static int level_chol_solve(LevelChol* C, BsrView A, const double* dlm, const double* b, double* x, double* r, double* z,
                            double* q, double* p, int max_iterations, double tolerance, int num_ctas,
                            DeviceScalars* scalars, cudaStream_t stream, long long* launches) {
  CholParams P;
  P.A = A; P.dlm = dlm; P.b = b; P.x = x; P.r = r; P.z = z; P.q = q; P.p = p;
  P.num_levels = C->num_levels; P.n_nodes = C->n_nodes;
  P.level_ptr = C->level_ptr; P.level_split = C->level_split; P.order = C->order; P.col_ptr = C->col_ptr;
  P.col_row = C->col_row; P.row_ptr = C->row_ptr; P.row_slot = C->row_slot; P.row_col = C->row_col;
  P.task_ptr = C->task_ptr; P.a2l = C->a2l; P.tasks = C->tasks; P.Lblk = C->Lblk; P.Ldiag = C->Ldiag; P.vt = C->vt;
  P.n_slots = C->n_slots; P.partials = C->partials; P.barrier = C->barrier; P.scalars = scalars;
  P.max_iterations = max_iterations; P.tolerance = tolerance; P.do_factor = 1;
  CUDA_TRY(cudaMemsetAsync(C->barrier, 0, 4 * sizeof(unsigned int), stream));
  int grid = num_ctas > 0 ? num_ctas : std::max(1, (A.n + 39) / 40);
  grid = std::max(1, std::min(grid, C->max_ctas));
  void* args[] = {&P};
  CUDA_TRY(cudaLaunchCooperativeKernel((void*)level_chol_pcg_kernel, dim3(grid), dim3(kPcgThreads), args, 0, stream));
  if (launches) (*launches)++;
  return 0;
}

}  // namespace pgo
