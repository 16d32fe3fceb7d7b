// pgo_kernels.cuh -- the sm_100a kernels of the pose-graph LM solver.
//
//  linearize_kernel   : per edge residual + both 6x6 Jacobians + J^T J blocks + J^T r, fused.
//                       Edge tiles are staged global->shared by 1-D bulk TMA (cp.async.bulk +
//                       mbarrier, 3-deep ring per warp) with L1 prefetch of the next tile's pose
//                       gathers; poses are gathered with 128-bit loads; every output block goes
//                       through a per-warp shared-memory staging tile so that the fp64 RED atomics
//                       (block-CSR diagonal, gradient) and the off-diagonal stores hit whole sectors.
//  pcg_kernel         : persistent block-Jacobi PCG (Chronopoulos-Gear form, 2 barriers per
//                       iteration: atomic grid barrier, or the cluster barrier for tiny graphs);
//                       its SpMV is bsr6_row() -- 6 lanes per 6x6 block row.
//  spmv_kernel        : the same bsr6_row() as a stand-alone launch (tests, bench, and the
//                       stream-ordered PCG of large / multi-GPU solves, with w.u fused in).
#pragma once

#include "pgo_common.cuh"
#include "pgo_edge_math.cuh"
#include "pgo_lm.cuh"

namespace pgo {

enum LinMode { kLinFull = 0, kLinCost = 1, kLinEval = 2 };

struct LinParams {
  int n_edges;
  int n_tiles;
  int n_own;                    // block rows stored here: poses >= n_own are halo copies owned by another rank (multi-GPU);
                                // their diagonal blocks / gradient are that rank's, and an edge's cost is its id_begin owner's
  const EdgeCoreTile* core;
  const EdgeInfoTile* info;     // nullptr when identity
  const double* poses;          // [N][8]
  const double* scale;          // [N][6]
  double* Hdiag;                // [N][36] panel
  double* Hoff;                 // [nnz][36] panel
  double* grad;                 // [N][6]
  DeviceScalars* scalars;
  int loss_type;
  double loss_a;
  const double* edge_loss;      // per-edge loss in processing order, encoded: 0 trivial, +a Huber(a), -a Cauchy(a); or nullptr
  // kLinEval outputs
  double* res_out;              // [E][6]
  double* jac_out;              // [E][2][36] row-major
  const LmState* lm;            // device-resident LM loop: return at once when lm->done (else nullptr)
};

__device__ __forceinline__ void load_pose(const double* poses, int i, double* p) {
  const double2* q = reinterpret_cast<const double2*>(poses + 8 * (size_t)i);
  const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
  p[0] = a.x; p[1] = a.y; p[2] = b.x; p[3] = b.y; p[4] = c.x; p[5] = c.y; p[6] = d.x;
}
__device__ __forceinline__ void load_vec6(const double* v, int i, double* s) {
  const double2* q = reinterpret_cast<const double2*>(v + 6 * (size_t)i);
  const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y; s[4] = c.x; s[5] = c.y;
}

// 3x3 = X^T Y over the 6 rows of two 6x3 panels
__device__ __forceinline__ void xty(const double (&X)[6][3], const double (&Y)[6][3], double (&P)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s = fma(X[k][i], Y[k][j], s);
      P[i][j] = s;
    }
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// kStages-deep per-warp ring of edge tiles (bulk TMA); tile t+2 is in flight and tile t+1's pose / scale gathers
// are prefetched into L1 while tile t is processed.
constexpr int kLinStages = 3;

constexpr int kLinInfoStages = 2;   // the 9 KB sqrt-information tiles are only double-buffered (shared-memory budget)

template <bool kIdentityInfo>
constexpr int lin_smem_bytes() {
  return kLinWarps * (kLinStages * kCoreTileBytes + (kIdentityInfo ? kInfoTileBytes : kLinInfoStages * kInfoTileBytes)) +
         kLinWarps * (kLinStages + kLinInfoStages) * 8 + kLinWarps * 64 * 4;
}

template <bool kIdentityInfo, int kMode, int kMinBlocks>
__global__ void __launch_bounds__(kLinWarps * 32, kMinBlocks) linearize_kernel(const LinParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (p.lm && p.lm->done) return;
  // per warp: core ring, then the info ring; the output staging tile [32][36] aliases the CURRENT info stage (its
  // sqrt-information is dead once the Jacobians are in registers) or, without information tiles, a dedicated buffer
  constexpr int kWarpBytes = kLinStages * kCoreTileBytes + (kIdentityInfo ? kInfoTileBytes : kLinInfoStages * kInfoTileBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbase = smem_raw + (size_t)warp * kWarpBytes;                 // core ring
  unsigned char* ibase = wbase + kLinStages * kCoreTileBytes;                  // info ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kLinWarps * kWarpBytes) + warp * (kLinStages + kLinInfoStages);
  uint64_t* ibars = bars + kLinStages;
  int* sidx = reinterpret_cast<int*>(smem_raw + (size_t)kLinWarps * kWarpBytes + kLinWarps * (kLinStages + kLinInfoStages) * 8) + warp * 64;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLinStages + kLinInfoStages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  const int gw = blockIdx.x * kLinWarps + warp;
  const int nw = gridDim.x * kLinWarps;
  auto issue_core = [&](int tile, int st) {
    mbar_expect_tx(&bars[st], kCoreTileBytes);
    tma_load_1d(wbase + (size_t)st * kCoreTileBytes, p.core + tile, kCoreTileBytes, &bars[st]);
  };
  auto issue_info = [&](int tile, int st) {
    if (!kIdentityInfo) {
      // the stage doubled as the generic-proxy staging tile two iterations ago: order those accesses before the async-proxy write
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&ibars[st], kInfoTileBytes);
      tma_load_1d(ibase + (size_t)st * kInfoTileBytes, p.info + tile, kInfoTileBytes, &ibars[st]);
    }
  };

  int stage = 0, istage = 0;
  uint32_t round = 0, iround = 0;   // completed trips around each ring = parity of the current stage's phase
  double cost_acc = 0.0;
  if (lane == 0) {
    if (gw < p.n_tiles) { issue_core(gw, 0); issue_info(gw, 0); }
    if (gw + nw < p.n_tiles) issue_core(gw + nw, 1);
  }
  auto prefetch_gathers = [&](int t, int st, uint32_t parity) {
    mbar_wait(&bars[st], parity);
    const EdgeCoreTile* nt = reinterpret_cast<const EdgeCoreTile*>(wbase + (size_t)st * kCoreTileBytes);
    if (t * kTile + lane < p.n_edges) {
      const int na = nt->a[lane], nbp = nt->b[lane];
      prefetch_l1(p.poses + 8 * (size_t)na);
      prefetch_l1(p.poses + 8 * (size_t)nbp);
      if (kMode != kLinCost) { prefetch_l1(p.scale + 6 * (size_t)na); prefetch_l1(p.scale + 6 * (size_t)nbp); }
    }
  };

  for (int tile = gw; tile < p.n_tiles; tile += nw) {
    const int next = tile + nw, next2 = tile + 2 * nw;
    const int stage1 = (stage + 1 == kLinStages) ? 0 : stage + 1;
    const int stage2 = (stage1 + 1 == kLinStages) ? 0 : stage1 + 1;
    if (lane == 0) {
      if (next2 < p.n_tiles) issue_core(next2, stage2);
      if (next < p.n_tiles) issue_info(next, istage ^ 1);
    }
    mbar_wait(&bars[stage], round & 1);
    if (next < p.n_tiles) prefetch_gathers(next, stage1, (stage1 == 0 ? round + 1 : round) & 1);
    if (!kIdentityInfo) mbar_wait(&ibars[istage], iround & 1);

    const EdgeCoreTile* ct = reinterpret_cast<const EdgeCoreTile*>(wbase + (size_t)stage * kCoreTileBytes);
    const EdgeInfoTile* it = reinterpret_cast<const EdgeInfoTile*>(ibase + (size_t)istage * kInfoTileBytes);
    double* stg = reinterpret_cast<double*>(ibase + (kIdentityInfo ? 0 : (size_t)istage * kInfoTileBytes));
    const int e = tile * kTile + lane;
    const bool valid = e < p.n_edges;
    const int a = valid ? ct->a[lane] : 0;
    const int b = valid ? ct->b[lane] : 0;

    double pa[7], pb[7], m[7];
    load_pose(p.poses, a, pa);
    load_pose(p.poses, b, pb);
#pragma unroll
    for (int k = 0; k < 7; ++k) m[k] = ct->meas[k][lane];
    auto S = [&](int i, int k) -> double { return it->S[i * 6 + k][lane]; };

    if (kMode == kLinCost) {
      double r[6];
      edge_residual_only<kIdentityInfo>(pa, pb, m, S, r);
      double sq = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) sq = fma(r[i], r[i], sq);
      double rho1;
      int lt = p.loss_type;
      double la = p.loss_a;
      if (p.edge_loss != nullptr) { const double code = valid ? __ldg(p.edge_loss + e) : 0.0; lt = code > 0.0 ? 1 : (code < 0.0 ? 2 : 0); la = fabs(code); }
      const double rho = loss_eval(lt, la, sq, rho1);
      if (valid && a < p.n_own) cost_acc += 0.5 * rho;
      __syncwarp();
      stage = stage1;
      if (stage == 0) ++round;
      istage ^= 1;
      if (istage == 0) ++iround;
      continue;
    }

    EdgePanels L;
    edge_linearize<kIdentityInfo>(pa, pb, m, S, L);
    double sq = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) sq = fma(L.r[i], L.r[i], sq);
    double rho1;
    int lt = p.loss_type;
    double la = p.loss_a;
    if (p.edge_loss != nullptr) { const double code = valid ? __ldg(p.edge_loss + e) : 0.0; lt = code > 0.0 ? 1 : (code < 0.0 ? 2 : 0); la = fabs(code); }
    const double rho = loss_eval(lt, la, sq, rho1);
    if (valid && a < p.n_own) cost_acc += 0.5 * rho;
    if (!valid) rho1 = 0.0;   // padded lanes contribute exact zeros

    double sa[6], sb[6];
    load_vec6(p.scale, a, sa);
    load_vec6(p.scale, b, sb);

    if (kMode == kLinEval) {
      // Problem::Evaluate parity outputs: corrected residual and Jacobians (row-major 6x6).
      const double sr = sqrt(rho1);
      if (valid) {
        double* ro = p.res_out + 6 * (size_t)e;
#pragma unroll
        for (int i = 0; i < 6; ++i) ro[i] = sr * L.r[i];
        double* ja = p.jac_out + 72 * (size_t)e;
        double* jb = ja + 36;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            ja[i * 6 + k] = -sr * L.B1[i][k] * sa[k];
            ja[i * 6 + 3 + k] = sr * L.C[i][k] * sa[3 + k];
            jb[i * 6 + k] = sr * L.B1[i][k] * sb[k];
            jb[i * 6 + 3 + k] = -sr * L.B2[i][k] * sb[3 + k];
          }
      }
    }

    // ---- outputs.  Every lane owns one edge, but consecutive lanes must hit consecutive addresses for the L2 to
    // see whole 32-byte sectors (fp64 RED and store throughput is per sector): each lane drops its block into a
    // per-warp shared-memory staging tile [32 edges][36], the warp then walks the tile element-major and issues
    // coalesced RED / stores (4 lanes per sector instead of 1).
    {
      __syncwarp();                                  // all lanes are done with the sqrt-information tile the staging aliases
      auto rot36 = [](int x) -> int { return x >= 36 ? x - 36 : x; };
      const unsigned full = 0xffffffffu;
      const unsigned lt = (1u << lane) - 1u;
      // ---- in-warp combination plan (segmented reduction by pose id).  A pose usually receives several contributions
      // from ONE tile: edges are processed in pose order, so lane l's end pose is lane l-1's begin pose along an odometry
      // chain and the cross edges of a pose sit next to its odometry edge.  Contributions to the same diagonal block /
      // gradient entry are added in shared memory and leave as ONE fp64 RED instead of two to four.
      //   a-side: lanes with the same begin pose form a group; its lowest lane (leader) owns the staging slot
      //   b-side: an end pose that is some lane's begin pose joins that group's slot ("hit"); the remaining end poses
      //           group among themselves and go out in a second, usually almost empty, pass
      const int ta = (valid && a < p.n_own) ? a : -1, tb = (valid && b < p.n_own) ? b : -1;
      const int ida = ta >= 0 ? ta : -1 - lane;
      const unsigned ga = __match_any_sync(full, ida);
      const int la = __ffs(ga) - 1, ra = __popc(ga & lt);
      int hit = -1;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const int idj = __shfl_sync(full, ida, j);
        if (tb >= 0 && idj == tb && hit < 0) hit = j;   // the lowest lane with that begin pose = its group's leader
      }
      const unsigned gh = __match_any_sync(full, hit >= 0 ? hit : -1 - lane);
      const int rh = __popc(gh & lt);
      const unsigned gb = __match_any_sync(full, (tb >= 0 && hit < 0) ? tb : -1 - lane);
      const int lb = __ffs(gb) - 1, rb = __popc(gb & lt);
      const bool b_left = tb >= 0 && hit < 0;
      const bool any_b_left = __any_sync(full, b_left);
      // ---- gradient J^T r (robustified: rho' J^T r), column-scaled: g_a = sa .* [-g1; gc], g_b = sb .* [g1; -g2] ----
      {
        double g1[3], g2[3], gc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double s1 = 0.0, s2 = 0.0, sc = 0.0;
#pragma unroll
          for (int i = 0; i < 6; ++i) { s1 = fma(L.B1[i][k], L.r[i], s1); s2 = fma(L.B2[i][k], L.r[i], s2); sc = fma(L.C[i][k], L.r[i], sc); }
          g1[k] = rho1 * s1; g2[k] = rho1 * s2; gc[k] = rho1 * sc;
        }
        double va[6], vb[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) { va[k] = -sa[k] * g1[k]; va[3 + k] = sa[3 + k] * gc[k]; vb[k] = sb[k] * g1[k]; vb[3 + k] = -sb[3 + k] * g2[k]; }
        // slots [0, 32): a-groups (+ b hits); slots [32, 64): left-over b groups
        if (ra == 0) {
#pragma unroll
          for (int k = 0; k < 6; ++k) stg[lane * 6 + k] = va[k];
        }
        if (b_left && rb == 0) {
#pragma unroll
          for (int k = 0; k < 6; ++k) stg[192 + lane * 6 + k] = vb[k];
        }
        sidx[lane] = ra == 0 ? ta : -1;
        sidx[32 + lane] = (b_left && rb == 0) ? tb : -1;
        __syncwarp();
        for (int r = 1; __any_sync(full, ra >= r); ++r) {
          if (ra == r) {
#pragma unroll
            for (int k = 0; k < 6; ++k) stg[la * 6 + k] += va[k];
          }
          __syncwarp();
        }
        for (int r = 0; __any_sync(full, hit >= 0 && rh >= r); ++r) {
          if (hit >= 0 && rh == r) {
#pragma unroll
            for (int k = 0; k < 6; ++k) stg[hit * 6 + k] += vb[k];
          }
          __syncwarp();
        }
        for (int r = 1; __any_sync(full, b_left && rb >= r); ++r) {
          if (b_left && rb == r) {
#pragma unroll
            for (int k = 0; k < 6; ++k) stg[192 + lb * 6 + k] += vb[k];
          }
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          if (j >= 6 && !any_b_left) break;
          const int item = j * 32 + lane;            // 0..383: [side][slot][6]
          const int blk = item / 6, el = item - blk * 6;
          const int t = sidx[blk];
          if (t >= 0) atomicAdd(p.grad + 6 * (size_t)t + el, stg[item]);
        }
        __syncwarp();
      }

      if (kMode == kLinFull) {
        // ---- 3x3 products of the panels ----
        double P11[3][3], P12[3][3], P22[3][3], P1C[3][3], PCC[3][3], PC2[3][3];
        xty(L.B1, L.B1, P11); xty(L.B1, L.B2, P12); xty(L.B2, L.B2, P22);
        xty(L.B1, L.C, P1C);  xty(L.C, L.C, PCC);   xty(L.C, L.B2, PC2);
        // element accessors of the unscaled J^T J blocks (rho' applied below)
        auto Haa = [&](int r, int c) -> double {
          return (r < 3) ? ((c < 3) ? P11[r][c] : -P1C[r][c - 3]) : ((c < 3) ? -P1C[c][r - 3] : PCC[r - 3][c - 3]);
        };
        auto Hbb = [&](int r, int c) -> double {
          return (r < 3) ? ((c < 3) ? P11[r][c] : -P12[r][c - 3]) : ((c < 3) ? -P12[c][r - 3] : P22[r - 3][c - 3]);
        };
        // H_ab = Ja^T Jb = [[-P11, P12], [C^T B1, -C^T B2]] ; C^T B1 = P1C^T
        auto Hab = [&](int r, int c) -> double {
          return (r < 3) ? ((c < 3) ? -P11[r][c] : P12[r][c - 3]) : ((c < 3) ? P1C[c][r - 3] : -PC2[r - 3][c - 3]);
        };
        // store (or add) my 36 values into staging slot `slot` (panel order, row rotated by the slot: conflict-free drains)
        auto put = [&](auto value, int slot, bool add) {
          const int sw = slot >> 2;
          double* base = stg + slot * 36;
#pragma unroll
          for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              // sw <= 7: the rotation can only wrap for the last elements of the row (decided at compile time)
              const int idx = (pidx(r, c) + 7 < 36) ? pidx(r, c) + sw : rot36(pidx(r, c) + sw);
              const double v = value(r, c);
              base[idx] = add ? base[idx] + v : v;
            }
        };
        // the warp drains the staging tile element-major: coalesced RED / stores (4 lanes per 32-byte sector)
        auto drain = [&](double* base, bool reduce_always) {
          auto out = [&](int t, int el, double v) {
            if (reduce_always) { if (t >= 0) atomicAdd(base + 36 * (size_t)t + el, v); }
            else if (t >= 0) base[36 * (size_t)t + el] = v;
            else if (t <= -2) atomicAdd(base + 36 * (size_t)(-t - 2) + el, v);
          };
          // pass 1 -- lane l owns element l of every block (32 lanes = 8 whole sectors);
          // pass 2 -- elements 32..35, four lanes per block (one sector), eight blocks per instruction
#pragma unroll 8
          for (int blk = 0; blk < 32; ++blk) {
            const int t = sidx[blk];
            out(t, lane, stg[blk * 36 + rot36(lane + (blk >> 2))]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int blk = j * 8 + (lane >> 2), el = 32 + (lane & 3);
            const int t = sidx[blk];
            out(t, el, stg[blk * 36 + rot36(el + (blk >> 2))]);
          }
          __syncwarp();
        };
        auto VA = [&](int r, int c) { return rho1 * sa[r] * sa[c] * Haa(r, c); };
        auto VB = [&](int r, int c) { return rho1 * sb[r] * sb[c] * Hbb(r, c); };
        // ---- diagonal blocks, pass 1: begin-pose groups plus the end poses that hit one of them ----
        if (ra == 0) put(VA, lane, false);
        sidx[lane] = ra == 0 ? ta : -1;
        __syncwarp();
        for (int r = 1; __any_sync(full, ra >= r); ++r) {
          if (ra == r) put(VA, la, true);
          __syncwarp();
        }
        for (int r = 0; __any_sync(full, hit >= 0 && rh >= r); ++r) {
          if (hit >= 0 && rh == r) put(VB, hit, true);
          __syncwarp();
        }
        drain(p.Hdiag, true);
        // ---- pass 2: end poses that no begin pose of this tile matched (the first lane of a chain, far cross edges) ----
        if (any_b_left) {
          if (b_left && rb == 0) put(VB, lane, false);
          sidx[lane] = (b_left && rb == 0) ? tb : -1;
          __syncwarp();
          for (int r = 1; __any_sync(full, b_left && rb >= r); ++r) {
            if (b_left && rb == r) put(VB, lb, true);
            __syncwarp();
          }
          drain(p.Hdiag, true);
        }
        // off-diagonal blocks: (a,b) = H_ab, (b,a) = H_ab^T; slot >= 0: sole producer (store), <= -2: shared slot (RED)
        put([&](int r, int c) { return rho1 * sa[r] * sb[c] * Hab(r, c); }, lane, false);
        sidx[lane] = valid ? ct->slot_ab[lane] : -1;
        __syncwarp();
        drain(p.Hoff, false);
        put([&](int r, int c) { return rho1 * sb[r] * sa[c] * Hab(c, r); }, lane, false);
        sidx[lane] = valid ? ct->slot_ba[lane] : -1;
        __syncwarp();
        drain(p.Hoff, false);
      }
    }
    __syncwarp();
    stage = stage1;
    if (stage == 0) ++round;
    istage ^= 1;
    if (istage == 0) ++iround;
  }
  cost_acc = warp_sum(cost_acc);
  if (lane == 0 && cost_acc != 0.0) atomicAdd(&p.scalars->cost, cost_acc);
}

// --------------------------------------------------------------------------------------------
// Block-CSR 6x6 SpMV row: lane `r` (0..5) of a 6-lane group returns row r of
//   y_i = (Hdiag_i + diag(d_i)) x_i + sum_j Hoff_ij x_j .
// Panel layout => the 6 lanes of a group read 96 contiguous bytes per 128-bit load.
// kCoherent: gather x with ld.global.cg (x was written by other CTAs of the same launch).
// --------------------------------------------------------------------------------------------
// (T = float: the operator copy the multilevel preconditioner sweeps over, pgo_amg.cuh -- half the bytes; the arithmetic
// stays fp64)
template <typename T> struct Pair2;
template <> struct Pair2<double> { typedef double2 type; };
template <> struct Pair2<float> { typedef float2 type; };
template <typename T>
__device__ __forceinline__ double2 ldg_pair(const typename Pair2<T>::type* p) {
  const typename Pair2<T>::type v = __ldg(p);
  return make_double2((double)v.x, (double)v.y);
}
template <bool kCoherent, typename T = double>
__device__ __forceinline__ double bsr6_row(const T* __restrict__ Hdiag, const T* __restrict__ Hoff,
                                           const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                           const double* x, const double* __restrict__ d, int i, int r, bool with_diag = true) {
  typedef typename Pair2<T>::type T2;
  auto ldx = [&](const double* q) -> double2 {
    return kCoherent ? __ldcg(reinterpret_cast<const double2*>(q)) : __ldg(reinterpret_cast<const double2*>(q));
  };
  const T2* hd = reinterpret_cast<const T2*>(Hdiag + 36 * (size_t)i) + r;
  const double2 h0 = ldg_pair<T>(hd), h1 = ldg_pair<T>(hd + 6), h2 = ldg_pair<T>(hd + 12);
  const double* xi = x + 6 * (size_t)i;
  const double2 x0 = ldx(xi), x1 = ldx(xi + 2), x2 = ldx(xi + 4);
  double acc = h0.x * x0.x;
  acc = fma(h0.y, x0.y, acc); acc = fma(h1.x, x1.x, acc); acc = fma(h1.y, x1.y, acc);
  acc = fma(h2.x, x2.x, acc); acc = fma(h2.y, x2.y, acc);
  if (!with_diag) acc = 0.0;   // multi-GPU: the all-reduced diagonal is applied by rank 0 only
  if (d != nullptr && with_diag) {
    const double xr = (r == 0) ? x0.x : (r == 1) ? x0.y : (r == 2) ? x1.x : (r == 3) ? x1.y : (r == 4) ? x2.x : x2.y;
    acc = fma(__ldg(d + 6 * (size_t)i + r), xr, acc);
  }
  const int p0 = __ldg(row_ptr + i), p1 = __ldg(row_ptr + i + 1);
  double acc2 = 0.0;
  int p = p0;
  for (; p + 1 < p1; p += 2) {
    const int j0 = __ldg(col_idx + p), j1 = __ldg(col_idx + p + 1);
    const T2* ha = reinterpret_cast<const T2*>(Hoff + 36 * (size_t)p) + r;
    const T2* hb = ha + 18;
    const double2 a0 = ldg_pair<T>(ha), a1 = ldg_pair<T>(ha + 6), a2 = ldg_pair<T>(ha + 12);
    const double2 b0 = ldg_pair<T>(hb), b1 = ldg_pair<T>(hb + 6), b2 = ldg_pair<T>(hb + 12);
    const double* xa = x + 6 * (size_t)j0;
    const double* xb = x + 6 * (size_t)j1;
    const double2 u0 = ldx(xa), u1 = ldx(xa + 2), u2 = ldx(xa + 4);
    const double2 v0 = ldx(xb), v1 = ldx(xb + 2), v2 = ldx(xb + 4);
    acc = fma(a0.x, u0.x, acc); acc = fma(a0.y, u0.y, acc); acc = fma(a1.x, u1.x, acc);
    acc = fma(a1.y, u1.y, acc); acc = fma(a2.x, u2.x, acc); acc = fma(a2.y, u2.y, acc);
    acc2 = fma(b0.x, v0.x, acc2); acc2 = fma(b0.y, v0.y, acc2); acc2 = fma(b1.x, v1.x, acc2);
    acc2 = fma(b1.y, v1.y, acc2); acc2 = fma(b2.x, v2.x, acc2); acc2 = fma(b2.y, v2.y, acc2);
  }
  if (p < p1) {
    const int j0 = __ldg(col_idx + p);
    const T2* ha = reinterpret_cast<const T2*>(Hoff + 36 * (size_t)p) + r;
    const double2 a0 = ldg_pair<T>(ha), a1 = ldg_pair<T>(ha + 6), a2 = ldg_pair<T>(ha + 12);
    const double* xa = x + 6 * (size_t)j0;
    const double2 u0 = ldx(xa), u1 = ldx(xa + 2), u2 = ldx(xa + 4);
    acc = fma(a0.x, u0.x, acc); acc = fma(a0.y, u0.y, acc); acc = fma(a1.x, u1.x, acc);
    acc = fma(a1.y, u1.y, acc); acc = fma(a2.x, u2.x, acc); acc = fma(a2.y, u2.y, acc);
  }
  return acc + acc2;
}

struct BsrView {
  int n;                       // block rows
  const double* Hdiag;
  const double* Hoff;
  const int* row_ptr;
  const int* col_idx;
};

// y = (H + diag(d)) x ; grid-stride over groups of 5 rows per warp.  dot_part (optional): per-CTA partial of y . x
// in a fixed summation order (the CG step length of the stream-ordered PCG rides on the SpMV).
template <bool kDot>
__global__ void __launch_bounds__(256) spmv_kernel(const BsrView A, const double* __restrict__ x,
                                                   const double* __restrict__ d, double* __restrict__ y,
                                                   bool with_diag, double* __restrict__ dot_part = nullptr,
                                                   const int* __restrict__ skip = nullptr) {
  __shared__ double red[8];
  const bool idle = kDot && skip != nullptr && *skip != 0;   // stream-ordered PCG: converged, keep the vectors frozen
  const int lane = threadIdx.x & 31;
  const int grp = lane / 6, r = lane - grp * 6;
  const int warps_per_cta = blockDim.x >> 5;
  const int gw = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  const int nw = gridDim.x * warps_per_cta;
  double acc = 0.0;
  if (!idle) {
    for (int base = gw * kRowsPerWarp; base < A.n; base += nw * kRowsPerWarp) {
      const int i = base + grp;
      if (grp < kRowsPerWarp && i < A.n) {
        const double v = bsr6_row<false>(A.Hdiag, A.Hoff, A.row_ptr, A.col_idx, x, d, i, r, with_diag);
        y[6 * (size_t)i + r] = v;
        if (kDot) acc = fma(v, __ldg(x + 6 * (size_t)i + r), acc);
      }
    }
  }
  if (kDot) {
    acc = warp_sum(acc);
    if (lane == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < warps_per_cta; ++k) t += red[k];
      dot_part[blockIdx.x] = t;
    }
  }
}

// --------------------------------------------------------------------------------------------
// Grid barrier for the persistent PCG kernel (cooperative launch => all CTAs co-resident).
// Monotonic arrival counter; thread 0 of each CTA arrives and spins.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - epoch) < 0);
    __threadfence();
  }
  __syncthreads();
}

struct PcgParams {
  BsrView A;
  const double* d;        // [N][6] LM diagonal (added to H)
  const double* Minv;     // [N][36] row-major inverse of (Hdiag + diag(d)) blocks
  const double* b;        // [N][6]
  double* x; double* r; double* u; double* w; double* p; double* s;
  double* partials;       // [2 parities][3 slots][gridDim.x]
  unsigned int* barrier;
  DeviceScalars* scalars;
  int max_iterations;
  double tolerance;
  const LmState* lm;      // device-resident LM loop (done flag), or nullptr
};

__device__ __forceinline__ double cta_sum(double v, double* red) {
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

__device__ __forceinline__ double reduce_partials(const double* part, int n, double* bcast) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    double t = 0.0;
    for (int k = lane; k < n; k += 32) t += __ldcg(part + k);
    t = warp_sum(t);
    if (lane == 0) *bcast = t;
  }
  __syncthreads();
  const double v = *bcast;
  __syncthreads();
  return v;
}


// kCluster: the whole launch is ONE thread-block cluster (small graphs) and the two barriers per iteration are the
// hardware cluster barrier instead of the atomic-counter grid barrier.
template <bool kCluster>
__device__ __forceinline__ void pcg_sync(unsigned int* counter, unsigned int& epoch) {
  if constexpr (kCluster) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else grid_barrier(counter, epoch);
}

template <bool kCluster>
__global__ void __launch_bounds__(kPcgThreads) pcg_kernel(const PcgParams P) {
  __shared__ double red[kPcgThreads / 32];
  __shared__ double bcast;
  if (P.lm && P.lm->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / 6, r = lane - grp * 6;
  const bool lane_on = grp < kRowsPerWarp;
  const int warps_per_cta = kPcgThreads / 32;
  const int gw = blockIdx.x * warps_per_cta + warp;
  const int nw = gridDim.x * warps_per_cta;
  const int n = P.A.n;
  const int G = gridDim.x;
  unsigned int epoch = 0;
  const unsigned gmask = 0xffffffffu;

  // u_i = Minv_i r_i for the row this lane group owns; rnew is this lane's component of r_i
  auto precond = [&](int i, double rnew) -> double {
    const double2* mi = reinterpret_cast<const double2*>(P.Minv + 36 * (size_t)i + 6 * r);
    const double2 m0 = __ldg(mi), m1 = __ldg(mi + 1), m2 = __ldg(mi + 2);
    const int g0 = grp * 6;
    const double r0 = __shfl_sync(gmask, rnew, g0), r1 = __shfl_sync(gmask, rnew, g0 + 1), r2 = __shfl_sync(gmask, rnew, g0 + 2);
    const double r3 = __shfl_sync(gmask, rnew, g0 + 3), r4 = __shfl_sync(gmask, rnew, g0 + 4), r5 = __shfl_sync(gmask, rnew, g0 + 5);
    return m0.x * r0 + m0.y * r1 + m1.x * r2 + m1.y * r3 + m2.x * r4 + m2.y * r5;
  };

  // ---- init: x = 0, r = b, u = Minv r, p = s = 0 ; gamma = r.u ----
  double acc = 0.0;
  for (int base = gw * kRowsPerWarp; base < n; base += nw * kRowsPerWarp) {
    const int i = base + grp;
    const bool on = lane_on && i < n;
    const size_t k = 6 * (size_t)(on ? i : 0) + r;
    const double rv = on ? P.b[k] : 0.0;
    const double uv = precond(on ? i : 0, rv);
    if (on) { P.x[k] = 0.0; P.r[k] = rv; P.u[k] = uv; P.p[k] = 0.0; P.s[k] = 0.0; acc += rv * uv; }
  }
  acc = cta_sum(acc, red);
  if (threadIdx.x == 0) P.partials[0 * 3 * G + 0 * G + blockIdx.x] = acc;
  pcg_sync<kCluster>(P.barrier, epoch);
  const double gamma0 = reduce_partials(P.partials + 0, G, &bcast);
  double gamma = gamma0, gamma_old = 0.0, alpha = 0.0, beta = 0.0;
  int iter = 0;
  int flag = 0;
  const double stop = P.tolerance * P.tolerance * gamma0;

  if (gamma0 > 0.0) {
    for (;;) {
      const int par = iter & 1;
      double* part = P.partials + (size_t)par * 3 * G;
      // ---- phase B: w = A u, delta = w.u ----
      acc = 0.0;
      for (int base = gw * kRowsPerWarp; base < n; base += nw * kRowsPerWarp) {
        const int i = base + grp;
        if (lane_on && i < n) {
          const double wv = bsr6_row<true>(P.A.Hdiag, P.A.Hoff, P.A.row_ptr, P.A.col_idx, P.u, P.d, i, r);
          const size_t k = 6 * (size_t)i + r;
          P.w[k] = wv;
          acc += wv * P.u[k];
        }
      }
      acc = cta_sum(acc, red);
      if (threadIdx.x == 0) part[1 * G + blockIdx.x] = acc;
      pcg_sync<kCluster>(P.barrier, epoch);
      const double delta = reduce_partials(part + 1 * G, G, &bcast);
      // ---- scalars (identical on every thread) ----
      if (iter == 0) { beta = 0.0; alpha = gamma / delta; }
      else { beta = gamma / gamma_old; alpha = gamma / (delta - beta * gamma / alpha); }
      if (!(alpha > 0.0) || !isfinite(alpha)) { flag = 2; break; }
      ++iter;
      // ---- phase A: p = u + beta p, s = w + beta s, x += alpha p, r -= alpha s, u = Minv r, gamma' = r.u ----
      acc = 0.0;
      for (int base = gw * kRowsPerWarp; base < n; base += nw * kRowsPerWarp) {
        const int i = base + grp;
        const bool on = lane_on && i < n;
        const size_t k = 6 * (size_t)(on ? i : 0) + r;
        double rv = 0.0;
        if (on) {
          const double pv = P.u[k] + beta * P.p[k];
          const double sv = P.w[k] + beta * P.s[k];
          P.p[k] = pv; P.s[k] = sv;
          P.x[k] += alpha * pv;
          rv = P.r[k] - alpha * sv;
          P.r[k] = rv;
        }
        const double uv = precond(on ? i : 0, rv);
        if (on) { P.u[k] = uv; acc += rv * uv; }
      }
      acc = cta_sum(acc, red);
      double* partn = P.partials + (size_t)(iter & 1) * 3 * G;
      if (threadIdx.x == 0) partn[0 * G + blockIdx.x] = acc;
      pcg_sync<kCluster>(P.barrier, epoch);
      gamma_old = gamma;
      gamma = reduce_partials(partn + 0 * G, G, &bcast);
      if (gamma <= stop) { flag = 0; break; }
      if (iter >= P.max_iterations) { flag = 1; break; }
    }
  }

  // ---- epilogue: x^T b, x^T (H + D) x, x^T D x for the model cost change ----
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int base = gw * kRowsPerWarp; base < n; base += nw * kRowsPerWarp) {
    const int i = base + grp;
    if (lane_on && i < n) {
      const double ax = bsr6_row<true>(P.A.Hdiag, P.A.Hoff, P.A.row_ptr, P.A.col_idx, P.x, P.d, i, r);
      const size_t k = 6 * (size_t)i + r;
      const double xv = P.x[k];
      a0 += xv * P.b[k]; a1 += xv * ax; a2 += xv * xv * P.d[k];
    }
  }
  a0 = cta_sum(a0, red); a1 = cta_sum(a1, red); a2 = cta_sum(a2, red);
  double* parte = P.partials + (size_t)((iter + 1) & 1) * 3 * G;
  if (threadIdx.x == 0) { parte[0 * G + blockIdx.x] = a0; parte[1 * G + blockIdx.x] = a1; parte[2 * G + blockIdx.x] = a2; }
  pcg_sync<kCluster>(P.barrier, epoch);
  const double e0 = reduce_partials(parte + 0 * G, G, &bcast);
  const double e1 = reduce_partials(parte + 1 * G, G, &bcast);
  const double e2 = reduce_partials(parte + 2 * G, G, &bcast);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    P.scalars->xtb = e0;
    P.scalars->xtAx = e1;
    P.scalars->xtDx = e2;
    P.scalars->pcg_gamma0 = gamma0;
    P.scalars->pcg_gamma = gamma;
    P.scalars->pcg_iterations = iter;
    P.scalars->pcg_flag = flag;
  }
}

// --------------------------------------------------------------------------------------------
// Per-pose helpers
// --------------------------------------------------------------------------------------------
// Jacobi scaling (trust_region_minimizer.cc: 1 / (1 + sqrt(column norm^2))) from the diagonal of the
// unscaled Hessian; 0 for constant / unused poses so that their columns vanish from the problem.
// mask [n][6]: 1 for the components of variable parameter blocks, 0 for constant ones (a pose may have only p or only q constant)
__global__ void jacobi_scale_kernel(int n, const double* __restrict__ Hdiag, const double* __restrict__ mask,
                                    int use_scaling, double* __restrict__ scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double s = 0.0;
    if (mask[6 * (size_t)i + c] > 0.0) s = use_scaling ? 1.0 / (1.0 + sqrt(Hdiag[36 * (size_t)i + pidx(c, c)])) : 1.0;
    scale[6 * (size_t)i + c] = s;
  }
}

// LevenbergMarquardtStrategy::ComputeStep prologue: diagonal (unless reused), D = diagonal / radius,
// and the block-Jacobi preconditioner Minv = (Hdiag + D)^-1 by Cholesky.
__global__ void lm_prepare_kernel(int n, const double* __restrict__ Hdiag, const unsigned char* __restrict__ active,
                                  int mode /*0: new diagonal, 1: reuse diagonal, 2: dlm given*/, double min_diag,
                                  double max_diag, double radius, double* __restrict__ diagonal,
                                  double* __restrict__ dlm, double* __restrict__ Minv, const LmState* lm = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (lm) {                       // device-resident LM loop: the radius and the reuse decision live on the device
    if (lm->done) return;
    mode = lm->reuse_diagonal ? 1 : 0;
    radius = lm->radius;
  }
  double A[6][6];
  double dd[6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) A[r][c] = Hdiag[36 * (size_t)i + pidx(r, c)];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    if (mode == 2) {
      dd[c] = dlm[6 * (size_t)i + c];
    } else {
      double v;
      if (mode == 1) v = diagonal[6 * (size_t)i + c];
      else { v = fmin(fmax(A[c][c], min_diag), max_diag); diagonal[6 * (size_t)i + c] = v; }
      dd[c] = v / radius;
      dlm[6 * (size_t)i + c] = dd[c];
    }
    A[c][c] += dd[c];
  }
  double* out = Minv + 36 * (size_t)i;
  if (!active[i]) {
#pragma unroll
    for (int k = 0; k < 36; ++k) out[k] = 0.0;
    return;
  }
  // Cholesky A = L L^T (in place, lower), then Minv = L^-T L^-1
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
  // Linv (lower)
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
      out[r * 6 + c] = ok ? t : ((r == c) ? 1.0 / (dd[r] > 0 ? dd[r] : 1.0) : 0.0);
    }
}

// x_cand = Plus(x, sign * y .* scale) for active poses; accumulates |x - x_cand|^2 and |x_cand|^2.
__global__ void __launch_bounds__(256) plus_kernel(int n, const double* __restrict__ x, const double* __restrict__ y,
                                                   const double* __restrict__ scale, const unsigned char* __restrict__ active,
                                                   double sign, double* __restrict__ out, DeviceScalars* scalars,
                                                   LmState* lm = nullptr, double* __restrict__ zero_hdiag = nullptr,
                                                   double* __restrict__ zero_grad = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (lm) {
    if (lm->done) return;
    if (i == 0) lm->t_solved = lm_globaltimer();      // the linear solve ends where this kernel starts
  }
  double sn = 0.0, xn = 0.0;
  if (i < n) {
    // the system the speculative linearisation is about to accumulate into starts from zero (folded memsets)
    if (zero_hdiag) {
      double2* z = reinterpret_cast<double2*>(zero_hdiag + 36 * (size_t)i);
#pragma unroll
      for (int k = 0; k < 18; ++k) z[k] = make_double2(0.0, 0.0);
    }
    if (zero_grad) {
      double2* z = reinterpret_cast<double2*>(zero_grad + 6 * (size_t)i);
#pragma unroll
      for (int k = 0; k < 3; ++k) z[k] = make_double2(0.0, 0.0);
    }
    double xi[7], d[6], o[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) xi[k] = x[8 * (size_t)i + k];
    if (active[i]) {
#pragma unroll
      for (int k = 0; k < 6; ++k) d[k] = sign * y[6 * (size_t)i + k] * scale[6 * (size_t)i + k];
      pose_plus(xi, d, o);
      // |x|^2 runs over the VARIABLE parameter blocks (p and q separately: either may be constant)
      const bool pvar = scale[6 * (size_t)i] > 0.0, qvar = scale[6 * (size_t)i + 3] > 0.0;
#pragma unroll
      for (int k = 0; k < 7; ++k) { const double t = xi[k] - o[k]; sn += t * t; if (k < 3 ? pvar : qvar) xn += o[k] * o[k]; }
    } else {
#pragma unroll
      for (int k = 0; k < 7; ++k) o[k] = xi[k];
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) out[8 * (size_t)i + k] = o[k];
    out[8 * (size_t)i + 7] = 0.0;
  }
  sn = warp_sum(sn); xn = warp_sum(xn);
  if ((threadIdx.x & 31) == 0 && (sn != 0.0 || xn != 0.0)) { atomicAdd(&scalars->step_norm2, sn); atomicAdd(&scalars->x_norm2, xn); }
}

// gradient norms as Ceres reports them: |x - Plus(x, -g)| with g = unscaled gradient = g_s / scale.
__global__ void __launch_bounds__(256) gradient_norm_kernel(int n, const double* __restrict__ x, const double* __restrict__ gs,
                                                            const double* __restrict__ scale, const unsigned char* __restrict__ active,
                                                            double* __restrict__ g_unscaled, DeviceScalars* scalars,
                                                            const LmState* lm = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (lm && lm->done) return;
  double mx = 0.0, l2 = 0.0;
  if (i < n) {
    double g[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double s = scale[6 * (size_t)i + k];
      g[k] = (active[i] && s > 0.0) ? gs[6 * (size_t)i + k] / s : 0.0;
      if (g_unscaled) g_unscaled[6 * (size_t)i + k] = g[k];
    }
    if (active[i]) {
      double xi[7], d[6], o[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) xi[k] = x[8 * (size_t)i + k];
#pragma unroll
      for (int k = 0; k < 6; ++k) d[k] = -g[k];
      pose_plus(xi, d, o);
#pragma unroll
      for (int k = 0; k < 7; ++k) { const double t = fabs(xi[k] - o[k]); mx = fmax(mx, t); l2 += t * t; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  l2 = warp_sum(l2);
  if ((threadIdx.x & 31) == 0 && (mx != 0.0 || l2 != 0.0)) {
    atomicMax(&scalars->gmax_bits, (unsigned long long)__double_as_longlong(mx));
    atomicAdd(&scalars->gnorm2, l2);
  }
}

// x_norm^2 over active poses (iteration zero)
__global__ void __launch_bounds__(256) xnorm_kernel(int n, const double* __restrict__ x, const unsigned char* __restrict__ active,
                                                    const double* __restrict__ mask, DeviceScalars* scalars) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double xn = 0.0;
  if (i < n && active[i]) {
    const bool pvar = mask[6 * (size_t)i] > 0.0, qvar = mask[6 * (size_t)i + 3] > 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) { const double t = x[8 * (size_t)i + k]; if (k < 3 ? pvar : qvar) xn += t * t; }
  }
  xn = warp_sum(xn);
  if ((threadIdx.x & 31) == 0 && xn != 0.0) atomicAdd(&scalars->x_norm2, xn);
}

}  // namespace pgo
