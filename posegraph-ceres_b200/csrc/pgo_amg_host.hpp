// pgo_amg_host.hpp -- host-side (pure C++, no CUDA) construction of the aggregation hierarchy behind the multilevel
// preconditioner of the damped normal equations (pgo_amg.cuh) and of the row partition of the multi-GPU path.
//
// Why: the reference solves (J^T J + D) y = J^T r exactly (SPARSE_NORMAL_CHOLESKY, REF/test/
// pose_graph_ceres_plus_finial.cpp:536).  On mesh-like graphs (sphere, grids, dense random loops) an exact factor fills
// in and block-Jacobi PCG needs 10^3..10^4 iterations per LM step.  A pose-graph Hessian is a connection Laplacian whose
// near-null space is the 6 rigid-body motions of a connected piece of the graph (EXACT null vectors of J^T J at any
// linearisation point: the cost is invariant under a global rigid motion), so the classical elasticity recipe applies:
// aggregate poses into patches, give every patch its 6 rigid-body modes as coarse unknowns (the coarse level is again a
// 6x6 block system), form the Galerkin operators P^T A P, and use a V-cycle with block-Jacobi smoothing as the PCG
// preconditioner.  Everything the V-cycle does is a 6x6 block SpMV or a per-pose vector operation: HBM-bound, like the
// rest of the path.
//
// Multi-GPU: poses (block rows of H) are partitioned into contiguous index ranges, one per rank; aggregates never cross
// a rank boundary, so every level keeps the same owner-computes structure and only halo values move.  Levels that have
// become small are replicated on every rank.  Every rank builds the GLOBAL hierarchy redundantly from the same input
// (deterministic, no communication) and then keeps its own slice.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace pgo {

struct AmgHostParams {
  double theta = 0.3;        // geometric strength: neighbour j of i is strong when 1/|p_i-p_j|^2 >= theta * max over both rows
  int coarsest_max = 16;     // stop coarsening at <= this many nodes: that level is inverted densely (amg_host_params sizes it)
  int max_levels = 12;
  int replicate_max = 32768; // multi-GPU: a level with at most this many nodes lives on every rank
  double stall_ratio = 0.85; // give up coarsening when a level keeps more than this share of its nodes
};

// The parameters every entry point uses for a graph of n_poses poses.  The first level of at most coarsest_max nodes is
// solved exactly (dense inverse, once per LM step): N/16 keeps that inversion a small share of an LM step on small graphs,
// 512 nodes (a 3072 x 3072 fp64 inverse, 75 MB) bounds it on large ones.
inline AmgHostParams amg_host_params(int n_poses) {
  AmgHostParams prm;
  prm.coarsest_max = std::min(512, std::max(16, n_poses / 16));
  if (const char* e = getenv("PGO_AMG_DENSE_MAX")) prm.coarsest_max = std::min(512, std::max(1, atoi(e)));
  if (const char* e = getenv("PGO_AMG_THETA")) prm.theta = atof(e);
  if (const char* e = getenv("PGO_AMG_REPLICATE_MAX")) prm.replicate_max = atoi(e);
  return prm;
}

// One level of the global hierarchy.  Nodes are numbered rank by rank: rank r owns [off[r], off[r+1]).
struct AmgGlobalLevel {
  int n = 0;
  std::vector<int> off;                 // [world + 1]
  std::vector<int> row_ptr, col_idx;    // symmetric off-diagonal block pattern, columns ascending
  std::vector<double> pos;              // [n][3] setup-time positions (strength measure only)
  std::vector<int> agg;                 // [n] -> node of the next level, -1 = not a variable; empty on the last level
  bool replicated = false;
};

// Greedy aggregation restricted to one rank's range [lo, hi): (1) a node whose strong neighbourhood is untouched becomes
// a root and takes it, (2) leftovers join the aggregate of their strongest aggregated neighbour, (3) what is left forms
// aggregates of its own connected leftovers.  Returns the number of aggregates created (ids start at `next`).
inline int amg_aggregate_range(const AmgGlobalLevel& L, int lo, int hi, const std::vector<unsigned char>& variable, double theta,
                               int next, std::vector<int>& agg) {
  const int first = next;
  auto d2 = [&](int i, int j) {
    const double dx = L.pos[3 * (size_t)i] - L.pos[3 * (size_t)j], dy = L.pos[3 * (size_t)i + 1] - L.pos[3 * (size_t)j + 1],
                 dz = L.pos[3 * (size_t)i + 2] - L.pos[3 * (size_t)j + 2];
    return dx * dx + dy * dy + dz * dz + 1e-12;
  };
  // row maxima of the strength 1/d^2 over ALL neighbours (also those on other ranks: the measure is symmetric)
  std::vector<double> rmax((size_t)(hi - lo), 0.0);
  for (int i = lo; i < hi; ++i) {
    double m = 0.0;
    for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) m = std::max(m, 1.0 / d2(i, L.col_idx[p]));
    rmax[i - lo] = m;
  }
  auto rowmax = [&](int j) -> double {
    if (j >= lo && j < hi) return rmax[j - lo];
    double m = 0.0;
    for (int p = L.row_ptr[j]; p < L.row_ptr[j + 1]; ++p) m = std::max(m, 1.0 / d2(j, L.col_idx[p]));
    return m;
  };
  auto strong = [&](int i, int j) -> bool {   // j in the same range, both variables
    if (j < lo || j >= hi || !variable[j]) return false;
    return 1.0 / d2(i, j) >= theta * std::max(rmax[i - lo], rowmax(j));
  };
  // pass 1
  for (int i = lo; i < hi; ++i) {
    if (!variable[i] || agg[i] >= 0) continue;
    bool free_nb = true;
    for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1] && free_nb; ++p) {
      const int j = L.col_idx[p];
      if (strong(i, j) && agg[j] >= 0) free_nb = false;
    }
    if (!free_nb) continue;
    agg[i] = next;
    for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) {
      const int j = L.col_idx[p];
      if (strong(i, j)) agg[j] = next;
    }
    ++next;
  }
  // pass 2: join the strongest aggregated neighbour (decided on the pass-1 state so the result does not depend on order)
  std::vector<int> join((size_t)(hi - lo), -1);
  for (int i = lo; i < hi; ++i) {
    if (!variable[i] || agg[i] >= 0) continue;
    double best = 0.0;
    for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) {
      const int j = L.col_idx[p];
      if (j < lo || j >= hi || !variable[j] || agg[j] < 0) continue;
      const double w = 1.0 / d2(i, j);
      if (w > best && strong(i, j)) { best = w; join[i - lo] = agg[j]; }
    }
  }
  for (int i = lo; i < hi; ++i) if (join[i - lo] >= 0) agg[i] = join[i - lo];
  // pass 3: leftovers (no aggregated strong neighbour): grow aggregates over ANY same-rank neighbours still free
  for (int i = lo; i < hi; ++i) {
    if (!variable[i] || agg[i] >= 0) continue;
    agg[i] = next;
    for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) {
      const int j = L.col_idx[p];
      if (j >= lo && j < hi && variable[j] && agg[j] < 0) agg[j] = next;
    }
    ++next;
  }
  return next - first;
}

// Global hierarchy from the level-0 pattern.  `variable[i]` = pose i is an unknown (used by an edge, not constant).
// off0: [world+1] contiguous ownership ranges of level 0.
inline void amg_build_global(int N, int world, const std::vector<int>& off0, const unsigned char* variable0, const int* row_ptr0,
                             const int* col_idx0, const double* pos0 /* [N][3] */, const AmgHostParams& prm,
                             std::vector<AmgGlobalLevel>* out) {
  out->clear();
  out->emplace_back();
  {
    AmgGlobalLevel& L = out->back();
    L.n = N; L.off = off0;
    L.row_ptr.assign(row_ptr0, row_ptr0 + N + 1);
    L.col_idx.assign(col_idx0, col_idx0 + row_ptr0[N]);
    L.pos.assign(pos0, pos0 + 3 * (size_t)N);
    L.replicated = false;   // level 0 always follows the partition (world == 1: the single rank owns everything)
  }
  std::vector<unsigned char> variable(variable0, variable0 + N);
  for (int lvl = 0; lvl + 1 < prm.max_levels; ++lvl) {
    AmgGlobalLevel& L = (*out)[lvl];
    int n_var = 0;
    for (int i = 0; i < L.n; ++i) n_var += variable[i] ? 1 : 0;
    if (n_var <= prm.coarsest_max) break;
    std::vector<int> agg((size_t)L.n, -1);
    std::vector<int> coff(world + 1, 0);
    int next = 0;
    if (L.replicated) {
      // every rank holds the whole level: aggregates may span the former rank ranges
      next = amg_aggregate_range(L, 0, L.n, variable, prm.theta, 0, agg);
      for (int r = 1; r <= world; ++r) coff[r] = next;
    } else {
      for (int r = 0; r < world; ++r) {
        coff[r] = next;
        next += amg_aggregate_range(L, L.off[r], L.off[r + 1], variable, prm.theta, next, agg);
      }
      coff[world] = next;
    }
    const int nc = next;
    if (nc == 0 || (double)nc > prm.stall_ratio * (double)n_var) break;   // coarsening stalled: this is the last level
    L.agg = agg;
    AmgGlobalLevel C;
    C.n = nc; C.off = coff;
    C.replicated = world > 1 && (L.replicated || nc <= prm.replicate_max);
    // centroids
    C.pos.assign(3 * (size_t)nc, 0.0);
    std::vector<int> cnt((size_t)nc, 0);
    for (int i = 0; i < L.n; ++i) {
      const int I = agg[i];
      if (I < 0) continue;
      cnt[I]++;
      for (int k = 0; k < 3; ++k) C.pos[3 * (size_t)I + k] += L.pos[3 * (size_t)i + k];
    }
    for (int I = 0; I < nc; ++I) for (int k = 0; k < 3; ++k) C.pos[3 * (size_t)I + k] /= std::max(cnt[I], 1);
    // coarse pattern: unique (I, J), I != J, over the fine blocks
    std::vector<int> start((size_t)nc + 1, 0);
    for (int i = 0; i < L.n; ++i) {
      const int I = agg[i];
      if (I < 0) continue;
      for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) {
        const int J = agg[L.col_idx[p]];
        if (J >= 0 && J != I) start[I + 1]++;
      }
    }
    for (int I = 0; I < nc; ++I) start[I + 1] += start[I];
    std::vector<int> cols((size_t)start[nc]);
    {
      std::vector<int> fill(start.begin(), start.end() - 1);
      for (int i = 0; i < L.n; ++i) {
        const int I = agg[i];
        if (I < 0) continue;
        for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) {
          const int J = agg[L.col_idx[p]];
          if (J >= 0 && J != I) cols[fill[I]++] = J;
        }
      }
    }
    C.row_ptr.assign((size_t)nc + 1, 0);
    C.col_idx.reserve(cols.size() / 2 + 16);
    for (int I = 0; I < nc; ++I) {
      int* b = cols.data() + start[I];
      int* e = cols.data() + start[I + 1];
      std::sort(b, e);
      e = std::unique(b, e);
      C.col_idx.insert(C.col_idx.end(), b, e);
      C.row_ptr[I + 1] = (int)C.col_idx.size();
    }
    out->push_back(std::move(C));
    variable.assign((size_t)nc, 1);
  }
  // once a level is replicated every coarser one is too
  for (size_t l = 1; l < out->size(); ++l) if ((*out)[l - 1].replicated) (*out)[l].replicated = true;
}

// ----------------------------------------------------------------------------------------------------------------
// One rank's slice of a level.  Local numbering: the rows this rank stores first (distributed: its owned nodes in
// global order; replicated: all nodes, local id == global id), then the halo columns in ascending global id (which
// groups them by owner rank).
// ----------------------------------------------------------------------------------------------------------------
struct AmgLocalLevel {
  bool replicated = false;
  int n_own = 0, n_halo = 0;
  int g0 = 0;                               // global id of local node 0
  std::vector<int> halo_gid;                // [n_halo]
  std::vector<int> row_ptr, col_idx;        // local block-CSR of the stored rows (off-diagonal part); level 0: the graph's own
  // ---- coarsening (empty on the last level) ----
  std::vector<int> agg;                     // [n_own + n_halo] -> LOCAL id on the next level, -1 = none
  int c_row0 = 0, c_row1 = 0;               // coarse rows (next level, local ids) this rank computes: [c_row0, c_row1)
  std::vector<int> mem_ptr, mem_idx;        // members (fine local rows) of each computed coarse row
  std::vector<int> gal_ptr;                 // per computed coarse block: first the diagonal blocks of rows c_row0..c_row1-1,
                                            // then their off-diagonal slots in CSR order
  std::vector<int> gal_row, gal_slot;       // contributions: fine row i and fine slot p (>= 0), or -1 = the diagonal block of i
  // ---- halo exchange plan (distributed levels of a multi-rank run) ----
  std::vector<int> nbr;                     // neighbour ranks, ascending
  std::vector<int> send_ptr, send_idx;      // per neighbour: the owned local ids it needs, ascending
  std::vector<int> recv_ptr;                // per neighbour: its slice [recv_ptr[k], recv_ptr[k+1]) of the halo region
  // ---- all-gather plan: this level is replicated but is filled from rank-owned row ranges ----
  std::vector<int> gather_off;              // [world+1] node ranges per rank (empty: nothing to gather)
  std::vector<int> gather_slot_off;         // [world+1] the matching off-diagonal slot ranges
};

inline int amg_owner_of(const std::vector<int>& off, int gid) {
  return (int)(std::upper_bound(off.begin(), off.end(), gid) - off.begin()) - 1;
}

// Slice the global hierarchy for `rank`.  `level0` describes the rank's level-0 rows as the graph stores them: row_ptr /
// col_idx with local column ids (owned g - off[rank], halo n_own + k), halo_gid, and -- multi-rank -- the exchange plan
// (nbr / send_ptr / send_idx / recv_ptr), which the graph derives from its EDGES: the level-0 halo also holds constant
// poses, which have no block in the pattern but whose values the residuals need.  world == 1: the global pattern.
inline void amg_localize(const std::vector<AmgGlobalLevel>& G, int rank, int world, const AmgLocalLevel& level0,
                         std::vector<AmgLocalLevel>* out) {
  const int nl = (int)G.size();
  out->assign(nl, AmgLocalLevel());
  (*out)[0] = level0;
  for (int l = 0; l < nl; ++l) {
    const AmgGlobalLevel& g = G[l];
    AmgLocalLevel& L = (*out)[l];
    L.replicated = g.replicated;
    if (g.replicated) { L.n_own = g.n; L.g0 = 0; }
    else { L.n_own = g.off[rank + 1] - g.off[rank]; L.g0 = g.off[rank]; }
  }
  (*out)[0].n_halo = (int)(*out)[0].halo_gid.size();

  for (int l = 0; l < nl; ++l) {
    const AmgGlobalLevel& g = G[l];
    AmgLocalLevel& L = (*out)[l];
    auto gid_of = [&](int loc) { return loc < L.n_own ? L.g0 + loc : L.halo_gid[loc - L.n_own]; };
    // ---- halo exchange plan ----
    if (!L.replicated && world > 1 && l > 0) {
      L.recv_ptr.clear(); L.nbr.clear();
      for (int k = 0; k < L.n_halo; ++k) {
        const int o = amg_owner_of(g.off, L.halo_gid[k]);
        if (L.nbr.empty() || L.nbr.back() != o) { L.nbr.push_back(o); L.recv_ptr.push_back(k); }
      }
      L.recv_ptr.push_back(L.n_halo);
      // what each neighbour needs from me: my owned nodes with a column owned by it (the pattern is symmetric, so the
      // neighbour's halo holds exactly these, in ascending global id)
      std::vector<std::vector<int>> need(L.nbr.size());
      for (int i = 0; i < L.n_own; ++i) {
        int last = -1;
        for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) {
          const int j = L.col_idx[p];
          if (j < L.n_own) continue;
          const int o = amg_owner_of(g.off, L.halo_gid[j - L.n_own]);
          if (o == last) continue;
          const int k = (int)(std::lower_bound(L.nbr.begin(), L.nbr.end(), o) - L.nbr.begin());
          if (need[k].empty() || need[k].back() != i) need[k].push_back(i);
          last = o;
        }
      }
      L.send_ptr.assign(1, 0);
      L.send_idx.clear();
      for (size_t k = 0; k < L.nbr.size(); ++k) {
        std::sort(need[k].begin(), need[k].end());
        need[k].erase(std::unique(need[k].begin(), need[k].end()), need[k].end());
        L.send_idx.insert(L.send_idx.end(), need[k].begin(), need[k].end());
        L.send_ptr.push_back((int)L.send_idx.size());
      }
    }
    if (l + 1 >= nl) break;
    // ---- coarsening maps towards level l + 1 ----
    const AmgGlobalLevel& gc = G[l + 1];
    AmgLocalLevel& C = (*out)[l + 1];
    const int n_loc = L.n_own + L.n_halo;
    L.agg.assign(n_loc, -1);
    std::vector<int> chalo;
    for (int i = 0; i < n_loc; ++i) {
      const int Gc = g.agg[gid_of(i)];
      if (Gc < 0) continue;
      if (C.replicated) L.agg[i] = Gc;
      else if (Gc >= gc.off[rank] && Gc < gc.off[rank + 1]) L.agg[i] = Gc - gc.off[rank];
      else { L.agg[i] = -2 - Gc; chalo.push_back(Gc); }   // resolved below
    }
    if (!C.replicated) {
      std::sort(chalo.begin(), chalo.end());
      chalo.erase(std::unique(chalo.begin(), chalo.end()), chalo.end());
      C.halo_gid = chalo;
      C.n_halo = (int)chalo.size();
      for (int i = 0; i < n_loc; ++i)
        if (L.agg[i] <= -2) {
          const int Gc = -L.agg[i] - 2;
          L.agg[i] = C.n_own + (int)(std::lower_bound(chalo.begin(), chalo.end(), Gc) - chalo.begin());
        }
    }
    // rows of the coarse level computed here
    if (!C.replicated) { L.c_row0 = 0; L.c_row1 = C.n_own; }
    else if (!L.replicated) { L.c_row0 = gc.off[rank]; L.c_row1 = gc.off[rank + 1]; }
    else { L.c_row0 = 0; L.c_row1 = gc.n; }
    if (C.replicated && !L.replicated && world > 1) {
      C.gather_off = gc.off;
      C.gather_slot_off.resize(world + 1);
      for (int r = 0; r <= world; ++r) C.gather_slot_off[r] = gc.row_ptr[gc.off[r]];
    }
    // local CSR of the coarse level from the global pattern
    {
      const int r0 = C.replicated ? 0 : gc.off[rank], r1 = C.replicated ? gc.n : gc.off[rank + 1];
      C.row_ptr.assign((size_t)(r1 - r0) + 1, 0);
      C.col_idx.clear();
      C.col_idx.reserve((size_t)(gc.row_ptr[r1] - gc.row_ptr[r0]));
      std::vector<int> tmp;
      for (int I = r0; I < r1; ++I) {
        tmp.clear();
        for (int p = gc.row_ptr[I]; p < gc.row_ptr[I + 1]; ++p) {
          const int J = gc.col_idx[p];
          int loc;
          if (C.replicated) loc = J;
          else if (J >= gc.off[rank] && J < gc.off[rank + 1]) loc = J - gc.off[rank];
          else {
            const auto it = std::lower_bound(C.halo_gid.begin(), C.halo_gid.end(), J);
            // every column of an owned coarse row is the aggregate of a column of an owned fine row, hence in the halo
            loc = C.n_own + (int)(it - C.halo_gid.begin());
          }
          tmp.push_back(loc);
        }
        std::sort(tmp.begin(), tmp.end());
        C.col_idx.insert(C.col_idx.end(), tmp.begin(), tmp.end());
        C.row_ptr[I - r0 + 1] = (int)C.col_idx.size();
      }
    }
    // members and Galerkin gather lists of the computed coarse rows
    const int ncomp = L.c_row1 - L.c_row0;
    L.mem_ptr.assign((size_t)ncomp + 1, 0);
    for (int i = 0; i < L.n_own; ++i) { const int I = L.agg[i]; if (I >= 0) L.mem_ptr[I - L.c_row0 + 1]++; }
    for (int k = 0; k < ncomp; ++k) L.mem_ptr[k + 1] += L.mem_ptr[k];
    L.mem_idx.resize((size_t)L.mem_ptr[ncomp]);
    {
      std::vector<int> fill(L.mem_ptr.begin(), L.mem_ptr.end() - 1);
      for (int i = 0; i < L.n_own; ++i) { const int I = L.agg[i]; if (I >= 0) L.mem_idx[fill[I - L.c_row0]++] = i; }
    }
    const int slot0 = C.row_ptr[L.c_row0], slot1 = C.row_ptr[L.c_row1];
    const int nblk = ncomp + (slot1 - slot0);
    auto block_of = [&](int I, int J) -> int {   // computed coarse block id of (I, J)
      if (I == J) return I - L.c_row0;
      const int* b = C.col_idx.data() + C.row_ptr[I];
      const int* e = C.col_idx.data() + C.row_ptr[I + 1];
      const int* it = std::lower_bound(b, e, J);
      return ncomp + (int)(it - C.col_idx.data()) - slot0;
    };
    L.gal_ptr.assign((size_t)nblk + 1, 0);
    for (int i = 0; i < L.n_own; ++i) {
      const int I = L.agg[i];
      if (I < 0) continue;
      L.gal_ptr[block_of(I, I) + 1]++;
      for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) {
        const int J = L.agg[L.col_idx[p]];
        if (J >= 0) L.gal_ptr[block_of(I, J) + 1]++;
      }
    }
    for (int k = 0; k < nblk; ++k) L.gal_ptr[k + 1] += L.gal_ptr[k];
    L.gal_row.resize((size_t)L.gal_ptr[nblk]);
    L.gal_slot.resize((size_t)L.gal_ptr[nblk]);
    {
      std::vector<int> fill(L.gal_ptr.begin(), L.gal_ptr.end() - 1);
      for (int i = 0; i < L.n_own; ++i) {
        const int I = L.agg[i];
        if (I < 0) continue;
        int q = fill[block_of(I, I)]++;
        L.gal_row[q] = i; L.gal_slot[q] = -1;
        for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) {
          const int J = L.agg[L.col_idx[p]];
          if (J < 0) continue;
          q = fill[block_of(I, J)]++;
          L.gal_row[q] = i; L.gal_slot[q] = p;
        }
      }
    }
  }
}

}  // namespace pgo
