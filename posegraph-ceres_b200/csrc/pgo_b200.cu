// pgo_b200.cu -- host side of libpgo_b200.so: graph upload + block structure analysis, the
// Levenberg-Marquardt driver (a restatement of ceres::internal::TrustRegionMinimizer +
// LevenbergMarquardtStrategy as the reference configures them,
// REF/test/pose_graph_ceres_plus_finial.cpp:500-514) and the C-ABI of include/pgo_b200.h.
// All numerical work happens in the kernels of pgo_kernels.cuh; there is no CPU fallback.
#include "../../include/pgo_b200.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "pgo_kernels.cuh"
#include "pgo_pool.cuh"

using namespace pgo;

static thread_local std::string g_last_error;
static int set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return set_error(PGO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,        \
                                            cudaGetErrorString(_e), __FILE__, __LINE__);         \
  } while (0)
#define NCCL_TRY(expr)                                                                           \
  do {                                                                                           \
    ncclResult_t _e = (expr);                                                                    \
    if (_e != ncclSuccess) return set_error(PGO_ERR_NCCL, "%s failed: %s (%s:%d)", #expr,        \
                                            ncclGetErrorString(_e), __FILE__, __LINE__);         \
  } while (0)
#define PGO_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != PGO_OK) return _rc; \
  } while (0)

#include "pgo_level_chol.cuh"
#include "pgo_candidates.cuh"

static double wall_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

namespace pgo { struct PcgMultiState; }

struct pgo_graph {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int N = 0, E = 0, T = 0;
  bool identity_info = true;
  bool has_dup_blocks = false;
  long long nnz_off = 0;
  int num_sms = 0;
  int pcg_max_ctas = 0;
  std::vector<unsigned char> active_h;
  std::vector<int> row_ptr_h, col_idx_h;
  // device
  double *poses = nullptr, *poses_cand = nullptr, *poses_snap = nullptr;
  double *scale = nullptr, *scale_eval = nullptr;
  EdgeCoreTile* core = nullptr;
  EdgeInfoTile* info = nullptr;
  double *Hdiag = nullptr, *Hoff = nullptr;
  double *Hdiag_alt = nullptr, *Hoff_alt = nullptr, *grad_alt = nullptr;   // second system for the speculative linearisation
  int *row_ptr = nullptr, *col_idx = nullptr;
  double *grad = nullptr, *grad_unscaled = nullptr;
  double *diagonal = nullptr, *dlm = nullptr, *Minv = nullptr;
  double *vx = nullptr, *vr = nullptr, *vu = nullptr, *vw = nullptr, *vp = nullptr, *vs = nullptr, *vb = nullptr;
  unsigned char* active = nullptr;
  DeviceScalars* scalars = nullptr;
  DeviceScalars* scalars_h = nullptr;  // pinned
  double* partials = nullptr;
  unsigned int* barrier = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // level-scheduled Cholesky preconditioner (optional)
  LevelChol* chol = nullptr;
  long long launches = 0;
  double setup_s = 0.0;
  std::vector<std::pair<void*, size_t>> blocks;   // device memory borrowed from the per-device pool
  std::vector<double> pose_stage;                 // host staging for the [N][7] <-> [N][8] pose layouts
  // symbolic factor analysis started on a helper thread by the one-shot entry point (overlaps upload + iteration zero)
  std::future<int> sym_future;
  std::unique_ptr<LevelCholSymbolic> sym;
  int sym_solver_type = -1;
  struct pgo::PcgMultiState* pcgm_state = nullptr;    // stream-ordered (multi-GPU) PCG
  struct pgo::PcgMultiState* pcgm_state_h = nullptr;  // pinned
  double *pcgm_part0 = nullptr, *pcgm_part1 = nullptr;
};

// Host-only structure analysis shared by pgo_graph_create and pgo_analyze_structure: variable poses (used by an
// edge and not constant) and the block-CSR pattern of the off-diagonal part of J^T J.
struct HostPattern {
  std::vector<unsigned char> active;
  std::vector<int> row_ptr, col_idx;
  std::vector<int> half_slot;             // [2E]: slot of block (a,b) at 2e, of (b,a) at 2e+1; -1 = none; <= -2: -(slot)-2, shared by several edges
  bool has_dup = false;
};
static void build_pattern(int N, int E, const int* edge_ids, const unsigned char* pose_const, HostPattern* out) {
  out->active.assign(N, 0);
  for (int e = 0; e < E; ++e) { out->active[edge_ids[2 * e]] = 1; out->active[edge_ids[2 * e + 1]] = 1; }
  if (pose_const) for (int i = 0; i < N; ++i) if (pose_const[i]) out->active[i] = 0;
  // bucket the half-edges (row -> col) by row with a counting sort, then order each (short) row by column
  struct Half { int col; int idx; };
  std::vector<int> start(N + 1, 0);
  size_t n_half = 0;
  for (int e = 0; e < E; ++e) {
    const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
    if (out->active[a] && out->active[b]) { start[a + 1]++; start[b + 1]++; n_half += 2; }
  }
  for (int i = 0; i < N; ++i) start[i + 1] += start[i];
  std::vector<Half> halves(n_half);
  {
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int e = 0; e < E; ++e) {
      const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
      if (out->active[a] && out->active[b]) { halves[fill[a]++] = {b, 2 * e}; halves[fill[b]++] = {a, 2 * e + 1}; }
    }
  }
  out->half_slot.assign(2 * (size_t)E, -1);
  out->row_ptr.assign(N + 1, 0);
  out->col_idx.clear();
  out->col_idx.reserve(n_half);
  for (int i = 0; i < N; ++i) {
    Half* hb = halves.data() + start[i];
    Half* he = halves.data() + start[i + 1];
    if (he - hb > 1) std::sort(hb, he, [](const Half& x, const Half& y) { return x.col < y.col || (x.col == y.col && x.idx < y.idx); });
    for (Half* k = hb; k < he;) {
      Half* k2 = k + 1;
      while (k2 < he && k2->col == k->col) ++k2;
      const int slot = (int)out->col_idx.size();
      out->col_idx.push_back(k->col);
      out->row_ptr[i + 1]++;
      const bool dup = k2 - k > 1;
      if (dup) out->has_dup = true;
      for (Half* q = k; q < k2; ++q) out->half_slot[q->idx] = dup ? -slot - 2 : slot;
      k = k2;
    }
  }
  for (int i = 0; i < N; ++i) out->row_ptr[i + 1] += out->row_ptr[i];
}

static int check_edges(int n_poses, int n_edges, const int* edge_ids) {
  for (int e = 0; e < n_edges; ++e) {
    const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
    if (a < 0 || a >= n_poses || b < 0 || b >= n_poses || a == b)
      return set_error(PGO_ERR_INVALID_ARGUMENT, "edge %d has invalid endpoints (%d, %d)", e, a, b);
  }
  return PGO_OK;
}

extern "C" int pgo_analyze_structure(int n_poses, int n_edges, const int* edge_ids, const unsigned char* pose_const,
                                     double max_fill_ratio, pgo_structure_info* info) {
  if (!info || n_poses <= 0 || n_edges < 0 || (n_edges > 0 && !edge_ids))
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_analyze_structure: null or empty input");
  PGO_TRY(check_edges(n_poses, n_edges, edge_ids));
  const double t0 = wall_s();
  std::memset(info, 0, sizeof *info);
  HostPattern pat;
  build_pattern(n_poses, n_edges, edge_ids, pose_const, &pat);
  const double t1 = wall_s();
  for (unsigned char a : pat.active) info->variable_poses += a;
  info->hessian_blocks = (long long)pat.col_idx.size() + n_poses;
  LevelCholSymbolic S;
  // max_fill_ratio > 0: the "cheap factor only" analysis PGO_LINEAR_AUTO runs (run_symbolic), with this fill limit
  const bool cheap_only = max_fill_ratio > 0.0;
  PGO_TRY(level_chol_symbolic(&S, n_poses, pat.active.data(), pat.row_ptr.data(), pat.col_idx.data(),
                              cheap_only ? max_fill_ratio : 1e30, cheap_only ? 64 : 8192, cheap_only ? 16 : (1 << 30)));
  if (getenv("PGO_PROFILE_HOST")) fprintf(stderr, "[pgo analyze] pattern %.1f us, symbolic %.1f us\n", 1e6 * (t1 - t0), 1e6 * (wall_s() - t1));
  info->factor_usable = S.usable ? 1 : 0;
  if (S.usable) {
    info->factor_blocks = S.n_slots + S.n_nodes;
    info->factor_levels = S.num_levels;
    info->factor_max_degree = S.max_degree;
    info->factor_tasks = (long long)S.tasks.size();
    // invariants of the schedule (checked here so that CPU tests cover the host logic):
    // nodes of one level are pairwise non-adjacent in the filled graph, every slot row is eliminated later
    std::vector<int> pos(n_poses, -1), lvl(n_poses, -1);
    for (int l = 0; l < S.num_levels; ++l)
      for (int k = S.level_ptr[l]; k < S.level_ptr[l + 1]; ++k) { pos[S.nodes[k].x] = k; lvl[S.nodes[k].x] = l; }
    for (int k = 0; k < S.n_nodes; ++k)
      for (int p = S.nodes[k].y; p < S.nodes[k].z; ++p) {
        const int u = S.col_row[p];
        if (pos[u] <= k || lvl[u] <= lvl[S.nodes[k].x]) return set_error(PGO_ERR_NUMERICAL, "level schedule violates elimination order");
      }
  }
  info->analysis_seconds = wall_s() - t0;
  return PGO_OK;
}

extern "C" const char* pgo_last_error(void) { return g_last_error.c_str(); }
extern "C" int pgo_abi_version(void) { return PGO_B200_ABI_VERSION; }
extern "C" int pgo_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" void pgo_default_options(pgo_solver_options* o) {
  o->max_num_iterations = 1000;  // REF/test/pose_graph_ceres_plus_finial.cpp:504
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
  o->loss_type = PGO_LOSS_HUBER;  // :463
  o->loss_a = 1.0;
  o->linear_solver_type = PGO_LINEAR_AUTO;
  o->pcg_max_iterations = 20000;
  o->pcg_tolerance = 1e-10;
  o->pcg_num_ctas = 0;
  o->direct_residual_accept = 1e-8;
  o->verbose = 0;
}

template <typename Tp>
static int dev_alloc(pgo_graph* g, Tp** p, size_t count) {
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(Tp);
  CUDA_TRY(pool_alloc(g->device, reinterpret_cast<void**>(p), bytes));
  g->blocks.emplace_back(static_cast<void*>(*p), bytes);
  return PGO_OK;
}

extern "C" void pgo_graph_destroy(pgo_graph* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->sym_future.valid()) g->sym_future.get();
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->own_stream && g->own_stream != g->stream) cudaStreamSynchronize(g->own_stream);
  if (g->comm) ncclCommDestroy(g->comm);
  if (g->chol) level_chol_destroy(g->chol, g->device);
  for (auto& blk : g->blocks) pool_free(g->device, blk.first, blk.second);
  pool_pinned_release(g->device, g->scalars_h);
  pool_pinned_release(g->device, g->pcgm_state_h);
  pool_event_release(g->device, g->ev0);
  pool_event_release(g->device, g->ev1);
  pool_event_release(g->device, g->ev2);
  pool_event_release(g->device, g->ev3);
  pool_stream_release(g->device, g->own_stream);
  delete g;
}

extern "C" void pgo_release_cached_memory(int device) { pool_release(device); }

static thread_local int g_symbolic_hint = -1;   // set by pgo_solve_pose_graph around its pgo_graph_create call
static int run_symbolic(const pgo_graph* g, int t, LevelCholSymbolic* S);

extern "C" int pgo_graph_create(pgo_graph** out, int device, int n_poses, int n_edges, const double* poses,
                                const int* edge_ids, const double* edge_meas, const double* edge_sqrt_info,
                                const unsigned char* pose_const) {
  if (!out || n_poses <= 0 || n_edges < 0 || !poses || (n_edges > 0 && (!edge_ids || !edge_meas)))
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_create: null or empty input");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(PGO_ERR_NO_DEVICE, "pgo_graph_create: no CUDA device available (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return set_error(PGO_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  PGO_TRY(check_edges(n_poses, n_edges, edge_ids));
  const double t0 = wall_s();
  static const bool prof = getenv("PGO_PROFILE_HOST") != nullptr;
  double tp = t0;
  auto lap = [&](const char* what) { if (prof) { const double t = wall_s(); fprintf(stderr, "[pgo create] %-28s %8.1f us\n", what, 1e6 * (t - tp)); tp = t; } };
  CUDA_TRY(cudaSetDevice(device));
  pgo_graph* g = new pgo_graph();
  g->device = device;
  g->N = n_poses; g->E = n_edges; g->T = (n_edges + kTile - 1) / kTile;
  const int N = g->N, E = g->E, T = g->T;
  auto fail = [&](int rc) { pgo_graph_destroy(g); return rc; };
#define G_TRY(expr) do { int _rc = (expr); if (_rc != PGO_OK) return fail(_rc); } while (0)
#define GC_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(set_error(PGO_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e))); } while (0)
  GC_TRY(pool_stream(device, &g->own_stream));
  g->stream = g->own_stream;
  GC_TRY(pool_event(device, &g->ev0));
  GC_TRY(pool_event(device, &g->ev1));
  GC_TRY(pool_event(device, &g->ev2));
  GC_TRY(pool_event(device, &g->ev3));
  g->num_sms = pool_num_sms(device);

  // ---- which poses are variables, block-CSR pattern of the off-diagonal part ----
  lap("stream/events");
  HostPattern pat;
  build_pattern(N, E, edge_ids, pose_const, &pat);
  lap("block-CSR pattern");
  g->active_h.swap(pat.active);
  g->row_ptr_h.swap(pat.row_ptr);
  g->col_idx_h.swap(pat.col_idx);
  g->nnz_off = (long long)g->col_idx_h.size();
  g->has_dup_blocks = pat.has_dup;
  if (g_symbolic_hint == PGO_LINEAR_AUTO || g_symbolic_hint == PGO_LINEAR_PCG_LEVEL_CHOLESKY) {
    // the pattern is final: analyse the elimination order on a helper thread while this thread packs and uploads
    g->sym.reset(new LevelCholSymbolic());
    g->sym_solver_type = g_symbolic_hint;
    pgo_graph* gp = g;
    const int t = g_symbolic_hint;
    g->sym_future = std::async(std::launch::async, [gp, t]() { return run_symbolic(gp, t, gp->sym.get()); });
  }

  // ---- identity information? (after the helper thread is off: the symbolic analysis is the longer leg) ----
  g->identity_info = true;
  if (edge_sqrt_info) {
    static const double eye[36] = {1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1};
    for (int e = 0; e < E && g->identity_info; ++e)
      if (std::memcmp(edge_sqrt_info + 36 * (size_t)e, eye, sizeof eye) != 0) {
        // memcmp also flags -0.0; confirm numerically
        for (int k = 0; k < 36; ++k) if (edge_sqrt_info[36 * (size_t)e + k] != eye[k]) { g->identity_info = false; break; }
      }
  }
  lap("identity scan");
  // ---- edge tiles (field-major, one warp per tile) ----
  std::vector<EdgeCoreTile> core_h(std::max(T, 1));
  std::memset(core_h.data(), 0, core_h.size() * sizeof(EdgeCoreTile));
  std::vector<EdgeInfoTile> info_h;
  if (!g->identity_info) { info_h.resize(std::max(T, 1)); std::memset(info_h.data(), 0, info_h.size() * sizeof(EdgeInfoTile)); }
  for (int e = 0; e < E; ++e) {
    EdgeCoreTile& t = core_h[e / kTile];
    const int l = e % kTile;
    const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
    t.a[l] = a; t.b[l] = b;
    t.slot_ab[l] = pat.half_slot[2 * (size_t)e];
    t.slot_ba[l] = pat.half_slot[2 * (size_t)e + 1];
    for (int k = 0; k < 7; ++k) t.meas[k][l] = edge_meas[7 * (size_t)e + k];
    if (!g->identity_info) for (int k = 0; k < 36; ++k) info_h[e / kTile].S[k][l] = edge_sqrt_info[36 * (size_t)e + k];
  }
  for (int e = E; e < T * kTile; ++e) { core_h[e / kTile].slot_ab[e % kTile] = -1; core_h[e / kTile].slot_ba[e % kTile] = -1; }

  lap("edge tiles");
  // ---- device allocations + uploads ----
  G_TRY(dev_alloc(g, &g->poses, (size_t)N * 8));
  G_TRY(dev_alloc(g, &g->poses_cand, (size_t)N * 8));
  G_TRY(dev_alloc(g, &g->poses_snap, (size_t)N * 8));
  G_TRY(dev_alloc(g, &g->scale, (size_t)N * 6));
  G_TRY(dev_alloc(g, &g->scale_eval, (size_t)N * 6));
  G_TRY(dev_alloc(g, &g->core, (size_t)std::max(T, 1)));
  if (!g->identity_info) G_TRY(dev_alloc(g, &g->info, (size_t)std::max(T, 1)));
  G_TRY(dev_alloc(g, &g->Hdiag, (size_t)N * 36));
  G_TRY(dev_alloc(g, &g->Hoff, (size_t)g->nnz_off * 36));
  G_TRY(dev_alloc(g, &g->row_ptr, (size_t)N + 1));
  G_TRY(dev_alloc(g, &g->col_idx, (size_t)g->nnz_off));
  G_TRY(dev_alloc(g, &g->grad, (size_t)N * 6));
  G_TRY(dev_alloc(g, &g->grad_unscaled, (size_t)N * 6));
  G_TRY(dev_alloc(g, &g->diagonal, (size_t)N * 6));
  G_TRY(dev_alloc(g, &g->dlm, (size_t)N * 6));
  G_TRY(dev_alloc(g, &g->Minv, (size_t)N * 36));
  for (double** v : {&g->vx, &g->vr, &g->vu, &g->vw, &g->vp, &g->vs, &g->vb}) G_TRY(dev_alloc(g, v, (size_t)N * 6));
  G_TRY(dev_alloc(g, &g->active, (size_t)N));
  G_TRY(dev_alloc(g, &g->scalars, 1));
  static_assert(sizeof(DeviceScalars) <= kPinnedBytes, "pinned scalars");
  GC_TRY(pool_pinned(device, reinterpret_cast<void**>(&g->scalars_h)));
  G_TRY(dev_alloc(g, &g->barrier, 4));

  lap("device allocations");
  GC_TRY(cudaMemcpyAsync(g->core, core_h.data(), (size_t)std::max(T, 1) * sizeof(EdgeCoreTile), cudaMemcpyHostToDevice, g->stream));
  if (!g->identity_info) GC_TRY(cudaMemcpyAsync(g->info, info_h.data(), (size_t)std::max(T, 1) * sizeof(EdgeInfoTile), cudaMemcpyHostToDevice, g->stream));
  GC_TRY(cudaMemcpyAsync(g->row_ptr, g->row_ptr_h.data(), ((size_t)N + 1) * sizeof(int), cudaMemcpyHostToDevice, g->stream));
  if (g->nnz_off) GC_TRY(cudaMemcpyAsync(g->col_idx, g->col_idx_h.data(), (size_t)g->nnz_off * sizeof(int), cudaMemcpyHostToDevice, g->stream));
  GC_TRY(cudaMemcpyAsync(g->active, g->active_h.data(), (size_t)N, cudaMemcpyHostToDevice, g->stream));
  {
    std::vector<double> se((size_t)N * 6);
    for (int i = 0; i < N; ++i) for (int k = 0; k < 6; ++k) se[6 * (size_t)i + k] = g->active_h[i] ? 1.0 : 0.0;
    GC_TRY(cudaMemcpyAsync(g->scale_eval, se.data(), se.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    GC_TRY(cudaMemcpyAsync(g->scale, se.data(), se.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  }
  // (pageable sources: cudaMemcpyAsync returns once they are staged, so the host vectors may go; the stream orders the rest)
  GC_TRY(cudaMemsetAsync(g->Hoff, 0, std::max<size_t>((size_t)g->nnz_off * 36, 1) * sizeof(double), g->stream));

  lap("uploads + memset");
  // persistent PCG grid: all CTAs must be co-resident (cooperative launch)
  static int per_sm = 0;
  if (per_sm == 0) GC_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_kernel<false>, kPcgThreads, 0));
  g->pcg_max_ctas = std::max(1, std::min(per_sm, 4) * g->num_sms);
  G_TRY(dev_alloc(g, &g->partials, (size_t)2 * 3 * g->pcg_max_ctas));

  *out = g;
  int rc = pgo_graph_set_poses(g, poses);
  if (rc != PGO_OK) { *out = nullptr; return fail(rc); }
  lap("poses");
  g->setup_s = wall_s() - t0;
  return PGO_OK;
#undef G_TRY
#undef GC_TRY
}

extern "C" int pgo_graph_set_stream(pgo_graph* g, void* cuda_stream) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  g->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : g->own_stream;
  return PGO_OK;
}

extern "C" int pgo_graph_num_poses(const pgo_graph* g) { return g ? g->N : 0; }
extern "C" int pgo_graph_num_edges(const pgo_graph* g) { return g ? g->E : 0; }

extern "C" int pgo_graph_set_poses(pgo_graph* g, const double* poses) {
  if (!g || !poses) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_set_poses: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  // [N][7] host -> [N][8] device: pad on the host, one contiguous copy (row-pitched DMA of 56-byte rows is slow)
  g->pose_stage.resize((size_t)g->N * 8);
  for (int i = 0; i < g->N; ++i) {
    std::memcpy(&g->pose_stage[8 * (size_t)i], poses + 7 * (size_t)i, 7 * sizeof(double));
    g->pose_stage[8 * (size_t)i + 7] = 0.0;
  }
  CUDA_TRY(cudaMemcpyAsync(g->poses, g->pose_stage.data(), (size_t)g->N * 8 * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PGO_OK;
}

extern "C" int pgo_graph_get_poses(pgo_graph* g, double* poses) {
  if (!g || !poses) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_get_poses: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  g->pose_stage.resize((size_t)g->N * 8);
  CUDA_TRY(cudaMemcpyAsync(g->pose_stage.data(), g->poses, (size_t)g->N * 8 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  for (int i = 0; i < g->N; ++i) std::memcpy(poses + 7 * (size_t)i, &g->pose_stage[8 * (size_t)i], 7 * sizeof(double));
  return PGO_OK;
}

extern "C" int pgo_graph_snapshot_poses(pgo_graph* g) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  CUDA_TRY(cudaMemcpyAsync(g->poses_snap, g->poses, (size_t)g->N * 8 * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PGO_OK;
}
extern "C" int pgo_graph_restore_poses(pgo_graph* g) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  CUDA_TRY(cudaMemcpyAsync(g->poses, g->poses_snap, (size_t)g->N * 8 * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PGO_OK;
}

// ------------------------------------------------------------------------------------------------
// NCCL plumbing: edges sharded across ranks, poses replicated.
// ------------------------------------------------------------------------------------------------
extern "C" int pgo_nccl_unique_id(unsigned char unique_id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_TRY(ncclGetUniqueId(&id));
  std::memcpy(unique_id, &id, 128);
  return PGO_OK;
}

extern "C" int pgo_graph_init_comm(pgo_graph* g, const unsigned char unique_id[128], int rank, int world_size) {
  if (!g || !unique_id || world_size < 1 || rank < 0 || rank >= world_size)
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_init_comm: bad arguments");
  CUDA_TRY(cudaSetDevice(g->device));
  if (world_size == 1) { g->rank = 0; g->world = 1; return PGO_OK; }
  ncclUniqueId id;
  std::memcpy(&id, unique_id, 128);
  NCCL_TRY(ncclCommInitRank(&g->comm, world_size, id, rank));
  g->rank = rank; g->world = world_size;
  // A pose is a variable when ANY rank's shard uses it: make the active set (and the unit scaling derived from it) global.
  NCCL_TRY(ncclAllReduce(g->active, g->active, (size_t)g->N, ncclUint8, ncclMax, g->comm, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->active_h.data(), g->active, (size_t)g->N, cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  {
    std::vector<double> se((size_t)g->N * 6);
    for (int i = 0; i < g->N; ++i) for (int k = 0; k < 6; ++k) se[6 * (size_t)i + k] = g->active_h[i] ? 1.0 : 0.0;
    CUDA_TRY(cudaMemcpyAsync(g->scale_eval, se.data(), se.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(g->scale, se.data(), se.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
  }
  return PGO_OK;
}

static int allreduce_sum(pgo_graph* g, double* buf, size_t count) {
  if (g->world <= 1) return PGO_OK;
  NCCL_TRY(ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, g->comm, g->stream));
  return PGO_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel launch helpers
// ------------------------------------------------------------------------------------------------
template <bool kIdent, int kMode, int kMinBlocks>
static int launch_linearize_t(pgo_graph* g, const LinParams& p) {
  constexpr int smem = lin_smem_bytes<kIdent>();
  auto kern = linearize_kernel<kIdent, kMode, kMinBlocks>;
  static bool attr_set = false;
  if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
  const int ctas = std::max(1, std::min((g->T + kLinWarps - 1) / kLinWarps, kMinBlocks * g->num_sms));
  kern<<<ctas, kLinWarps * 32, smem, g->stream>>>(p);
  g->launches++;
  CUDA_TRY(cudaGetLastError());
  return PGO_OK;
}

static int launch_linearize(pgo_graph* g, int mode, const double* poses, const double* scale, int loss_type,
                            double loss_a, double* res_out = nullptr, double* jac_out = nullptr) {
  LinParams p;
  p.n_edges = g->E; p.n_tiles = g->T; p.core = g->core; p.info = g->info; p.poses = poses; p.scale = scale;
  p.Hdiag = g->Hdiag; p.Hoff = g->Hoff; p.grad = g->grad; p.scalars = g->scalars;
  p.loss_type = loss_type; p.loss_a = loss_a; p.res_out = res_out; p.jac_out = jac_out;
  if (g->E == 0) return PGO_OK;
  static const int occ = getenv("PGO_LIN_OCC") ? atoi(getenv("PGO_LIN_OCC")) : 2;   // CTAs per SM of the full kernel (tuning knob)
  if (g->identity_info) {
    if (mode == kLinFull) return occ >= 3 ? launch_linearize_t<true, kLinFull, 3>(g, p) : launch_linearize_t<true, kLinFull, 2>(g, p);
    return launch_linearize_t<true, kLinEval, 2>(g, p);
  }
  if (mode == kLinFull) return occ >= 3 ? launch_linearize_t<false, kLinFull, 3>(g, p) : launch_linearize_t<false, kLinFull, 2>(g, p);
  return launch_linearize_t<false, kLinEval, 2>(g, p);
}

static int zero_system(pgo_graph* g, bool hessian) {
  if (hessian) {
    CUDA_TRY(cudaMemsetAsync(g->Hdiag, 0, (size_t)g->N * 36 * sizeof(double), g->stream));
    if (g->has_dup_blocks && g->nnz_off) CUDA_TRY(cudaMemsetAsync(g->Hoff, 0, (size_t)g->nnz_off * 36 * sizeof(double), g->stream));
  }
  CUDA_TRY(cudaMemsetAsync(g->grad, 0, (size_t)g->N * 6 * sizeof(double), g->stream));
  return PGO_OK;
}

static int zero_scalars(pgo_graph* g) {
  CUDA_TRY(cudaMemsetAsync(g->scalars, 0, sizeof(DeviceScalars), g->stream));
  return PGO_OK;
}
static int fetch_scalars(pgo_graph* g) {
  // every rank must take the same LM decisions: rank 0's scalars are authoritative
  if (g->world > 1) NCCL_TRY(ncclBroadcast(g->scalars, g->scalars, sizeof(DeviceScalars), ncclUint8, 0, g->comm, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->scalars_h, g->scalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PGO_OK;
}

// full linearization at `poses` with column scaling `scale`: H, g (all-reduced across ranks)
static int linearize_full(pgo_graph* g, const double* poses, const double* scale, int loss_type, double loss_a) {
  PGO_TRY(zero_system(g, true));
  PGO_TRY(launch_linearize(g, kLinFull, poses, scale, loss_type, loss_a));
  if (g->world > 1) {
    PGO_TRY(allreduce_sum(g, g->Hdiag, (size_t)g->N * 36));
    PGO_TRY(allreduce_sum(g, g->grad, (size_t)g->N * 6));
    PGO_TRY(allreduce_sum(g, &g->scalars->cost, 1));
  }
  return PGO_OK;
}

// The LM loop linearises speculatively at every candidate point into the alternate system (H, g) while the current one
// stays intact: an accepted step keeps the swap, a rejected one swaps back.  One stream synchronisation per iteration.
static void swap_system(pgo_graph* g) {
  std::swap(g->Hdiag, g->Hdiag_alt);
  std::swap(g->Hoff, g->Hoff_alt);
  std::swap(g->grad, g->grad_alt);
}

static BsrView bsr_view(const pgo_graph* g) {
  BsrView A;
  A.n = g->N; A.Hdiag = g->Hdiag; A.Hoff = g->Hoff; A.row_ptr = g->row_ptr; A.col_idx = g->col_idx;
  return A;
}

constexpr int kStreamPcgMinPoses = 200000;
constexpr int kClusterPcgMaxPoses = 400;   // measured: at 2500 poses the 16-SM cluster is already 2.5x slower than the full grid

static int pcg_grid(const pgo_graph* g, const pgo_solver_options* o) {
  int want = (g->N + (kPcgThreads / 32) * kRowsPerWarp - 1) / ((kPcgThreads / 32) * kRowsPerWarp);
  if (o && o->pcg_num_ctas > 0) want = o->pcg_num_ctas;
  return std::max(1, std::min(want, g->pcg_max_ctas));
}

// Single-GPU persistent PCG: (H + diag(dlm)) x = b, x in g->vx. Results land in g->scalars.
static int launch_pcg(pgo_graph* g, const pgo_solver_options* o, const double* b) {
  PcgParams P;
  P.A = bsr_view(g);
  P.d = g->dlm; P.Minv = g->Minv; P.b = b;
  P.x = g->vx; P.r = g->vr; P.u = g->vu; P.w = g->vw; P.p = g->vp; P.s = g->vs;
  P.partials = g->partials; P.barrier = g->barrier; P.scalars = g->scalars;
  P.max_iterations = o->pcg_max_iterations; P.tolerance = o->pcg_tolerance;
  CUDA_TRY(cudaMemsetAsync(g->barrier, 0, 4 * sizeof(unsigned int), g->stream));
  // small graphs: one 16-CTA cluster, hardware barrier (an iteration is a few microseconds, the barrier dominates)
  static int cluster_ctas = -1;
  if (cluster_ctas < 0) {
    cluster_ctas = 0;
    cudaFuncSetAttribute(pcg_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaGetLastError();
    for (int cs : {16, 8}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs); cfg.blockDim = dim3(kPcgThreads);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, pcg_kernel<true>, &cfg) == cudaSuccess && ncl >= 1) { cluster_ctas = cs; break; }
      cudaGetLastError();
    }
  }
  if (cluster_ctas > 0 && g->N <= kClusterPcgMaxPoses && o->pcg_num_ctas <= 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_ctas); cfg.blockDim = dim3(kPcgThreads); cfg.stream = g->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, pcg_kernel<true>, P));
  } else {
    void* args[] = {&P};
    const int grid = pcg_grid(g, o);
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)pcg_kernel<false>, dim3(grid), dim3(kPcgThreads), args, 0, g->stream));
  }
  g->launches++;
  return PGO_OK;
}

#include "pgo_pcg_multi.cuh"

// ------------------------------------------------------------------------------------------------
// C-ABI: evaluate / linearize / hessian / spmv / linear solve
// ------------------------------------------------------------------------------------------------
extern "C" int pgo_graph_evaluate(pgo_graph* g, int loss_type, double loss_a, double* cost, double* residuals,
                                  double* gradient, double* jacobians) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  double *res_d = nullptr, *jac_d = nullptr;
  const size_t res_bytes = std::max<size_t>((size_t)g->E * 6, 1) * sizeof(double), jac_bytes = std::max<size_t>((size_t)g->E * 72, 1) * sizeof(double);
  CUDA_TRY(pool_alloc(g->device, reinterpret_cast<void**>(&res_d), res_bytes));
  CUDA_TRY(pool_alloc(g->device, reinterpret_cast<void**>(&jac_d), jac_bytes));
  int rc = PGO_OK;
  do {
    if ((rc = zero_system(g, false)) != PGO_OK) break;
    if ((rc = zero_scalars(g)) != PGO_OK) break;
    if ((rc = launch_linearize(g, kLinEval, g->poses, g->scale_eval, loss_type, loss_a, res_d, jac_d)) != PGO_OK) break;
    if (g->world > 1) {
      if ((rc = allreduce_sum(g, g->grad, (size_t)g->N * 6)) != PGO_OK) break;
      if ((rc = allreduce_sum(g, &g->scalars->cost, 1)) != PGO_OK) break;
    }
    if ((rc = fetch_scalars(g)) != PGO_OK) break;
    if (cost) *cost = g->scalars_h->cost;
    cudaError_t ce = cudaSuccess;
    if (residuals && g->E) ce = cudaMemcpy(residuals, res_d, (size_t)g->E * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && jacobians && g->E) ce = cudaMemcpy(jacobians, jac_d, (size_t)g->E * 72 * sizeof(double), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && gradient) ce = cudaMemcpy(gradient, g->grad, (size_t)g->N * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) rc = set_error(PGO_ERR_CUDA, "evaluate copy-back failed: %s", cudaGetErrorString(ce));
  } while (0);
  cudaStreamSynchronize(g->stream);
  pool_free(g->device, res_d, res_bytes); pool_free(g->device, jac_d, jac_bytes);
  return rc;
}

extern "C" int pgo_graph_linearize(pgo_graph* g, int loss_type, double loss_a, const double* scale, double* cost,
                                   float* elapsed_ms) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  if (scale) CUDA_TRY(cudaMemcpyAsync(g->scale, scale, (size_t)g->N * 6 * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  else CUDA_TRY(cudaMemcpyAsync(g->scale, g->scale_eval, (size_t)g->N * 6 * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
  PGO_TRY(zero_system(g, true));
  PGO_TRY(zero_scalars(g));
  CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
  PGO_TRY(launch_linearize(g, kLinFull, g->poses, g->scale, loss_type, loss_a));
  CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
  if (g->world > 1) {
    PGO_TRY(allreduce_sum(g, g->Hdiag, (size_t)g->N * 36));
    PGO_TRY(allreduce_sum(g, g->grad, (size_t)g->N * 6));
    PGO_TRY(allreduce_sum(g, &g->scalars->cost, 1));
  }
  PGO_TRY(fetch_scalars(g));
  if (cost) *cost = g->scalars_h->cost;
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, g->ev0, g->ev1));
  return PGO_OK;
}

extern "C" int pgo_graph_get_hessian(pgo_graph* g, long long* nnzb, int* row_ptr, int* col_idx, double* values,
                                     double* gradient) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  const long long total = g->nnz_off + g->N;
  if (nnzb) *nnzb = total;
  if (gradient) CUDA_TRY(cudaMemcpy(gradient, g->grad, (size_t)g->N * 6 * sizeof(double), cudaMemcpyDeviceToHost));
  if (!row_ptr || !col_idx || !values) return PGO_OK;
  std::vector<double> hd((size_t)g->N * 36), ho((size_t)std::max<long long>(g->nnz_off, 1) * 36);
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  CUDA_TRY(cudaMemcpy(hd.data(), g->Hdiag, hd.size() * sizeof(double), cudaMemcpyDeviceToHost));
  if (g->nnz_off) CUDA_TRY(cudaMemcpy(ho.data(), g->Hoff, (size_t)g->nnz_off * 36 * sizeof(double), cudaMemcpyDeviceToHost));
  long long k = 0;
  for (int i = 0; i < g->N; ++i) {
    row_ptr[i] = (int)k;
    col_idx[k] = i;
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) values[36 * k + r * 6 + c] = hd[36 * (size_t)i + pidx(r, c)];
    ++k;
    for (int p = g->row_ptr_h[i]; p < g->row_ptr_h[i + 1]; ++p) {
      col_idx[k] = g->col_idx_h[p];
      for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) values[36 * k + r * 6 + c] = ho[36 * (size_t)p + pidx(r, c)];
      ++k;
    }
  }
  row_ptr[g->N] = (int)k;
  return PGO_OK;
}

extern "C" int pgo_graph_spmv(pgo_graph* g, const double* x, const double* d, double* y, int repeats, float* elapsed_ms) {
  if (!g || !x || !y) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_spmv: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  const size_t nv = (size_t)g->N * 6 * sizeof(double);
  CUDA_TRY(cudaMemcpyAsync(g->vu, x, nv, cudaMemcpyHostToDevice, g->stream));
  if (d) CUDA_TRY(cudaMemcpyAsync(g->dlm, d, nv, cudaMemcpyHostToDevice, g->stream));
  const int warps = (g->N + kRowsPerWarp - 1) / kRowsPerWarp;
  const int ctas = std::max(1, std::min((warps + 7) / 8, 8 * g->num_sms));
  CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
  for (int k = 0; k < std::max(repeats, 1); ++k) {
    spmv_kernel<false><<<ctas, 256, 0, g->stream>>>(bsr_view(g), g->vu, d ? g->dlm : nullptr, g->vw, g->rank == 0);
    g->launches++;
  }
  CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
  CUDA_TRY(cudaGetLastError());
  if (g->world > 1) PGO_TRY(allreduce_sum(g, g->vw, (size_t)g->N * 6));
  CUDA_TRY(cudaMemcpyAsync(y, g->vw, nv, cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, g->ev0, g->ev1));
  return PGO_OK;
}

// Host-only symbolic analysis for solver type t.  AUTO takes the factor only when it is cheap: chain-like graphs (<= 64
// levels, node degree <= 16, fill <= 8x); mesh-like graphs (sphere, grids, dense random loops) go to block-Jacobi PCG.
static int run_symbolic(const pgo_graph* g, int t, LevelCholSymbolic* S) {
  const bool autosel = t == PGO_LINEAR_AUTO;
  return level_chol_symbolic(S, g->N, g->active_h.data(), g->row_ptr_h.data(), g->col_idx_h.data(), autosel ? 8.0 : 1e30,
                             autosel ? 64 : 8192, autosel ? 16 : (1 << 30));
}

static int resolve_linear_solver(pgo_graph* g, const pgo_solver_options* o) {
  int t = o->linear_solver_type;
  if (g->world > 1) return PGO_LINEAR_PCG_BLOCK_JACOBI;  // the factor is not distributed
  if (t == PGO_LINEAR_AUTO || t == PGO_LINEAR_PCG_LEVEL_CHOLESKY) {
    if (!g->chol) {
      LevelChol* c = nullptr;
      int rc = PGO_OK;
      if (g->sym_future.valid() && g->sym_solver_type == t) {
        rc = g->sym_future.get();                      // started by pgo_solve_pose_graph
      } else {
        if (g->sym_future.valid()) g->sym_future.get();
        g->sym.reset(new LevelCholSymbolic());
        rc = run_symbolic(g, t, g->sym.get());
      }
      if (rc == PGO_OK) rc = level_chol_analyze(&c, g->device, g->N, *g->sym, g->stream);
      g->sym.reset();
      if (rc == PGO_OK) g->chol = c;
      else { if (c) level_chol_destroy(c, g->device); if (t == PGO_LINEAR_PCG_LEVEL_CHOLESKY) return rc; }
    }
    if (g->chol && g->chol->usable) return PGO_LINEAR_PCG_LEVEL_CHOLESKY;
    if (t == PGO_LINEAR_PCG_LEVEL_CHOLESKY) return set_error(PGO_ERR_NUMERICAL, "level Cholesky analysis failed");
    return PGO_LINEAR_PCG_BLOCK_JACOBI;
  }
  return PGO_LINEAR_PCG_BLOCK_JACOBI;
}

// LevenbergMarquardtStrategy::ComputeStep: D = diagonal / radius (new, reused or given), then solve
// (H + D) x = b with the chosen solver; x -> g->vx, stats -> g->scalars (after fetch).
static int linear_solve_device(pgo_graph* g, const pgo_solver_options* o, int solver, const double* b, const LmDiagonal& lm) {
  if (g->world == 1 && solver == PGO_LINEAR_PCG_LEVEL_CHOLESKY) {
    // the LM diagonal is formed inside the factor kernel
    return level_chol_solve(g->chol, bsr_view(g), lm, g->active, b, g->vx, g->vr, g->vu, g->vw, g->vp, g->vs,
                            std::min(o->pcg_max_iterations, 200), o->pcg_tolerance, std::max(o->pcg_tolerance, o->direct_residual_accept),
                            o->pcg_num_ctas, g->scalars,
                            g->stream, &g->launches);
  }
  const int tpb = 128;
  lm_prepare_kernel<<<(g->N + tpb - 1) / tpb, tpb, 0, g->stream>>>(g->N, g->Hdiag, g->active, lm.mode, lm.min_diag, lm.max_diag,
                                                                   lm.radius, lm.diagonal, lm.dlm, g->Minv);
  g->launches++;
  static const bool force_stream_pcg = getenv("PGO_FORCE_STREAM_PCG") != nullptr;   // tests: multi-GPU code path on one GPU
  // large graphs: separate launches at full occupancy beat the persistent kernel (whose grid barriers only pay when an
  // iteration is a few microseconds long)
  if (g->world > 1 || force_stream_pcg || g->N >= kStreamPcgMinPoses) return pcg_multi(g, o, b);
  return launch_pcg(g, o, b);
}

extern "C" int pgo_graph_linear_solve(pgo_graph* g, const pgo_solver_options* options, const double* d,
                                      const double* b, double* y, int* iterations, double* relative_residual,
                                      float* elapsed_ms) {
  if (!g || !options || !d || !b || !y) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_linear_solve: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  const size_t nv = (size_t)g->N * 6 * sizeof(double);
  const int solver = resolve_linear_solver(g, options);
  if (solver < 0) return solver;
  CUDA_TRY(cudaMemcpyAsync(g->dlm, d, nv, cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->vb, b, nv, cudaMemcpyHostToDevice, g->stream));
  PGO_TRY(zero_scalars(g));
  CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
  const LmDiagonal lm = {2, 0.0, 0.0, 1.0, g->diagonal, g->dlm};
  PGO_TRY(linear_solve_device(g, options, solver, g->vb, lm));
  CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
  PGO_TRY(fetch_scalars(g));
  CUDA_TRY(cudaMemcpy(y, g->vx, nv, cudaMemcpyDeviceToHost));
  if (iterations) *iterations = g->scalars_h->pcg_iterations;
  if (relative_residual) *relative_residual = g->scalars_h->pcg_gamma0 > 0 ? std::sqrt(g->scalars_h->pcg_gamma / g->scalars_h->pcg_gamma0) : 0.0;
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, g->ev0, g->ev1));
  return PGO_OK;
}

// ------------------------------------------------------------------------------------------------
// ceres::Solve: TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy (ceres 1.13 flow),
// the same sequence of decisions as oracle/pgo_oracle.c:oracle_solve, driven from the host with
// two stream synchronisations per iteration.
// ------------------------------------------------------------------------------------------------
extern "C" int pgo_graph_solve(pgo_graph* g, const pgo_solver_options* opt, pgo_solver_summary* summary,
                               pgo_iteration_summary* log, int log_cap) {
  if (!g || !opt || !summary) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_solve: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  const double t_begin = wall_s();
  std::memset(summary, 0, sizeof *summary);
  const long long launches0 = g->launches;
  const int N = g->N;
  const int tpb = 128, nblk = (N + tpb - 1) / tpb;
  float ms = 0.f;
  auto push_log = [&](const pgo_iteration_summary& it) {
    if (log && summary->num_iterations < log_cap) log[summary->num_iterations] = it;
    summary->num_iterations++;
  };
  const int solver = resolve_linear_solver(g, opt);
  if (solver < 0) return solver;
  summary->linear_solver_used = solver;
  if (!g->Hdiag_alt) {
    PGO_TRY(dev_alloc(g, &g->Hdiag_alt, (size_t)N * 36));
    PGO_TRY(dev_alloc(g, &g->Hoff_alt, (size_t)g->nnz_off * 36));
    PGO_TRY(dev_alloc(g, &g->grad_alt, (size_t)N * 6));
  }
  summary->hessian_blocks = g->nnz_off + g->N;
  if (solver == PGO_LINEAR_PCG_LEVEL_CHOLESKY) { summary->factor_blocks = g->chol->factor_blocks; summary->factor_levels = g->chol->num_levels; }
  summary->time_setup_s = g->setup_s;

  // ---- IterationZero: evaluate, Jacobi scaling from the unscaled diagonal, re-linearize scaled ----
  CUDA_TRY(cudaMemcpyAsync(g->scale, g->scale_eval, (size_t)N * 6 * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
  PGO_TRY(zero_scalars(g));
  CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
  PGO_TRY(linearize_full(g, g->poses, g->scale, opt->loss_type, opt->loss_a));
  summary->num_linearizations++;
  if (opt->jacobi_scaling) {
    jacobi_scale_kernel<<<nblk, tpb, 0, g->stream>>>(N, g->Hdiag, g->active, 1, g->scale);
    g->launches++;
    PGO_TRY(zero_scalars(g));
    PGO_TRY(linearize_full(g, g->poses, g->scale, opt->loss_type, opt->loss_a));
    summary->num_linearizations++;
  }
  CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
  xnorm_kernel<<<(N + 255) / 256, 256, 0, g->stream>>>(N, g->poses, g->active, g->scalars);
  gradient_norm_kernel<<<(N + 255) / 256, 256, 0, g->stream>>>(N, g->poses, g->grad, g->scale, g->active, nullptr, g->scalars);
  g->launches += 2;
  PGO_TRY(fetch_scalars(g));
  CUDA_TRY(cudaEventElapsedTime(&ms, g->ev0, g->ev1));
  summary->time_linearize_ms += ms;

  double x_cost = g->scalars_h->cost;
  double x_norm = std::sqrt(g->scalars_h->x_norm2);
  double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_consecutive_invalid = 0, iter = 0;
  pgo_iteration_summary it;
  std::memset(&it, 0, sizeof it);
  it.cost = x_cost; it.trust_region_radius = radius;
  { long long bits = (long long)g->scalars_h->gmax_bits; double gm; std::memcpy(&gm, &bits, 8); it.gradient_max_norm = gm; }
  it.gradient_norm = std::sqrt(g->scalars_h->gnorm2);
  summary->initial_cost = x_cost;
  push_log(it);
  if (!std::isfinite(x_cost)) {
    summary->termination_type = PGO_FAILURE;
    snprintf(summary->message, sizeof summary->message, "Initial cost is not finite.");
    summary->final_cost = x_cost;
    summary->time_total_s = wall_s() - t_begin;
    return PGO_OK;
  }

  for (;;) {
    if (iter >= opt->max_num_iterations) { summary->termination_type = PGO_NO_CONVERGENCE; snprintf(summary->message, sizeof summary->message, "Maximum number of iterations reached. Number of iterations: %d.", iter); break; }
    if (it.gradient_max_norm <= opt->gradient_tolerance) { summary->termination_type = PGO_CONVERGENCE; snprintf(summary->message, sizeof summary->message, "Gradient tolerance reached. Gradient max norm: %e <= %e", it.gradient_max_norm, opt->gradient_tolerance); break; }
    if (radius < opt->min_trust_region_radius) { summary->termination_type = PGO_CONVERGENCE; snprintf(summary->message, sizeof summary->message, "Minimum trust region radius reached."); break; }
    ++iter;
    { const double gm = it.gradient_max_norm, gn = it.gradient_norm; std::memset(&it, 0, sizeof it); it.iteration = iter; it.gradient_max_norm = gm; it.gradient_norm = gn; }

    // ---- LevenbergMarquardtStrategy::ComputeStep: D = diag / radius, solve (H + D) y = g, step = -y ----
    PGO_TRY(zero_scalars(g));
    CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
    const LmDiagonal lm = {reuse_diagonal ? 1 : 0, opt->min_lm_diagonal, opt->max_lm_diagonal, radius, g->diagonal, g->dlm};
    const double th0 = wall_s();
    PGO_TRY(linear_solve_device(g, opt, solver, g->grad, lm));
    CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
    const double th1 = wall_s();
    reuse_diagonal = true;
    // ---- candidate = Plus(x, -y .* scale), |step|, |x_cand|, and -- speculatively -- the full linearisation at the
    //      candidate (cost, H, g, gradient norms) into the alternate system: everything the decision needs in one sync
    plus_kernel<<<(N + 255) / 256, 256, 0, g->stream>>>(N, g->poses, g->vx, g->scale, g->active, -1.0, g->poses_cand, g->scalars);
    g->launches++;
    swap_system(g);
    CUDA_TRY(cudaEventRecord(g->ev2, g->stream));
    PGO_TRY(linearize_full(g, g->poses_cand, g->scale, opt->loss_type, opt->loss_a));
    CUDA_TRY(cudaEventRecord(g->ev3, g->stream));
    summary->num_linearizations++;
    gradient_norm_kernel<<<(N + 255) / 256, 256, 0, g->stream>>>(N, g->poses_cand, g->grad, g->scale, g->active, nullptr, g->scalars);
    g->launches++;
    PGO_TRY(fetch_scalars(g));
    const double th2 = wall_s();
    CUDA_TRY(cudaEventElapsedTime(&ms, g->ev0, g->ev1));
    summary->time_linear_solver_ms += ms;
    { float ms2 = 0.f; CUDA_TRY(cudaEventElapsedTime(&ms2, g->ev2, g->ev3)); summary->time_linearize_ms += ms2; }
    if (opt->verbose >= 2)
      fprintf(stderr, "[pgo host] it %d: enqueue solver %.1f us, enqueue plus+linearize + wait %.1f us, solver on GPU %.1f us\n", iter,
              1e6 * (th1 - th0), 1e6 * (th2 - th1), 1e3 * ms);
    const DeviceScalars sc = *g->scalars_h;
    summary->total_pcg_iterations += sc.pcg_iterations;
    it.linear_solver_iterations = sc.pcg_iterations;
    it.pcg_relative_residual = sc.pcg_gamma0 > 0 ? std::sqrt(std::fabs(sc.pcg_gamma) / sc.pcg_gamma0) : 0.0;
    it.trust_region_radius = radius;

    // model_cost_change = -(J step)^T (r + J step / 2) = y^T g - y^T H y / 2
    const double model_cost_change = sc.xtb - 0.5 * (sc.xtAx - sc.xtDx);
    bool step_valid = sc.pcg_flag < 2 && std::isfinite(model_cost_change) && model_cost_change > 0.0;   // 2: CG breakdown, 3: pivot failure
    if (opt->verbose)
      fprintf(stderr, "[pgo] it %d radius %.3e pcg %d (flag %d, rel %.2e) model %.6e\n", iter, radius, sc.pcg_iterations,
              sc.pcg_flag, it.pcg_relative_residual, model_cost_change);
    if (!step_valid) {
      swap_system(g);   // discard the speculative system
      it.step_is_valid = 0; it.cost = x_cost;
      if (++num_consecutive_invalid >= opt->max_num_consecutive_invalid_steps) {
        summary->termination_type = PGO_FAILURE;
        snprintf(summary->message, sizeof summary->message, "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %d", opt->max_num_consecutive_invalid_steps);
        push_log(it);
        break;
      }
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      summary->num_unsuccessful_steps++;
      push_log(it);
      continue;
    }
    num_consecutive_invalid = 0;
    it.step_is_valid = 1;
    double cand_cost = sc.cost;
    if (!std::isfinite(cand_cost)) cand_cost = 1.7976931348623157e308;
    const double step_norm = std::sqrt(sc.step_norm2);
    it.step_norm = step_norm;
    if (step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
      summary->termination_type = PGO_CONVERGENCE;
      snprintf(summary->message, sizeof summary->message, "Parameter tolerance reached. Relative step_norm: %e <= %e.", step_norm / (x_norm + opt->parameter_tolerance), opt->parameter_tolerance);
      swap_system(g);
      it.cost = x_cost; push_log(it);
      break;
    }
    const double cost_change = x_cost - cand_cost;
    it.cost_change = cost_change;
    if (std::fabs(cost_change) <= opt->function_tolerance * x_cost) {
      summary->termination_type = PGO_CONVERGENCE;
      snprintf(summary->message, sizeof summary->message, "Function tolerance reached. |cost_change|/cost: %e <= %e", std::fabs(cost_change) / x_cost, opt->function_tolerance);
      swap_system(g);
      it.cost = x_cost; push_log(it);
      break;
    }
    const double relative_decrease = cost_change / model_cost_change;
    it.relative_decrease = relative_decrease;
    if (relative_decrease > opt->min_relative_decrease) {
      // HandleSuccessfulStep: x = candidate; its linearisation is already the current system
      std::swap(g->poses, g->poses_cand);
      x_norm = std::sqrt(sc.x_norm2);
      x_cost = sc.cost;
      { long long bits = (long long)sc.gmax_bits; double gm; std::memcpy(&gm, &bits, 8); it.gradient_max_norm = gm; }
      it.gradient_norm = std::sqrt(sc.gnorm2);
      it.step_is_successful = 1; it.cost = x_cost;
      summary->num_successful_steps++;
      double t = 2.0 * relative_decrease - 1.0;
      t = 1.0 - t * t * t;
      if (t < 1.0 / 3.0) t = 1.0 / 3.0;
      radius = std::min(radius / t, opt->max_trust_region_radius);
      decrease_factor = 2.0; reuse_diagonal = false;
    } else {
      swap_system(g);   // HandleUnsuccessfulStep: discard the speculative system
      it.step_is_successful = 0; it.cost = x_cost;
      summary->num_unsuccessful_steps++;
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
    push_log(it);
  }
  summary->final_cost = x_cost;
  summary->kernel_launches = g->launches - launches0;
  summary->time_total_s = wall_s() - t_begin;
  return PGO_OK;
}

// ------------------------------------------------------------------------------------------------
// Loop-edge candidate search (the caller side of the path): REF/test/generate_edges_from_trajectory_origion.cpp
// ------------------------------------------------------------------------------------------------
extern "C" int pgo_edge_candidates(int device, int n_frames, const double* positions, double search_radius, int min_frame_gap,
                                   long long* row_ptr, int* candidates, long long capacity, long long* total) {
  if (n_frames <= 0 || !positions || !row_ptr || !total || min_frame_gap < 0 || !(search_radius >= 0.0))
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_edge_candidates: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(PGO_ERR_NO_DEVICE, "pgo_edge_candidates: no CUDA device available (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return set_error(PGO_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  CUDA_TRY(cudaSetDevice(device));
  const int n = n_frames;
  // the reference keeps poses in CV_32F matrices: positions are rounded to float before the search
  std::vector<float> soa(3 * (size_t)n);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) soa[(size_t)k * n + i] = (float)positions[3 * (size_t)i + k];
  const float r = (float)search_radius;
  const float r2 = r * r;
  cudaStream_t stream = nullptr;
  CUDA_TRY(pool_stream(device, &stream));
  float* d_pos = nullptr; int* d_counts = nullptr; long long* d_ptr = nullptr; int* d_idx = nullptr;
  const size_t pos_bytes = soa.size() * sizeof(float), cnt_bytes = (size_t)n * sizeof(int), ptr_bytes = ((size_t)n + 1) * sizeof(long long);
  size_t idx_bytes = 0;
  int rc = PGO_OK;
  auto cleanup = [&]() {
    cudaStreamSynchronize(stream);
    pool_free(device, d_pos, pos_bytes); pool_free(device, d_counts, cnt_bytes); pool_free(device, d_ptr, ptr_bytes);
    if (d_idx) pool_free(device, d_idx, idx_bytes);
    pool_stream_release(device, stream);
  };
#define CAND_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rc = set_error(PGO_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); cleanup(); return rc; } } while (0)
  CAND_TRY(pool_alloc(device, reinterpret_cast<void**>(&d_pos), pos_bytes));
  CAND_TRY(pool_alloc(device, reinterpret_cast<void**>(&d_counts), cnt_bytes));
  CAND_TRY(pool_alloc(device, reinterpret_cast<void**>(&d_ptr), ptr_bytes));
  CAND_TRY(cudaMemcpyAsync(d_pos, soa.data(), pos_bytes, cudaMemcpyHostToDevice, stream));
  const int sms = pool_num_sms(device);
  // short trajectories: one frame per warp (parallelism); long ones: eight frames per warp (L2 traffic / 8)
  const bool blocked = n >= 16 * 8 * sms;
  const int groups = blocked ? (n + 7) / 8 : n;
  int per_sm = 1;   // grid = what is resident: groups are dealt longest first, a queued CTA would start late with a long one
  CAND_TRY(blocked ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, edge_candidates_kernel<true, 8>, 256, 0)
                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, edge_candidates_kernel<true, 1>, 256, 0));
  const int ctas = std::max(1, std::min((groups + 7) / 8, std::max(per_sm, 1) * sms));
  if (blocked) edge_candidates_kernel<false, 8><<<ctas, 256, 0, stream>>>(n, d_pos, d_pos + n, d_pos + 2 * (size_t)n, r2, min_frame_gap, d_counts, nullptr, nullptr);
  else edge_candidates_kernel<false, 1><<<ctas, 256, 0, stream>>>(n, d_pos, d_pos + n, d_pos + 2 * (size_t)n, r2, min_frame_gap, d_counts, nullptr, nullptr);
  CAND_TRY(cudaGetLastError());
  std::vector<int> counts(n);
  CAND_TRY(cudaMemcpyAsync(counts.data(), d_counts, cnt_bytes, cudaMemcpyDeviceToHost, stream));
  CAND_TRY(cudaStreamSynchronize(stream));
  row_ptr[0] = 0;
  for (int c = 0; c < n; ++c) row_ptr[c + 1] = row_ptr[c] + counts[c];
  *total = row_ptr[n];
  if (candidates != nullptr) {
    if (capacity < *total) { cleanup(); return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_edge_candidates: capacity %lld < %lld candidates", capacity, *total); }
    idx_bytes = std::max<size_t>((size_t)*total, 1) * sizeof(int);
    CAND_TRY(pool_alloc(device, reinterpret_cast<void**>(&d_idx), idx_bytes));
    CAND_TRY(cudaMemcpyAsync(d_ptr, row_ptr, ptr_bytes, cudaMemcpyHostToDevice, stream));
    if (blocked) edge_candidates_kernel<true, 8><<<ctas, 256, 0, stream>>>(n, d_pos, d_pos + n, d_pos + 2 * (size_t)n, r2, min_frame_gap, nullptr, d_ptr, d_idx);
    else edge_candidates_kernel<true, 1><<<ctas, 256, 0, stream>>>(n, d_pos, d_pos + n, d_pos + 2 * (size_t)n, r2, min_frame_gap, nullptr, d_ptr, d_idx);
    CAND_TRY(cudaGetLastError());
    if (*total > 0) CAND_TRY(cudaMemcpyAsync(candidates, d_idx, (size_t)*total * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CAND_TRY(cudaStreamSynchronize(stream));
  }
#undef CAND_TRY
  cleanup();
  return PGO_OK;
}

extern "C" int pgo_solve_pose_graph(int device, int n_poses, double* poses, int n_edges, const int* edge_ids,
                                    const double* edge_meas, const double* edge_sqrt_info,
                                    const unsigned char* pose_const, const pgo_solver_options* options,
                                    pgo_solver_summary* summary, pgo_iteration_summary* iteration_log,
                                    int iteration_log_capacity) {
  pgo_solver_options defaults;
  if (!options) { pgo_default_options(&defaults); options = &defaults; }
  pgo_graph* g = nullptr;
  g_symbolic_hint = options->linear_solver_type;     // pgo_graph_create starts the symbolic analysis on a helper thread
  const int crc = pgo_graph_create(&g, device, n_poses, n_edges, poses, edge_ids, edge_meas, edge_sqrt_info, pose_const);
  g_symbolic_hint = -1;
  PGO_TRY(crc);
  int rc = pgo_graph_solve(g, options, summary, iteration_log, iteration_log_capacity);
  if (rc == PGO_OK) rc = pgo_graph_get_poses(g, poses);
  pgo_graph_destroy(g);
  return rc;
}
