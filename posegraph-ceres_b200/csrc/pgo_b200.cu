// pgo_b200.cu -- host side of libpgo_b200.so: graph upload + block structure analysis, the
// Levenberg-Marquardt driver (a restatement of ceres::internal::TrustRegionMinimizer +
// LevenbergMarquardtStrategy as the reference configures them,
// REF/test/pose_graph_ceres_plus_finial.cpp:500-514) and the C-ABI of include/pgo_b200.h.
// All numerical work happens in the kernels of pgo_kernels.cuh; there is no CPU fallback.
#include "../../include/pgo_b200.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "pgo_kernels.cuh"
#include "pgo_pool.cuh"
#include "pgo_amg_host.hpp"

namespace pgo { struct PcgMultiState; struct Amg; }

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

static void amg_destroy(pgo::Amg* M, int device);
static double wall_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct pgo_graph {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int N = 0, E = 0, T = 0;          // LOCAL poses (owned + halo copies), local edges, edge tiles
  int n_own = 0;                    // block rows / variables stored here (== N on one GPU); local ids [n_own, N) are halo poses
  int N_global = 0, E_global = 0;   // the whole problem (== N, E on one GPU)
  int g0 = 0;                       // global id of local pose 0
  std::vector<int> part_off;        // [world + 1] contiguous pose ranges per rank
  std::vector<int> halo_gid;        // [N - n_own] global ids of the halo poses, ascending
  // level-0 halo exchange plan (multi-GPU): neighbours, what they need from here, where their values land
  std::vector<int> nbr, send_ptr, recv_ptr, send_idx_h;
  int* send_idx = nullptr;
  double* halo_sendbuf = nullptr;
  double* gather_buf = nullptr;     // [N_global][8] pose all-gather staging (multi-GPU get_poses)
  // global structure kept for the multilevel hierarchy (host): pattern, variable flags, setup-time positions
  std::vector<int> grow_ptr_h, gcol_idx_h;
  std::vector<unsigned char> gactive_h;
  std::vector<double> pos0_h;       // [N_global][3]
  pgo::Amg* amg = nullptr;
  // one-shot entry point: graphs are kept per topology (see graph_cache below)
  unsigned long long topo_hash = 0;
  std::vector<int> topo_edge_ids;            // [E][2] as passed by the caller
  std::vector<unsigned char> topo_const;     // [N]
  std::vector<int> edge_pos;                 // [E] tile position of the caller's edge e (edges are processed in pose order)
  std::vector<int> edge_gid;                 // multi-GPU: caller's (global) index of local edge e; empty on one GPU
  double* edge_loss = nullptr;               // [E] per-edge loss in processing order (encoded), valid while edge_loss_set
  bool edge_loss_set = false;
  bool edges_reordered = false;
  std::vector<EdgeCoreTile> core_host;       // packed tiles (indices stay, measurements are refreshed)
  std::vector<EdgeInfoTile> info_host;
  // device-resident LM loop (pgo_lm.cuh)
  pgo::LmState* lm_state = nullptr;
  pgo::LmState* lm_ring = nullptr;          // pinned: 3 look-behind slots + 1
  cudaEvent_t lm_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  pgo_iteration_summary* lm_log = nullptr;
  int lm_log_cap = 0;
  // one LM iteration captured as a CUDA graph (solvers without a host-polled inner loop); valid for lm_graph_key
  cudaGraphExec_t lm_graph = nullptr;
  std::vector<unsigned char> lm_graph_key;
  int lm_graph_kernels = 0;
  bool failed = false;                       // a collective call returned an error on this rank
  long long comm_calls = 0, comm_bytes = 0;   // NCCL calls / payload bytes sent by this rank (multi-GPU)
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
static void build_pattern(int N, int E, const int* edge_ids, const unsigned char* pose_const, HostPattern* out, int n_own = -1) {
  // n_own < N (multi-GPU slice): rows exist only for the owned poses [0, n_own); poses >= n_own are halo columns
  if (n_own < 0) n_own = N;
  out->active.assign(N, 0);
  for (int e = 0; e < E; ++e) { out->active[edge_ids[2 * e]] = 1; out->active[edge_ids[2 * e + 1]] = 1; }
  if (pose_const) for (int i = 0; i < N; ++i) if (pose_const[i] == 1) out->active[i] = 0;   // 2 / 3: only p / only q constant
  // bucket the half-edges (row -> col) by row with a counting sort, then order each (short) row by column
  struct Half { int col; int idx; };
  std::vector<int> start(n_own + 1, 0);
  size_t n_half = 0;
  for (int e = 0; e < E; ++e) {
    const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
    if (out->active[a] && out->active[b]) {
      if (a < n_own) { start[a + 1]++; ++n_half; }
      if (b < n_own) { start[b + 1]++; ++n_half; }
    }
  }
  for (int i = 0; i < n_own; ++i) start[i + 1] += start[i];
  std::vector<Half> halves(n_half);
  {
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int e = 0; e < E; ++e) {
      const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
      if (out->active[a] && out->active[b]) {
        if (a < n_own) halves[fill[a]++] = {b, 2 * e};
        if (b < n_own) halves[fill[b]++] = {a, 2 * e + 1};
      }
    }
  }
  out->half_slot.assign(2 * (size_t)E, -1);
  out->row_ptr.assign(n_own + 1, 0);
  out->col_idx.clear();
  out->col_idx.reserve(n_half);
  for (int i = 0; i < n_own; ++i) {
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
  for (int i = 0; i < n_own; ++i) out->row_ptr[i + 1] += out->row_ptr[i];
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
  o->pcg_tolerance = 1e-8;      // in the M^-1 norm; a direct factorisation of these systems is no more accurate (kappa ~ 1e6..1e9)
  o->pcg_num_ctas = 0;
  o->direct_residual_accept = 1e-8;
  o->verbose = 0;
  o->edge_loss_type = nullptr;
  o->edge_loss_a = nullptr;
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
  // a rank that failed inside a collective solve must not leave its peers blocked in NCCL: abort instead of a graceful destroy
  if (g->amg) amg_destroy(g->amg, g->device);   // (its peer windows close with a last collective: before the communicator goes)
  if (g->comm) { if (g->failed) ncclCommAbort(g->comm); else ncclCommDestroy(g->comm); }
  if (g->chol) level_chol_destroy(g->chol, g->device);
  if (g->lm_graph) cudaGraphExecDestroy(g->lm_graph);
  for (auto& blk : g->blocks) pool_free(g->device, blk.first, blk.second);
  pool_pinned_release(g->device, g->scalars_h);
  pool_pinned_release(g->device, g->pcgm_state_h);
  pool_pinned_release(g->device, g->lm_ring);
  for (int k = 0; k < 4; ++k) pool_event_release(g->device, g->lm_ev[k]);
  pool_event_release(g->device, g->ev0);
  pool_event_release(g->device, g->ev1);
  pool_event_release(g->device, g->ev2);
  pool_event_release(g->device, g->ev3);
  pool_stream_release(g->device, g->own_stream);
  delete g;
}

static void graph_cache_clear(int device);
extern "C" void pgo_release_cached_memory(int device) { graph_cache_clear(device); pool_release(device); }

static thread_local int g_symbolic_hint = -1;   // set by pgo_solve_pose_graph around its pgo_graph_create call
static thread_local bool g_keep_host_tiles = false;   // ditto: the graph will be cached per topology
static int run_symbolic(const pgo_graph* g, int t, LevelCholSymbolic* S);

// What one rank holds.  One GPU: the whole problem (n_own == n_loc, no halo).  Multi-GPU: the block rows of its own
// poses; edges that touch them; copies ("halo") of the other endpoint of every cut edge.
struct LocalProblem {
  int n_loc = 0, n_own = 0, n_edges = 0;
  const double* poses = nullptr;             // [n_loc][7]
  const int* edge_ids = nullptr;             // [n_edges][2] LOCAL ids
  const double* edge_meas = nullptr;
  const double* edge_sqrt_info = nullptr;
  const unsigned char* pose_const = nullptr; // [n_loc]
};

static int halo_exchange(pgo_graph* g, const std::vector<int>& nbr, const std::vector<int>& send_ptr, const std::vector<int>& recv_ptr,
                         const int* send_idx, int n_own, double* v, int width, const int* skip);

static int graph_create_local(pgo_graph* g, const LocalProblem& in) {
  const int device = g->device;
  const double t0 = wall_s();
  static const bool prof = getenv("PGO_PROFILE_HOST") != nullptr;
  double tp = t0;
  auto lap = [&](const char* what) { if (prof) { const double t = wall_s(); fprintf(stderr, "[pgo create] %-28s %8.1f us\n", what, 1e6 * (t - tp)); tp = t; } };
  g->N = in.n_loc; g->n_own = in.n_own; g->E = in.n_edges; g->T = (in.n_edges + kTile - 1) / kTile;
  const int N = g->N, E = g->E, T = g->T, n_own = g->n_own;
  CUDA_TRY(pool_stream(device, &g->own_stream));
  g->stream = g->own_stream;
  CUDA_TRY(pool_event(device, &g->ev0));
  CUDA_TRY(pool_event(device, &g->ev1));
  CUDA_TRY(pool_event(device, &g->ev2));
  CUDA_TRY(pool_event(device, &g->ev3));
  g->num_sms = pool_num_sms(device);

  // ---- which poses are variables, block-CSR pattern of the off-diagonal part ----
  lap("stream/events");
  HostPattern pat;
  build_pattern(N, E, in.edge_ids, in.pose_const, &pat, n_own);
  lap("block-CSR pattern");
  g->active_h.swap(pat.active);
  g->row_ptr_h.swap(pat.row_ptr);
  g->col_idx_h.swap(pat.col_idx);
  g->nnz_off = (long long)g->col_idx_h.size();
  g->has_dup_blocks = pat.has_dup;
  if (g->world == 1 && (g_symbolic_hint == PGO_LINEAR_AUTO || g_symbolic_hint == PGO_LINEAR_PCG_LEVEL_CHOLESKY)) {
    // the pattern is final: analyse the elimination order on a helper thread while this thread packs and uploads
    g->sym.reset(new LevelCholSymbolic());
    g->sym_solver_type = g_symbolic_hint;
    pgo_graph* gp = g;
    const int t = g_symbolic_hint;
    g->sym_future = std::async(std::launch::async, [gp, t]() { return run_symbolic(gp, t, gp->sym.get()); });
  }

  // ---- identity information? (after the helper thread is off: the symbolic analysis is the longer leg) ----
  g->identity_info = true;
  if (in.edge_sqrt_info) {
    static const double eye[36] = {1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1};
    for (int e = 0; e < E && g->identity_info; ++e)
      if (std::memcmp(in.edge_sqrt_info + 36 * (size_t)e, eye, sizeof eye) != 0) {
        // memcmp also flags -0.0; confirm numerically
        for (int k = 0; k < 36; ++k) if (in.edge_sqrt_info[36 * (size_t)e + k] != eye[k]) { g->identity_info = false; break; }
      }
  }
  if (g->world > 1) {
    // every rank must compile the same kernel variant: information is identity only if it is on all ranks
    int flag = g->identity_info ? 1 : 0, *flag_d = nullptr;
    PGO_TRY(dev_alloc(g, &flag_d, 1));
    CUDA_TRY(cudaMemcpyAsync(flag_d, &flag, sizeof(int), cudaMemcpyHostToDevice, g->stream));
    NCCL_TRY(ncclAllReduce(flag_d, flag_d, 1, ncclInt, ncclMin, g->comm, g->stream));
    CUDA_TRY(cudaMemcpyAsync(&flag, flag_d, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    g->identity_info = flag != 0;
  }
  lap("identity scan");
  // ---- edge tiles (field-major, one warp per tile) ----
  std::vector<EdgeCoreTile> core_h(std::max(T, 1));
  std::memset(core_h.data(), 0, core_h.size() * sizeof(EdgeCoreTile));
  std::vector<EdgeInfoTile> info_h;
  if (!g->identity_info) { info_h.resize(std::max(T, 1)); std::memset(info_h.data(), 0, info_h.size() * sizeof(EdgeInfoTile)); }
  static const double eye_row[6][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}, {0, 0, 0, 1, 0, 0}, {0, 0, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 1}};
  // Processing order = pose order (counting sort by the larger endpoint, stable): the persistent kernel walks the tiles
  // in index order, so every pose's diagonal block and gradient then receive their contributions while their lines are
  // still in L2 (a 1M-pose diagonal is 288 MB; in generator order -- all odometry edges, then all cross edges -- every
  // line made two round trips to HBM).  Odometry chains keep their order: lane l's end pose is lane l-1's begin pose.
  g->edge_pos.resize((size_t)E);
  {
    std::vector<int> cnt((size_t)N + 1, 0);
    for (int e = 0; e < E; ++e) cnt[std::max(in.edge_ids[2 * e], in.edge_ids[2 * e + 1]) + 1]++;
    for (int i = 0; i < N; ++i) cnt[i + 1] += cnt[i];
    bool identity = true;
    for (int e = 0; e < E; ++e) {
      const int p = cnt[std::max(in.edge_ids[2 * e], in.edge_ids[2 * e + 1])]++;
      g->edge_pos[e] = p;
      identity &= p == e;
    }
    g->edges_reordered = !identity;
  }
  for (int e = 0; e < E; ++e) {
    const int pos = g->edge_pos[e];
    EdgeCoreTile& t = core_h[pos / kTile];
    const int l = pos % kTile;
    const int a = in.edge_ids[2 * e], b = in.edge_ids[2 * e + 1];
    t.a[l] = a; t.b[l] = b;
    t.slot_ab[l] = pat.half_slot[2 * (size_t)e];
    t.slot_ba[l] = pat.half_slot[2 * (size_t)e + 1];
    for (int k = 0; k < 7; ++k) t.meas[k][l] = in.edge_meas[7 * (size_t)e + k];
    if (!g->identity_info)
      for (int k = 0; k < 36; ++k) info_h[pos / kTile].S[k][l] = in.edge_sqrt_info ? in.edge_sqrt_info[36 * (size_t)e + k] : eye_row[k / 6][k % 6];
  }
  for (int e = E; e < T * kTile; ++e) { core_h[e / kTile].slot_ab[e % kTile] = -1; core_h[e / kTile].slot_ba[e % kTile] = -1; }

  lap("edge tiles");
  if (g_keep_host_tiles) { g->core_host = core_h; g->info_host = info_h; }
  // ---- device allocations + uploads ----
  PGO_TRY(dev_alloc(g, &g->poses, (size_t)N * 8));
  PGO_TRY(dev_alloc(g, &g->poses_cand, (size_t)N * 8));
  PGO_TRY(dev_alloc(g, &g->poses_snap, (size_t)N * 8));
  PGO_TRY(dev_alloc(g, &g->scale, (size_t)N * 6));
  PGO_TRY(dev_alloc(g, &g->scale_eval, (size_t)N * 6));
  PGO_TRY(dev_alloc(g, &g->core, (size_t)std::max(T, 1)));
  if (!g->identity_info) PGO_TRY(dev_alloc(g, &g->info, (size_t)std::max(T, 1)));
  PGO_TRY(dev_alloc(g, &g->Hdiag, (size_t)n_own * 36));
  PGO_TRY(dev_alloc(g, &g->Hoff, (size_t)g->nnz_off * 36));
  PGO_TRY(dev_alloc(g, &g->row_ptr, (size_t)n_own + 1));
  PGO_TRY(dev_alloc(g, &g->col_idx, (size_t)g->nnz_off));
  PGO_TRY(dev_alloc(g, &g->grad, (size_t)n_own * 6));
  PGO_TRY(dev_alloc(g, &g->grad_unscaled, (size_t)n_own * 6));
  PGO_TRY(dev_alloc(g, &g->diagonal, (size_t)n_own * 6));
  PGO_TRY(dev_alloc(g, &g->dlm, (size_t)n_own * 6));
  PGO_TRY(dev_alloc(g, &g->Minv, (size_t)n_own * 36));
  // vectors that an SpMV gathers from carry the halo tail
  for (double** v : {&g->vx, &g->vr, &g->vu, &g->vw, &g->vp, &g->vs, &g->vb}) PGO_TRY(dev_alloc(g, v, (size_t)N * 6));
  PGO_TRY(dev_alloc(g, &g->active, (size_t)N));
  PGO_TRY(dev_alloc(g, &g->scalars, 1));
  static_assert(sizeof(DeviceScalars) <= kPinnedBytes, "pinned scalars");
  CUDA_TRY(pool_pinned(device, reinterpret_cast<void**>(&g->scalars_h)));
  PGO_TRY(dev_alloc(g, &g->barrier, 4));

  lap("device allocations");
  CUDA_TRY(cudaMemcpyAsync(g->core, core_h.data(), (size_t)std::max(T, 1) * sizeof(EdgeCoreTile), cudaMemcpyHostToDevice, g->stream));
  if (!g->identity_info) CUDA_TRY(cudaMemcpyAsync(g->info, info_h.data(), (size_t)std::max(T, 1) * sizeof(EdgeInfoTile), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->row_ptr, g->row_ptr_h.data(), ((size_t)n_own + 1) * sizeof(int), cudaMemcpyHostToDevice, g->stream));
  if (g->nnz_off) CUDA_TRY(cudaMemcpyAsync(g->col_idx, g->col_idx_h.data(), (size_t)g->nnz_off * sizeof(int), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->active, g->active_h.data(), (size_t)N, cudaMemcpyHostToDevice, g->stream));
  {
    // unit column scaling = the mask of variable components: 0 for constant poses, and for the constant half of a pose
    // whose p (code 2) or q (code 3) alone is held constant -- those columns vanish from the problem
    std::vector<double> se((size_t)N * 6);
    for (int i = 0; i < N; ++i) {
      const int code = in.pose_const ? in.pose_const[i] : 0;
      for (int k = 0; k < 6; ++k) se[6 * (size_t)i + k] = (g->active_h[i] && !(code == 2 && k < 3) && !(code == 3 && k >= 3)) ? 1.0 : 0.0;
    }
    CUDA_TRY(cudaMemcpyAsync(g->scale_eval, se.data(), se.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(g->scale, se.data(), se.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  }
  // (pageable sources: cudaMemcpyAsync returns once they are staged, so the host vectors may go; the stream orders the rest)
  CUDA_TRY(cudaMemsetAsync(g->Hoff, 0, std::max<size_t>((size_t)g->nnz_off * 36, 1) * sizeof(double), g->stream));
  for (double** v : {&g->vx, &g->vr, &g->vu, &g->vw, &g->vp, &g->vs, &g->vb})
    CUDA_TRY(cudaMemsetAsync(*v, 0, std::max<size_t>((size_t)N * 6, 1) * sizeof(double), g->stream));

  lap("uploads + memset");
  // persistent PCG grid: all CTAs must be co-resident (cooperative launch)
  int per_sm = 0;
  if (!pool_cache_get(device, kCachePcgPerSm, &per_sm)) {
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_kernel<false>, kPcgThreads, 0));
    pool_cache_set(device, kCachePcgPerSm, per_sm);
  }
  g->pcg_max_ctas = std::max(1, std::min(per_sm, 4) * g->num_sms);
  PGO_TRY(dev_alloc(g, &g->partials, (size_t)2 * 3 * g->pcg_max_ctas));

  // [n_loc][7] host -> [n_loc][8] device
  g->pose_stage.resize((size_t)N * 8);
  for (int i = 0; i < N; ++i) {
    std::memcpy(&g->pose_stage[8 * (size_t)i], in.poses + 7 * (size_t)i, 7 * sizeof(double));
    g->pose_stage[8 * (size_t)i + 7] = 0.0;
  }
  CUDA_TRY(cudaMemcpyAsync(g->poses, g->pose_stage.data(), (size_t)N * 8 * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  lap("poses");
  g->setup_s = wall_s() - t0;
  return PGO_OK;
}

static int check_create_args(int device, int n_poses, int n_edges, const double* poses, const int* edge_ids, const double* edge_meas) {
  if (n_poses <= 0 || n_edges < 0 || !poses || (n_edges > 0 && (!edge_ids || !edge_meas)))
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_create: null or empty input");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(PGO_ERR_NO_DEVICE, "pgo_graph_create: no CUDA device available (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return set_error(PGO_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  return check_edges(n_poses, n_edges, edge_ids);
}

extern "C" int pgo_graph_create(pgo_graph** out, int device, int n_poses, int n_edges, const double* poses,
                                const int* edge_ids, const double* edge_meas, const double* edge_sqrt_info,
                                const unsigned char* pose_const) {
  if (!out) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_create: null output");
  PGO_TRY(check_create_args(device, n_poses, n_edges, poses, edge_ids, edge_meas));
  CUDA_TRY(cudaSetDevice(device));
  pgo_graph* g = new pgo_graph();
  g->device = device;
  g->N_global = n_poses; g->E_global = n_edges; g->g0 = 0;
  g->part_off = {0, n_poses};
  LocalProblem in;
  in.n_loc = in.n_own = n_poses; in.n_edges = n_edges; in.poses = poses; in.edge_ids = edge_ids; in.edge_meas = edge_meas;
  in.edge_sqrt_info = edge_sqrt_info; in.pose_const = pose_const;
  const int rc = graph_create_local(g, in);
  if (rc != PGO_OK) { pgo_graph_destroy(g); return rc; }
  // setup-time positions: the strength measure of the multilevel hierarchy (built on first use)
  g->pos0_h.resize(3 * (size_t)n_poses);
  for (int i = 0; i < n_poses; ++i) for (int k = 0; k < 3; ++k) g->pos0_h[3 * (size_t)i + k] = poses[7 * (size_t)i + k];
  *out = g;
  return PGO_OK;
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU: owner-computes row partition.  Every rank passes the SAME global graph; rank r keeps the block rows of
// the poses [off[r], off[r+1]) (contiguous index ranges: pose graphs are trajectory-ordered, so index locality is
// spatial locality), every edge that touches one of them, and halo copies of the other endpoints of cut edges.
// ------------------------------------------------------------------------------------------------
struct HostPartition {
  std::vector<int> off;                      // [world + 1]
  int n_own = 0, g0 = 0;
  std::vector<int> halo_gid;                 // ascending
  std::vector<int> edge_sel;                 // global edge ids kept here, ascending
  std::vector<int> local_ids;                // [2 * kept]
  std::vector<int> nbr, send_ptr, send_idx, recv_ptr;
};

static void partition_ranges(int n_poses, int world, std::vector<int>* off) {
  off->resize(world + 1);
  for (int r = 0; r <= world; ++r) (*off)[r] = (int)(((long long)n_poses * r) / world);
}

static void build_partition(int n_poses, int n_edges, const int* edge_ids, int rank, int world, HostPartition* P) {
  partition_ranges(n_poses, world, &P->off);
  const int lo = P->off[rank], hi = P->off[rank + 1];
  P->g0 = lo; P->n_own = hi - lo;
  P->edge_sel.clear(); P->halo_gid.clear();
  for (int e = 0; e < n_edges; ++e) {
    const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
    const bool oa = a >= lo && a < hi, ob = b >= lo && b < hi;
    if (!oa && !ob) continue;
    P->edge_sel.push_back(e);
    if (!oa) P->halo_gid.push_back(a);
    if (!ob) P->halo_gid.push_back(b);
  }
  std::sort(P->halo_gid.begin(), P->halo_gid.end());
  P->halo_gid.erase(std::unique(P->halo_gid.begin(), P->halo_gid.end()), P->halo_gid.end());
  auto local_of = [&](int gid) -> int {
    if (gid >= lo && gid < hi) return gid - lo;
    return P->n_own + (int)(std::lower_bound(P->halo_gid.begin(), P->halo_gid.end(), gid) - P->halo_gid.begin());
  };
  P->local_ids.resize(2 * P->edge_sel.size());
  for (size_t k = 0; k < P->edge_sel.size(); ++k) {
    const int e = P->edge_sel[k];
    P->local_ids[2 * k] = local_of(edge_ids[2 * e]);
    P->local_ids[2 * k + 1] = local_of(edge_ids[2 * e + 1]);
  }
  // exchange plan from the cut edges (symmetric by construction: a cut edge puts each endpoint into the other owner's halo)
  P->nbr.clear(); P->recv_ptr.clear();
  for (size_t k = 0; k < P->halo_gid.size(); ++k) {
    const int o = amg_owner_of(P->off, P->halo_gid[k]);
    if (P->nbr.empty() || P->nbr.back() != o) { P->nbr.push_back(o); P->recv_ptr.push_back((int)k); }
  }
  P->recv_ptr.push_back((int)P->halo_gid.size());
  std::vector<std::vector<int>> need(P->nbr.size());
  for (size_t k = 0; k < P->edge_sel.size(); ++k) {
    const int e = P->edge_sel[k];
    const int a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
    const bool oa = a >= lo && a < hi, ob = b >= lo && b < hi;
    if (oa == ob) continue;
    const int mine = oa ? a : b, other = oa ? b : a;
    const int o = amg_owner_of(P->off, other);
    const int q = (int)(std::lower_bound(P->nbr.begin(), P->nbr.end(), o) - P->nbr.begin());
    need[q].push_back(mine - lo);
  }
  P->send_ptr.assign(1, 0);
  P->send_idx.clear();
  for (size_t q = 0; q < P->nbr.size(); ++q) {
    std::sort(need[q].begin(), need[q].end());
    need[q].erase(std::unique(need[q].begin(), need[q].end()), need[q].end());
    P->send_idx.insert(P->send_idx.end(), need[q].begin(), need[q].end());
    P->send_ptr.push_back((int)P->send_idx.size());
  }
}

extern "C" int pgo_nccl_unique_id(unsigned char unique_id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_TRY(ncclGetUniqueId(&id));
  std::memcpy(unique_id, &id, 128);
  return PGO_OK;
}

extern "C" int pgo_graph_create_partitioned(pgo_graph** out, int device, int n_poses, int n_edges, const double* poses,
                                            const int* edge_ids, const double* edge_meas, const double* edge_sqrt_info,
                                            const unsigned char* pose_const, const unsigned char unique_id[128], int rank,
                                            int world_size) {
  if (!out || !unique_id || world_size < 1 || rank < 0 || rank >= world_size)
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_create_partitioned: bad arguments");
  if (world_size == 1) return pgo_graph_create(out, device, n_poses, n_edges, poses, edge_ids, edge_meas, edge_sqrt_info, pose_const);
  PGO_TRY(check_create_args(device, n_poses, n_edges, poses, edge_ids, edge_meas));
  if (n_poses < world_size) return set_error(PGO_ERR_INVALID_ARGUMENT, "fewer poses (%d) than ranks (%d)", n_poses, world_size);
  CUDA_TRY(cudaSetDevice(device));
  pgo_graph* g = new pgo_graph();
  g->device = device;
  g->rank = rank; g->world = world_size;
  g->N_global = n_poses; g->E_global = n_edges;
  auto fail = [&](int rc) { pgo_graph_destroy(g); return rc; };
  {
    ncclUniqueId id;
    std::memcpy(&id, unique_id, 128);
    const ncclResult_t e = ncclCommInitRank(&g->comm, world_size, id, rank);
    if (e != ncclSuccess) return fail(set_error(PGO_ERR_NCCL, "ncclCommInitRank failed: %s", ncclGetErrorString(e)));
  }
  HostPartition P;
  build_partition(n_poses, n_edges, edge_ids, rank, world_size, &P);
  g->part_off = P.off; g->g0 = P.g0; g->halo_gid = P.halo_gid; g->edge_gid = P.edge_sel;
  g->nbr = P.nbr; g->send_ptr = P.send_ptr; g->recv_ptr = P.recv_ptr; g->send_idx_h = P.send_idx;
  // local slice of the inputs
  const int n_loc = P.n_own + (int)P.halo_gid.size(), ne = (int)P.edge_sel.size();
  std::vector<double> lposes((size_t)n_loc * 7), lmeas((size_t)ne * 7), linfo;
  std::vector<unsigned char> lconst((size_t)n_loc, 0);
  auto gid_of = [&](int i) { return i < P.n_own ? P.g0 + i : P.halo_gid[i - P.n_own]; };
  for (int i = 0; i < n_loc; ++i) {
    const int gi = gid_of(i);
    std::memcpy(&lposes[7 * (size_t)i], poses + 7 * (size_t)gi, 7 * sizeof(double));
    if (pose_const) lconst[i] = pose_const[gi];
  }
  if (edge_sqrt_info) linfo.resize((size_t)ne * 36);
  for (int k = 0; k < ne; ++k) {
    const int e = P.edge_sel[k];
    std::memcpy(&lmeas[7 * (size_t)k], edge_meas + 7 * (size_t)e, 7 * sizeof(double));
    if (edge_sqrt_info) std::memcpy(&linfo[36 * (size_t)k], edge_sqrt_info + 36 * (size_t)e, 36 * sizeof(double));
  }
  LocalProblem in;
  in.n_loc = n_loc; in.n_own = P.n_own; in.n_edges = ne; in.poses = lposes.data(); in.edge_ids = P.local_ids.data();
  in.edge_meas = lmeas.data(); in.edge_sqrt_info = edge_sqrt_info ? linfo.data() : nullptr; in.pose_const = lconst.data();
  int rc = graph_create_local(g, in);
  if (rc != PGO_OK) return fail(rc);
  // global structure for the hierarchy: pattern and variable flags of the WHOLE graph (every rank, deterministic)
  {
    HostPattern gp;
    build_pattern(n_poses, n_edges, edge_ids, pose_const, &gp);
    g->gactive_h.swap(gp.active);
    g->grow_ptr_h.swap(gp.row_ptr);
    g->gcol_idx_h.swap(gp.col_idx);
    g->pos0_h.resize(3 * (size_t)n_poses);
    for (int i = 0; i < n_poses; ++i) for (int k = 0; k < 3; ++k) g->pos0_h[3 * (size_t)i + k] = poses[7 * (size_t)i + k];
  }
  size_t max_send = P.send_idx.size();
  if (!P.send_idx.empty()) {
    rc = dev_alloc(g, &g->send_idx, P.send_idx.size());
    if (rc != PGO_OK) return fail(rc);
    if (cudaMemcpyAsync(g->send_idx, P.send_idx.data(), P.send_idx.size() * sizeof(int), cudaMemcpyHostToDevice, g->stream) != cudaSuccess ||
        cudaStreamSynchronize(g->stream) != cudaSuccess)
      return fail(set_error(PGO_ERR_CUDA, "uploading the halo plan failed"));
  }
  rc = dev_alloc(g, &g->halo_sendbuf, std::max<size_t>(max_send, 1) * 8);
  if (rc != PGO_OK) return fail(rc);
  // a pose that is constant / unused on its owner must be inactive in every halo copy as well: owners are authoritative
  {
    std::vector<double> act((size_t)n_loc * 6, 0.0);
    for (int i = 0; i < n_loc; ++i) {
      const int gi = gid_of(i), code = lconst[i];
      g->active_h[i] = g->gactive_h[gi];
      if (g->gactive_h[gi]) for (int k = 0; k < 6; ++k) act[6 * (size_t)i + k] = (!(code == 2 && k < 3) && !(code == 3 && k >= 3)) ? 1.0 : 0.0;
    }
    if (cudaMemcpyAsync(g->scale_eval, act.data(), act.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream) != cudaSuccess ||
        cudaMemcpyAsync(g->scale, act.data(), act.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream) != cudaSuccess ||
        cudaMemcpyAsync(g->active, g->active_h.data(), (size_t)n_loc, cudaMemcpyHostToDevice, g->stream) != cudaSuccess ||
        cudaStreamSynchronize(g->stream) != cudaSuccess)
      return fail(set_error(PGO_ERR_CUDA, "uploading the variable flags failed"));
  }
  *out = g;
  return PGO_OK;
}

static unsigned long long mix64(unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  for (unsigned long long v : {a, b, c, d}) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 31; }
  return h;
}

// One rank's complete host-side setup from the global inputs (shared by pgo_analyze_partition; mirrors what
// pgo_graph_create_partitioned + the first multilevel solve build).
struct HostRankSetup {
  HostPartition part;
  HostPattern local_pat;
  std::vector<AmgLocalLevel> levels;
};
static void host_rank_setup(int n_poses, int n_edges, const double* poses, const int* edge_ids, const unsigned char* pose_const,
                            int rank, int world, const std::vector<AmgGlobalLevel>& G, HostRankSetup* out) {
  (void)poses;
  build_partition(n_poses, n_edges, edge_ids, rank, world, &out->part);
  const HostPartition& P = out->part;
  const int n_loc = P.n_own + (int)P.halo_gid.size();
  std::vector<unsigned char> lconst((size_t)n_loc, 0);
  for (int i = 0; i < n_loc; ++i) {
    const int gi = i < P.n_own ? P.g0 + i : P.halo_gid[i - P.n_own];
    if (pose_const) lconst[i] = pose_const[gi];
  }
  build_pattern(n_loc, (int)P.edge_sel.size(), P.local_ids.data(), lconst.data(), &out->local_pat, P.n_own);
  AmgLocalLevel l0;
  l0.row_ptr = out->local_pat.row_ptr; l0.col_idx = out->local_pat.col_idx; l0.halo_gid = P.halo_gid;
  l0.nbr = P.nbr; l0.send_ptr = P.send_ptr; l0.send_idx = P.send_idx; l0.recv_ptr = P.recv_ptr;
  amg_localize(G, rank, world, l0, &out->levels);
}

extern "C" int pgo_analyze_partition(int n_poses, int n_edges, const double* poses, const int* edge_ids,
                                     const unsigned char* pose_const, int rank, int world_size, pgo_partition_info* info) {
  if (!info || !poses || n_poses <= 0 || n_edges < 0 || (n_edges > 0 && !edge_ids) || world_size < 1 || world_size > 64 || rank < 0 ||
      rank >= world_size || n_poses < world_size)
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_analyze_partition: bad arguments");
  PGO_TRY(check_edges(n_poses, n_edges, edge_ids));
  std::memset(info, 0, sizeof *info);
  // global structure + hierarchy (what every rank computes redundantly)
  HostPattern gp;
  build_pattern(n_poses, n_edges, edge_ids, pose_const, &gp);
  std::vector<double> pos0(3 * (size_t)n_poses);
  for (int i = 0; i < n_poses; ++i) for (int k = 0; k < 3; ++k) pos0[3 * (size_t)i + k] = poses[7 * (size_t)i + k];
  std::vector<int> off;
  partition_ranges(n_poses, world_size, &off);
  const AmgHostParams prm = amg_host_params(n_poses);
  std::vector<AmgGlobalLevel> G;
  amg_build_global(n_poses, world_size, off, gp.active.data(), gp.row_ptr.data(), gp.col_idx.data(), pos0.data(), prm, &G);
  HostRankSetup me;
  host_rank_setup(n_poses, n_edges, poses, edge_ids, pose_const, rank, world_size, G, &me);
  const HostPartition& P = me.part;
  info->n_own = P.n_own; info->n_halo = (int)P.halo_gid.size(); info->n_local_edges = (int)P.edge_sel.size();
  for (size_t k = 0; k < P.edge_sel.size(); ++k)
    if ((P.local_ids[2 * k] >= P.n_own) != (P.local_ids[2 * k + 1] >= P.n_own)) info->n_cut_edges++;
  info->n_neighbours = (int)P.nbr.size();
  info->send_total = (int)P.send_idx.size(); info->recv_total = (int)P.halo_gid.size();
  for (size_t q = 0; q < P.nbr.size(); ++q) {
    info->send_to[P.nbr[q]] = P.send_ptr[q + 1] - P.send_ptr[q];
    info->recv_from[P.nbr[q]] = P.recv_ptr[q + 1] - P.recv_ptr[q];
  }
  const int nl = (int)me.levels.size();
  info->amg_levels = nl;
  bool ok = true;
  unsigned long long hs = 0, hr = 0;
  for (int l = 0; l < nl && l < 16; ++l) {
    const AmgLocalLevel& L = me.levels[l];
    info->level_nodes[l] = G[l].n; info->level_own[l] = L.n_own; info->level_halo[l] = L.n_halo;
    info->level_replicated[l] = L.replicated ? 1 : 0;
    info->level_blocks[l] = (long long)L.n_own + (long long)L.col_idx.size();
    info->level_send[l] = (int)L.send_idx.size(); info->level_recv[l] = L.replicated ? 0 : L.n_halo;
    for (size_t q = 0; q < L.nbr.size(); ++q) {
      for (int k = L.send_ptr[q]; k < L.send_ptr[q + 1]; ++k) hs += mix64(rank, L.nbr[q], (unsigned long long)l << 32 | (unsigned)(L.g0 + L.send_idx[k]), k - L.send_ptr[q]);
      for (int k = L.recv_ptr[q]; k < L.recv_ptr[q + 1]; ++k) hr += mix64(L.nbr[q], rank, (unsigned long long)l << 32 | (unsigned)L.halo_gid[k], k - L.recv_ptr[q]);
    }
    // ---- invariants ----
    // (peer-memory exchanges, pgo_peer.cuh: two staging buffers per channel are enough only when partners are symmetric)
    for (size_t q = 0; q < L.nbr.size(); ++q)
      if ((L.send_ptr[q + 1] > L.send_ptr[q]) != (L.recv_ptr[q + 1] > L.recv_ptr[q])) ok = false;
    const int n_loc = L.n_own + L.n_halo;
    for (size_t p = 0; p < L.col_idx.size(); ++p) if (L.col_idx[p] < 0 || L.col_idx[p] >= n_loc) ok = false;
    if (l + 1 < nl) {
      const AmgLocalLevel& C = me.levels[l + 1];
      if ((int)L.agg.size() != n_loc) ok = false;
      // aggregates never cross a rank boundary: the owner of a node owns its aggregate
      const AmgGlobalLevel& g = G[l];
      const AmgGlobalLevel& gc = G[l + 1];
      // (a replicated level lives whole on every rank: its aggregates are free to span the former ranges)
      for (int i = 0; i < g.n && ok && !g.replicated; ++i)
        if (g.agg[i] >= 0 && amg_owner_of(g.off, i) != amg_owner_of(gc.off, g.agg[i])) ok = false;
      // every stored fine block lands in exactly one gather list; member lists cover the stored variable rows
      long long want = 0, members = 0;
      for (int i = 0; i < L.n_own; ++i) {
        if (L.agg[i] < 0) continue;
        ++members; ++want;
        for (int p = L.row_ptr[i]; p < L.row_ptr[i + 1]; ++p) if (L.agg[L.col_idx[p]] >= 0) ++want;
      }
      if (want != (long long)L.gal_row.size() || members != (long long)L.mem_idx.size()) ok = false;
      if ((int)L.gal_ptr.size() - 1 != (L.c_row1 - L.c_row0) + (C.row_ptr[L.c_row1] - C.row_ptr[L.c_row0])) ok = false;
      for (int a : L.agg) if (a >= C.n_own + C.n_halo) ok = false;
    }
  }
  // ---- cross-rank: what I send to q is exactly what q expects from me, in the same order, on every distributed level ----
  if (world_size > 1 && world_size <= 16) {
    for (int q = 0; q < world_size; ++q) {
      if (q == rank) continue;
      HostRankSetup peer;
      host_rank_setup(n_poses, n_edges, poses, edge_ids, pose_const, q, world_size, G, &peer);
      for (int l = 0; l < nl; ++l) {
        const AmgLocalLevel& A = me.levels[l];
        const AmgLocalLevel& B = peer.levels[l];
        if (A.replicated) continue;
        std::vector<int> sent, expected;
        for (size_t k = 0; k < A.nbr.size(); ++k)
          if (A.nbr[k] == q) for (int t = A.send_ptr[k]; t < A.send_ptr[k + 1]; ++t) sent.push_back(A.g0 + A.send_idx[t]);
        for (size_t k = 0; k < B.nbr.size(); ++k)
          if (B.nbr[k] == rank) for (int t = B.recv_ptr[k]; t < B.recv_ptr[k + 1]; ++t) expected.push_back(B.halo_gid[t]);
        if (sent != expected) ok = false;
      }
    }
  }
  info->plan_checksum = hs; info->recv_checksum = hr;
  info->consistent = ok ? 1 : 0;
  return PGO_OK;
}

// Host-only: the aggregate of every node on every level of the global hierarchy (tests / design tools).
// agg_out: concatenation over the levels 0..L-2 of agg[level_nodes[l]]; returns the number of levels in *n_levels.
extern "C" int pgo_amg_aggregates(int n_poses, int n_edges, const double* poses, const int* edge_ids, const unsigned char* pose_const,
                                  int world_size, int* n_levels, int* level_nodes /* [16] */, int* agg_out, long long agg_capacity) {
  if (!poses || !n_levels || !level_nodes || n_poses <= 0 || n_edges < 0 || (n_edges > 0 && !edge_ids) || world_size < 1)
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_amg_aggregates: bad arguments");
  PGO_TRY(check_edges(n_poses, n_edges, edge_ids));
  HostPattern gp;
  build_pattern(n_poses, n_edges, edge_ids, pose_const, &gp);
  std::vector<double> pos0(3 * (size_t)n_poses);
  for (int i = 0; i < n_poses; ++i) for (int k = 0; k < 3; ++k) pos0[3 * (size_t)i + k] = poses[7 * (size_t)i + k];
  std::vector<int> off;
  partition_ranges(n_poses, world_size, &off);
  const AmgHostParams prm = amg_host_params(n_poses);
  std::vector<AmgGlobalLevel> G;
  amg_build_global(n_poses, world_size, off, gp.active.data(), gp.row_ptr.data(), gp.col_idx.data(), pos0.data(), prm, &G);
  *n_levels = (int)G.size();
  long long need = 0;
  for (size_t l = 0; l < G.size() && l < 16; ++l) { level_nodes[l] = G[l].n; if (l + 1 < G.size()) need += G[l].n; }
  if (agg_out) {
    if (agg_capacity < need) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_amg_aggregates: capacity %lld < %lld", agg_capacity, need);
    long long k = 0;
    for (size_t l = 0; l + 1 < G.size(); ++l) for (int i = 0; i < G[l].n; ++i) agg_out[k++] = G[l].agg[i];
  }
  return PGO_OK;
}

extern "C" int pgo_graph_rank(const pgo_graph* g) { return g ? g->rank : 0; }
extern "C" int pgo_graph_world_size(const pgo_graph* g) { return g ? g->world : 0; }
extern "C" int pgo_graph_num_local_poses(const pgo_graph* g) { return g ? g->n_own : 0; }
extern "C" int pgo_graph_num_halo_poses(const pgo_graph* g) { return g ? g->N - g->n_own : 0; }
extern "C" int pgo_graph_num_local_edges(const pgo_graph* g) { return g ? g->E : 0; }

extern "C" int pgo_graph_set_stream(pgo_graph* g, void* cuda_stream) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  g->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : g->own_stream;
  return PGO_OK;
}

extern "C" int pgo_graph_num_poses(const pgo_graph* g) { return g ? g->N_global : 0; }
extern "C" int pgo_graph_num_edges(const pgo_graph* g) { return g ? g->E_global : 0; }

// poses: the GLOBAL [n_poses][7] array on every rank
extern "C" int pgo_graph_set_poses(pgo_graph* g, const double* poses) {
  if (!g || !poses) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_set_poses: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  // [N][7] host -> [N][8] device: pad on the host, one contiguous copy (row-pitched DMA of 56-byte rows is slow)
  g->pose_stage.resize((size_t)g->N * 8);
  for (int i = 0; i < g->N; ++i) {
    const int gi = i < g->n_own ? g->g0 + i : g->halo_gid[i - g->n_own];
    std::memcpy(&g->pose_stage[8 * (size_t)i], poses + 7 * (size_t)gi, 7 * sizeof(double));
    g->pose_stage[8 * (size_t)i + 7] = 0.0;
  }
  CUDA_TRY(cudaMemcpyAsync(g->poses, g->pose_stage.data(), (size_t)g->N * 8 * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PGO_OK;
}

// poses: the GLOBAL [n_poses][7] array; multi-GPU: collective (all-gather of the owned slices), every rank gets all poses
extern "C" int pgo_graph_get_poses(pgo_graph* g, double* poses) {
  if (!g || !poses) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_get_poses: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  const int Ng = g->N_global;
  g->pose_stage.resize((size_t)Ng * 8);
  const double* src = g->poses;
  if (g->world > 1) {
    if (!g->gather_buf) PGO_TRY(dev_alloc(g, &g->gather_buf, (size_t)Ng * 8));
    CUDA_TRY(cudaMemcpyAsync(g->gather_buf + 8 * (size_t)g->g0, g->poses, (size_t)g->n_own * 8 * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
    NCCL_TRY(ncclGroupStart());
    for (int r = 0; r < g->world; ++r) {
      const size_t cnt = (size_t)(g->part_off[r + 1] - g->part_off[r]) * 8;
      double* p = g->gather_buf + 8 * (size_t)g->part_off[r];
      if (cnt) NCCL_TRY(ncclBroadcast(p, p, cnt, ncclDouble, r, g->comm, g->stream));
    }
    NCCL_TRY(ncclGroupEnd());
    src = g->gather_buf;
  }
  CUDA_TRY(cudaMemcpyAsync(g->pose_stage.data(), src, (size_t)Ng * 8 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  for (int i = 0; i < Ng; ++i) std::memcpy(poses + 7 * (size_t)i, &g->pose_stage[8 * (size_t)i], 7 * sizeof(double));
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
// halo exchange (multi-GPU): pack the owned boundary entries, NCCL send / receive straight into the halo tail
// ------------------------------------------------------------------------------------------------
namespace pgo {
__global__ void halo_pack_kernel(int n_send, int width, const int* __restrict__ send_idx, const double* __restrict__ v,
                                 double* __restrict__ buf, const int* skip) {
  if (skip && *skip) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = t / width, c = t - k * width;
  if (k >= n_send) return;
  buf[(size_t)k * width + c] = v[(size_t)send_idx[k] * width + c];
}
}  // namespace pgo

static int halo_exchange(pgo_graph* g, const std::vector<int>& nbr, const std::vector<int>& send_ptr, const std::vector<int>& recv_ptr,
                         const int* send_idx, int n_own, double* v, int width, const int* skip) {
  if (g->world <= 1 || nbr.empty()) return PGO_OK;
  const int n_send = send_ptr.back();
  if (n_send > 0) {
    const int items = n_send * width;
    halo_pack_kernel<<<(items + 255) / 256, 256, 0, g->stream>>>(n_send, width, send_idx, v, g->halo_sendbuf, skip);
    g->launches++;
  }
  NCCL_TRY(ncclGroupStart());
  for (size_t k = 0; k < nbr.size(); ++k) {
    const int ns = send_ptr[k + 1] - send_ptr[k], nr = recv_ptr[k + 1] - recv_ptr[k];
    if (ns > 0) NCCL_TRY(ncclSend(g->halo_sendbuf + (size_t)send_ptr[k] * width, (size_t)ns * width, ncclDouble, nbr[k], g->comm, g->stream));
    if (nr > 0) NCCL_TRY(ncclRecv(v + ((size_t)n_own + recv_ptr[k]) * width, (size_t)nr * width, ncclDouble, nbr[k], g->comm, g->stream));
  }
  NCCL_TRY(ncclGroupEnd());
  g->comm_calls++;
  g->comm_bytes += (long long)n_send * width * 8;
  return PGO_OK;
}
// level-0 vectors / poses / scales
static int halo_exchange0(pgo_graph* g, double* v, int width, const int* skip = nullptr) {
  return halo_exchange(g, g->nbr, g->send_ptr, g->recv_ptr, g->send_idx, g->n_own, v, width, skip);
}

static int allreduce_sum(pgo_graph* g, double* buf, size_t count) {
  if (g->world <= 1) return PGO_OK;
  NCCL_TRY(ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, g->comm, g->stream));
  g->comm_calls++;
  g->comm_bytes += (long long)count * 8;
  return PGO_OK;
}

// host [N_global][width] -> device local layout (owned rows, then the halo copies)
static int upload_global(pgo_graph* g, const double* src, int width, double* dst_dev, bool with_halo) {
  if (g->world <= 1) {
    CUDA_TRY(cudaMemcpyAsync(dst_dev, src, (size_t)g->N * width * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    return PGO_OK;
  }
  CUDA_TRY(cudaMemcpyAsync(dst_dev, src + (size_t)g->g0 * width, (size_t)g->n_own * width * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  if (with_halo && g->N > g->n_own) {
    std::vector<double> h((size_t)(g->N - g->n_own) * width);
    for (int k = 0; k < g->N - g->n_own; ++k) std::memcpy(&h[(size_t)k * width], src + (size_t)g->halo_gid[k] * width, width * sizeof(double));
    CUDA_TRY(cudaMemcpyAsync(dst_dev + (size_t)g->n_own * width, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));   // `h` is pageable and goes out of scope
  }
  return PGO_OK;
}
// device owned rows [n_own][width] -> host [N_global][width] on every rank (collective when world > 1)
static int download_global(pgo_graph* g, const double* src_dev, int width, double* dst) {
  if (g->world <= 1) {
    CUDA_TRY(cudaMemcpyAsync(dst, src_dev, (size_t)g->N * width * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    return PGO_OK;
  }
  if (width > 8) return set_error(PGO_ERR_INVALID_ARGUMENT, "download_global: width");
  if (!g->gather_buf) PGO_TRY(dev_alloc(g, &g->gather_buf, (size_t)g->N_global * 8));
  CUDA_TRY(cudaMemcpyAsync(g->gather_buf + (size_t)width * g->g0, src_dev, (size_t)g->n_own * width * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
  NCCL_TRY(ncclGroupStart());
  for (int r = 0; r < g->world; ++r) {
    const size_t cnt = (size_t)(g->part_off[r + 1] - g->part_off[r]) * width;
    double* p = g->gather_buf + (size_t)width * g->part_off[r];
    if (cnt) NCCL_TRY(ncclBroadcast(p, p, cnt, ncclDouble, r, g->comm, g->stream));
  }
  NCCL_TRY(ncclGroupEnd());
  CUDA_TRY(cudaMemcpyAsync(dst, g->gather_buf, (size_t)g->N_global * width * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PGO_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel launch helpers
// ------------------------------------------------------------------------------------------------
template <bool kIdent, int kMode, int kMinBlocks>
static int launch_linearize_t(pgo_graph* g, const LinParams& p) {
  constexpr int smem = lin_smem_bytes<kIdent>();
  auto kern = linearize_kernel<kIdent, kMode, kMinBlocks>;
  // the attribute belongs to the device that was current when it was set: remember it per device, not per process
  constexpr int key = kCacheLinAttrBase + (kIdent ? 0 : 8) + (kMode == kLinFull ? 0 : (kMode == kLinCost ? 2 : 4)) + (kMinBlocks >= 3 ? 1 : 0);
  int have = 0;
  if (!pool_cache_get(g->device, key, &have)) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pool_cache_set(g->device, key, 1);
  }
  const int ctas = std::max(1, std::min((g->T + kLinWarps - 1) / kLinWarps, kMinBlocks * g->num_sms));
  kern<<<ctas, kLinWarps * 32, smem, g->stream>>>(p);
  g->launches++;
  CUDA_TRY(cudaGetLastError());
  return PGO_OK;
}

struct LinTarget { double* Hdiag; double* Hoff; double* grad; };   // the system a linearisation accumulates into
static int launch_linearize(pgo_graph* g, int mode, const double* poses, const double* scale, int loss_type,
                            double loss_a, double* res_out = nullptr, double* jac_out = nullptr, const LinTarget* target = nullptr,
                            const LmState* lm = nullptr) {
  LinParams p;
  p.n_edges = g->E; p.n_tiles = g->T; p.n_own = g->n_own; p.core = g->core; p.info = g->info; p.poses = poses; p.scale = scale;
  p.Hdiag = target ? target->Hdiag : g->Hdiag; p.Hoff = target ? target->Hoff : g->Hoff; p.grad = target ? target->grad : g->grad;
  p.scalars = g->scalars; p.lm = lm; p.edge_loss = g->edge_loss_set ? g->edge_loss : nullptr;
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
    CUDA_TRY(cudaMemsetAsync(g->Hdiag, 0, (size_t)g->n_own * 36 * sizeof(double), g->stream));
    if (g->has_dup_blocks && g->nnz_off) CUDA_TRY(cudaMemsetAsync(g->Hoff, 0, (size_t)g->nnz_off * 36 * sizeof(double), g->stream));
  }
  CUDA_TRY(cudaMemsetAsync(g->grad, 0, (size_t)g->n_own * 6 * sizeof(double), g->stream));
  return PGO_OK;
}

static int zero_scalars(pgo_graph* g) {
  CUDA_TRY(cudaMemsetAsync(g->scalars, 0, sizeof(DeviceScalars), g->stream));
  return PGO_OK;
}
// multi-GPU: every rank reduced over its own rows / edges; the all-reduced values are bit-identical on every rank, so all
// ranks take the same LM decisions.  (The PCG statistics already are global.)
static int reduce_scalars(pgo_graph* g) {
  if (g->world > 1) {
    static_assert(offsetof(DeviceScalars, step_norm2) == 8 && offsetof(DeviceScalars, x_norm2) == 16, "cost, step_norm2, x_norm2 are contiguous");
    NCCL_TRY(ncclAllReduce(&g->scalars->cost, &g->scalars->cost, 3, ncclDouble, ncclSum, g->comm, g->stream));
    NCCL_TRY(ncclAllReduce(&g->scalars->gnorm2, &g->scalars->gnorm2, 1, ncclDouble, ncclSum, g->comm, g->stream));
    NCCL_TRY(ncclAllReduce(&g->scalars->gmax_bits, &g->scalars->gmax_bits, 1, ncclUint64, ncclMax, g->comm, g->stream));
    g->comm_calls += 3; g->comm_bytes += 40;
  }
  return PGO_OK;
}
static int fetch_scalars(pgo_graph* g) {
  PGO_TRY(reduce_scalars(g));
  CUDA_TRY(cudaMemcpyAsync(g->scalars_h, g->scalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return PGO_OK;
}

// full linearization at `poses` with column scaling `scale`: H, g of the owned block rows (multi-GPU: cut edges are
// evaluated by both owners, each keeping its own rows -- no exchange; the halo tails of poses and scale must be current)
static int linearize_full(pgo_graph* g, const double* poses, const double* scale, int loss_type, double loss_a) {
  PGO_TRY(zero_system(g, true));
  PGO_TRY(launch_linearize(g, kLinFull, poses, scale, loss_type, loss_a));
  return PGO_OK;
}

static BsrView bsr_view(const pgo_graph* g) {
  BsrView A;
  A.n = g->n_own; A.Hdiag = g->Hdiag; A.Hoff = g->Hoff; A.row_ptr = g->row_ptr; A.col_idx = g->col_idx;
  return A;
}

constexpr int kStreamPcgMinPoses = 200000;
constexpr int kAmgMinPoses = 512;          // PGO_LINEAR_AUTO: below this a mesh-like graph stays with block-Jacobi PCG
constexpr int kClusterPcgMaxPoses = 400;   // measured: at 2500 poses the 16-SM cluster is already 2.5x slower than the full grid

static int pcg_grid(const pgo_graph* g, const pgo_solver_options* o) {
  int want = (g->N + (kPcgThreads / 32) * kRowsPerWarp - 1) / ((kPcgThreads / 32) * kRowsPerWarp);
  if (o && o->pcg_num_ctas > 0) want = o->pcg_num_ctas;
  return std::max(1, std::min(want, g->pcg_max_ctas));
}

// Single-GPU persistent PCG: (H + diag(dlm)) x = b, x in g->vx. Results land in g->scalars.
static int launch_pcg(pgo_graph* g, const pgo_solver_options* o, const double* b, const LmState* lm = nullptr) {
  PcgParams P;
  P.lm = lm;
  P.A = bsr_view(g);
  P.d = g->dlm; P.Minv = g->Minv; P.b = b;
  P.x = g->vx; P.r = g->vr; P.u = g->vu; P.w = g->vw; P.p = g->vp; P.s = g->vs;
  P.partials = g->partials; P.barrier = g->barrier; P.scalars = g->scalars;
  P.max_iterations = o->pcg_max_iterations; P.tolerance = o->pcg_tolerance;
  CUDA_TRY(cudaMemsetAsync(g->barrier, 0, 4 * sizeof(unsigned int), g->stream));
  // small graphs: one 16-CTA cluster, hardware barrier (an iteration is a few microseconds, the barrier dominates)
  int cluster_ctas = 0;
  if (!pool_cache_get(g->device, kCachePcgCluster, &cluster_ctas)) {
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
    pool_cache_set(g->device, kCachePcgCluster, cluster_ctas);
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

// The multilevel hierarchy of this graph: global aggregation (every rank computes the same one), then this rank's slice.
static int graph_amg_hierarchy(pgo_graph* g, const AmgHostParams& prm, std::vector<AmgGlobalLevel>* G, std::vector<AmgLocalLevel>* L) {
  AmgLocalLevel l0;
  l0.row_ptr = g->row_ptr_h; l0.col_idx = g->col_idx_h; l0.halo_gid = g->halo_gid;
  l0.nbr = g->nbr; l0.send_ptr = g->send_ptr; l0.send_idx = g->send_idx_h; l0.recv_ptr = g->recv_ptr;
  if ((size_t)g->N_global * 3 != g->pos0_h.size()) return set_error(PGO_ERR_NUMERICAL, "amg: setup-time positions are missing");
  if (g->world == 1) {
    amg_build_global(g->N, 1, g->part_off, g->active_h.data(), g->row_ptr_h.data(), g->col_idx_h.data(), g->pos0_h.data(), prm, G);
  } else {
    amg_build_global(g->N_global, g->world, g->part_off, g->gactive_h.data(), g->grow_ptr_h.data(), g->gcol_idx_h.data(),
                     g->pos0_h.data(), prm, G);
    std::vector<int>().swap(g->grow_ptr_h);
    std::vector<int>().swap(g->gcol_idx_h);
  }
  amg_localize(*G, g->rank, g->world, l0, L);
  return PGO_OK;
}
#include "pgo_peer.cuh"
#include "pgo_amg.cuh"

// ------------------------------------------------------------------------------------------------
// C-ABI: evaluate / linearize / hessian / spmv / linear solve
// ------------------------------------------------------------------------------------------------
extern "C" int pgo_graph_evaluate(pgo_graph* g, int loss_type, double loss_a, double* cost, double* residuals,
                                  double* gradient, double* jacobians) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  if (g->world > 1 && (residuals || jacobians))
    return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_evaluate: per-edge outputs are not available on a partitioned graph (cost and gradient are)");
  double *res_d = nullptr, *jac_d = nullptr;
  const size_t res_bytes = std::max<size_t>((size_t)g->E * 6, 1) * sizeof(double), jac_bytes = std::max<size_t>((size_t)g->E * 72, 1) * sizeof(double);
  CUDA_TRY(pool_alloc(g->device, reinterpret_cast<void**>(&res_d), res_bytes));
  {
    const cudaError_t e2 = pool_alloc(g->device, reinterpret_cast<void**>(&jac_d), jac_bytes);
    if (e2 != cudaSuccess) {
      pool_free(g->device, res_d, res_bytes);
      return set_error(PGO_ERR_CUDA, "pgo_graph_evaluate: allocating %zu bytes failed: %s", jac_bytes, cudaGetErrorString(e2));
    }
  }
  int rc = PGO_OK;
  do {
    if ((rc = zero_system(g, false)) != PGO_OK) break;
    if ((rc = zero_scalars(g)) != PGO_OK) break;
    if ((rc = launch_linearize(g, kLinEval, g->poses, g->scale_eval, loss_type, loss_a, res_d, jac_d)) != PGO_OK) break;
    if ((rc = fetch_scalars(g)) != PGO_OK) break;
    if (cost) *cost = g->scalars_h->cost;
    cudaError_t ce = cudaSuccess;
    // the kernel writes per-edge outputs in processing order; hand them back in the caller's edge order
    auto fetch = [&](double* dst, const double* src_d, int width) {
      if (!g->edges_reordered) return cudaMemcpy(dst, src_d, (size_t)g->E * width * sizeof(double), cudaMemcpyDeviceToHost);
      std::vector<double> tmp((size_t)g->E * width);
      const cudaError_t e2 = cudaMemcpy(tmp.data(), src_d, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost);
      if (e2 != cudaSuccess) return e2;
      for (int e = 0; e < g->E; ++e) std::memcpy(dst + (size_t)e * width, tmp.data() + (size_t)g->edge_pos[e] * width, width * sizeof(double));
      return cudaSuccess;
    };
    if (residuals && g->E) ce = fetch(residuals, res_d, 6);
    if (ce == cudaSuccess && jacobians && g->E) ce = fetch(jacobians, jac_d, 72);
    if (ce != cudaSuccess) rc = set_error(PGO_ERR_CUDA, "evaluate copy-back failed: %s", cudaGetErrorString(ce));
    if (rc == PGO_OK && gradient) rc = download_global(g, g->grad, 6, gradient);
  } while (0);
  cudaStreamSynchronize(g->stream);
  pool_free(g->device, res_d, res_bytes); pool_free(g->device, jac_d, jac_bytes);
  return rc;
}

extern "C" int pgo_graph_set_edge_losses(pgo_graph* g, const int* loss_type, const double* loss_a) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  if (!loss_type || !loss_a) { g->edge_loss_set = false; return PGO_OK; }
  std::vector<double> code((size_t)std::max(g->E, 1), 0.0);
  for (int e = 0; e < g->E; ++e) {
    const int ge = g->edge_gid.empty() ? e : g->edge_gid[e];
    const int t = loss_type[ge];
    const double a = loss_a[ge];
    if (t != PGO_LOSS_TRIVIAL && t != PGO_LOSS_HUBER && t != PGO_LOSS_CAUCHY) return set_error(PGO_ERR_INVALID_ARGUMENT, "edge %d: unknown loss type %d", ge, t);
    if (t != PGO_LOSS_TRIVIAL && !(a > 0.0)) return set_error(PGO_ERR_INVALID_ARGUMENT, "edge %d: loss scale must be positive", ge);
    code[g->edge_pos[e]] = t == PGO_LOSS_HUBER ? a : (t == PGO_LOSS_CAUCHY ? -a : 0.0);
  }
  if (!g->edge_loss) PGO_TRY(dev_alloc(g, &g->edge_loss, (size_t)std::max(g->E, 1)));
  CUDA_TRY(cudaMemcpyAsync(g->edge_loss, code.data(), code.size() * sizeof(double), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  g->edge_loss_set = true;
  return PGO_OK;
}

extern "C" int pgo_graph_linearize(pgo_graph* g, int loss_type, double loss_a, const double* scale, double* cost,
                                   float* elapsed_ms) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  if (scale) PGO_TRY(upload_global(g, scale, 6, g->scale, true));
  else CUDA_TRY(cudaMemcpyAsync(g->scale, g->scale_eval, (size_t)g->N * 6 * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
  PGO_TRY(zero_system(g, true));
  PGO_TRY(zero_scalars(g));
  CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
  PGO_TRY(launch_linearize(g, kLinFull, g->poses, g->scale, loss_type, loss_a));
  CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
  PGO_TRY(fetch_scalars(g));
  if (cost) *cost = g->scalars_h->cost;
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, g->ev0, g->ev1));
  return PGO_OK;
}

extern "C" int pgo_graph_get_hessian(pgo_graph* g, long long* nnzb, int* row_ptr, int* col_idx, double* values,
                                     double* gradient) {
  if (!g) return set_error(PGO_ERR_INVALID_ARGUMENT, "null graph");
  CUDA_TRY(cudaSetDevice(g->device));
  if (g->world > 1) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_get_hessian: not available on a partitioned graph");
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
  // x, d, y are GLOBAL arrays; multi-GPU: the owned slice of x goes up and the halo tail arrives by the exchange
  PGO_TRY(upload_global(g, x, 6, g->vu, false));
  if (d) PGO_TRY(upload_global(g, d, 6, g->dlm, false));
  const int warps = (g->n_own + kRowsPerWarp - 1) / kRowsPerWarp;
  const int ctas = std::max(1, std::min((warps + 7) / 8, 8 * g->num_sms));
  CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
  for (int k = 0; k < std::max(repeats, 1); ++k) {
    PGO_TRY(halo_exchange0(g, g->vu, 6));
    spmv_kernel<false><<<ctas, 256, 0, g->stream>>>(bsr_view(g), g->vu, d ? g->dlm : nullptr, g->vw, true);
    g->launches++;
  }
  CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
  CUDA_TRY(cudaGetLastError());
  PGO_TRY(download_global(g, g->vw, 6, y));
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, g->ev0, g->ev1));
  return PGO_OK;
}

// Host-only symbolic analysis for solver type t.  AUTO takes the factor only when it is cheap: chain-like graphs (<= 64
// levels, node degree <= 16, fill <= 8x); mesh-like graphs (sphere, grids, dense random loops) go to block-Jacobi PCG.
static int run_symbolic(const pgo_graph* g, int t, LevelCholSymbolic* S) {
  const bool autosel = t == PGO_LINEAR_AUTO;
  const int rc = level_chol_symbolic(S, g->N, g->active_h.data(), g->row_ptr_h.data(), g->col_idx_h.data(), autosel ? 8.0 : 1e30,
                                     autosel ? 64 : 8192, autosel ? 16 : (1 << 30));
  if (rc != PGO_OK) S->error = g_last_error;     // carried to the thread that asked (this may be the helper thread)
  return rc;
}

static int resolve_linear_solver(pgo_graph* g, const pgo_solver_options* o) {
  int t = o->linear_solver_type;
  auto want_amg = [&]() -> int {
    if (!g->amg) {
      const int rc = amg_create(g, &g->amg);
      if (rc != PGO_OK) { amg_destroy(g->amg, g->device); g->amg = nullptr; return rc; }
    }
    return PGO_LINEAR_PCG_AMG;
  };
  if (g->world > 1) return want_amg();        // the one solver that follows the row partition (the factor is not distributed)
  if (t == PGO_LINEAR_PCG_AMG) return want_amg();
  if (t == PGO_LINEAR_AUTO || t == PGO_LINEAR_PCG_LEVEL_CHOLESKY) {
    if (!g->chol) {
      LevelChol* c = nullptr;
      int rc = PGO_OK;
      if (g->sym_future.valid() && g->sym_solver_type == t) {
        rc = g->sym_future.get();                      // started by pgo_solve_pose_graph (helper thread)
        if (rc != PGO_OK && g->sym && !g->sym->error.empty()) set_error(rc, "%s", g->sym->error.c_str());
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
    // mesh-like graph (no cheap exact factor): the multilevel preconditioner; tiny graphs stay with block-Jacobi
    if (g->n_own >= kAmgMinPoses) return want_amg();
    return PGO_LINEAR_PCG_BLOCK_JACOBI;
  }
  return PGO_LINEAR_PCG_BLOCK_JACOBI;
}

// LevenbergMarquardtStrategy::ComputeStep: D = diagonal / radius (new, reused or given), then solve
// (H + D) x = b with the chosen solver; x -> g->vx, stats -> g->scalars (after fetch).
// lm_state != nullptr: inside the device-resident LM loop -- radius and diagonal reuse are read from it on the device.
static int linear_solve_device(pgo_graph* g, const pgo_solver_options* o, int solver, const double* b, const LmDiagonal& lm,
                               const LmState* lm_state = nullptr) {
  if (g->world == 1 && solver == PGO_LINEAR_PCG_LEVEL_CHOLESKY) {
    // the LM diagonal is formed inside the factor kernel
    return level_chol_solve(g->chol, bsr_view(g), lm, lm_state, g->active, b, g->vx, g->vr, g->vu, g->vw, g->vp, g->vs,
                            std::min(o->pcg_max_iterations, 200), o->pcg_tolerance, std::max(o->pcg_tolerance, o->direct_residual_accept),
                            o->pcg_num_ctas, g->scalars,
                            g->stream, &g->launches);
  }
  const int tpb = 128;
  lm_prepare_kernel<<<(g->n_own + tpb - 1) / tpb, tpb, 0, g->stream>>>(g->n_own, g->Hdiag, g->active, lm.mode, lm.min_diag, lm.max_diag,
                                                                       lm.radius, lm.diagonal, lm.dlm, g->Minv, lm_state);
  g->launches++;
  if (solver == PGO_LINEAR_PCG_AMG) return amg_pcg_solve(g, o, b);
  static const bool force_stream_pcg = getenv("PGO_FORCE_STREAM_PCG") != nullptr;   // tests: the stream-ordered form on a small graph
  // large graphs: separate launches at full occupancy beat the persistent kernel (whose grid barriers only pay when an
  // iteration is a few microseconds long)
  if (force_stream_pcg || g->N >= kStreamPcgMinPoses) return pcg_multi(g, o, b);
  return launch_pcg(g, o, b, lm_state);
}
// does the solver poll the device from the host (its own convergence loop)?  Then the LM loop cannot run ahead of it.
static bool solver_is_host_polled(const pgo_graph* g, int solver) {
  static const bool force_stream_pcg = getenv("PGO_FORCE_STREAM_PCG") != nullptr;
  if (solver == PGO_LINEAR_PCG_AMG) return true;
  if (solver == PGO_LINEAR_PCG_BLOCK_JACOBI && (force_stream_pcg || g->N >= kStreamPcgMinPoses)) return true;
  return false;
}

extern "C" int pgo_graph_linear_solve(pgo_graph* g, const pgo_solver_options* options, const double* d,
                                      const double* b, double* y, int* iterations, double* relative_residual,
                                      float* elapsed_ms) {
  if (!g || !options || !d || !b || !y) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_linear_solve: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  const int solver = resolve_linear_solver(g, options);
  if (solver < 0) return solver;
  PGO_TRY(upload_global(g, d, 6, g->dlm, false));
  PGO_TRY(upload_global(g, b, 6, g->vb, false));
  PGO_TRY(zero_scalars(g));
  CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
  const LmDiagonal lm = {2, 0.0, 0.0, 1.0, g->diagonal, g->dlm};
  PGO_TRY(linear_solve_device(g, options, solver, g->vb, lm));
  CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->scalars_h, g->scalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  PGO_TRY(download_global(g, g->vx, 6, y));
  if (iterations) *iterations = g->scalars_h->pcg_iterations;
  if (relative_residual) *relative_residual = g->scalars_h->pcg_gamma0 > 0 ? std::sqrt(g->scalars_h->pcg_gamma / g->scalars_h->pcg_gamma0) : 0.0;
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, g->ev0, g->ev1));
  return PGO_OK;
}

// ------------------------------------------------------------------------------------------------
// ceres::Solve: TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy (ceres 1.13 flow), the same sequence of
// decisions as oracle/pgo_oracle.c:oracle_solve -- taken ON THE DEVICE (pgo_lm.cuh).  One iteration is: linear solve,
// candidate = Plus(x, -y .* scale), speculative linearisation at the candidate into the alternate system (cost, H, g and
// the gradient norms arrive with the step statistics), decision kernel, commit kernel.  The host enqueues iterations
// ahead of the GPU and watches the `done` flag two iterations behind, so no iteration waits for a host round trip;
// iterations enqueued past the end return at once.  (Linear solvers with a host-polled inner loop -- the multilevel and
// the stream-ordered PCG of large graphs -- keep the LM loop in step with the host: there an iteration is milliseconds.)
// ------------------------------------------------------------------------------------------------
static void lm_message(const LmState& st, const pgo_solver_options* opt, char* out, size_t cap) {
  switch (st.reason) {
    case kLmMaxIterations: snprintf(out, cap, "Maximum number of iterations reached. Number of iterations: %d.", (int)st.reason_value); break;
    case kLmGradientTolerance: snprintf(out, cap, "Gradient tolerance reached. Gradient max norm: %e <= %e", st.reason_value, opt->gradient_tolerance); break;
    case kLmMinRadius: snprintf(out, cap, "Minimum trust region radius reached."); break;
    case kLmInvalidSteps: snprintf(out, cap, "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %d", opt->max_num_consecutive_invalid_steps); break;
    case kLmParameterTolerance: snprintf(out, cap, "Parameter tolerance reached. Relative step_norm: %e <= %e.", st.reason_value, opt->parameter_tolerance); break;
    case kLmFunctionTolerance: snprintf(out, cap, "Function tolerance reached. |cost_change|/cost: %e <= %e", st.reason_value, opt->function_tolerance); break;
    case kLmInitialCostNotFinite: snprintf(out, cap, "Initial cost is not finite."); break;
    default: snprintf(out, cap, "The minimizer did not terminate."); break;
  }
}

static int graph_solve_impl(pgo_graph* g, const pgo_solver_options* opt, pgo_solver_summary* summary, pgo_iteration_summary* log, int log_cap);
extern "C" int pgo_graph_solve(pgo_graph* g, const pgo_solver_options* opt, pgo_solver_summary* summary,
                               pgo_iteration_summary* log, int log_cap) {
  const int rc = graph_solve_impl(g, opt, summary, log, log_cap);
  if (rc != PGO_OK && g && g->world > 1) g->failed = true;   // pgo_graph_destroy then aborts the communicator (peers may be inside a collective)
  return rc;
}
static int graph_solve_impl(pgo_graph* g, const pgo_solver_options* opt, pgo_solver_summary* summary, pgo_iteration_summary* log, int log_cap) {
  if (!g || !opt || !summary) return set_error(PGO_ERR_INVALID_ARGUMENT, "pgo_graph_solve: null argument");
  CUDA_TRY(cudaSetDevice(g->device));
  const double t_begin = wall_s();
  std::memset(summary, 0, sizeof *summary);
  const long long launches0 = g->launches;
  const int N = g->N;            // local poses incl. halo copies (pose / scale arrays)
  const int R = g->n_own;        // block rows = variables stored here
  const int tpb = 128, nblk = (R + tpb - 1) / tpb;
  const long long comm_calls0 = g->comm_calls, comm_bytes0 = g->comm_bytes;
  const long long peer_x0 = amg_peer_pushes(g), peer_b0 = amg_peer_bytes(g);
  const int solver = resolve_linear_solver(g, opt);
  if (solver < 0) return solver;
  summary->linear_solver_used = solver;
  if (!g->Hdiag_alt) {
    PGO_TRY(dev_alloc(g, &g->Hdiag_alt, (size_t)R * 36));
    PGO_TRY(dev_alloc(g, &g->Hoff_alt, (size_t)g->nnz_off * 36));
    PGO_TRY(dev_alloc(g, &g->grad_alt, (size_t)R * 6));
    CUDA_TRY(cudaMemsetAsync(g->Hoff_alt, 0, std::max<size_t>((size_t)g->nnz_off * 36, 1) * sizeof(double), g->stream));
  }
  // device-resident LM state, the iteration log, and a pinned ring the host reads the state from
  const int dev_log_cap = std::max(2, std::min(opt->max_num_iterations + 2, 1 << 16));
  if (!g->lm_state) {
    PGO_TRY(dev_alloc(g, &g->lm_state, 1));
    CUDA_TRY(pool_pinned(g->device, reinterpret_cast<void**>(&g->lm_ring)));
    static_assert(4 * sizeof(LmState) <= kPinnedBytes, "pinned LM state ring");
    for (int k = 0; k < 4; ++k) CUDA_TRY(pool_event(g->device, &g->lm_ev[k]));
  }
  if (g->lm_log_cap < dev_log_cap) {
    PGO_TRY(dev_alloc(g, &g->lm_log, (size_t)dev_log_cap));
    g->lm_log_cap = dev_log_cap;
  }
  LmState* st = g->lm_state;
  summary->hessian_blocks = g->nnz_off + R;
  if (solver == PGO_LINEAR_PCG_AMG) { summary->amg_levels = g->amg->num_levels; summary->amg_blocks = g->amg->blocks_all_levels; }
  if (solver == PGO_LINEAR_PCG_LEVEL_CHOLESKY) { summary->factor_blocks = g->chol->factor_blocks; summary->factor_levels = g->chol->num_levels; }
  summary->time_setup_s = g->setup_s;
  LmOptions lo;
  lo.max_num_iterations = opt->max_num_iterations;
  lo.function_tolerance = opt->function_tolerance; lo.gradient_tolerance = opt->gradient_tolerance; lo.parameter_tolerance = opt->parameter_tolerance;
  lo.initial_radius = opt->initial_trust_region_radius; lo.max_radius = opt->max_trust_region_radius; lo.min_radius = opt->min_trust_region_radius;
  lo.min_relative_decrease = opt->min_relative_decrease; lo.max_consecutive_invalid = opt->max_num_consecutive_invalid_steps;
  lo.verbose = opt->verbose;

  // ---- IterationZero: evaluate, Jacobi scaling from the unscaled diagonal, re-linearize scaled ----
  CUDA_TRY(cudaMemcpyAsync(g->scale, g->scale_eval, (size_t)N * 6 * sizeof(double), cudaMemcpyDeviceToDevice, g->stream));
  PGO_TRY(zero_scalars(g));
  CUDA_TRY(cudaEventRecord(g->ev0, g->stream));
  PGO_TRY(linearize_full(g, g->poses, g->scale, opt->loss_type, opt->loss_a));
  summary->num_linearizations++;
  if (opt->jacobi_scaling) {
    jacobi_scale_kernel<<<nblk, tpb, 0, g->stream>>>(R, g->Hdiag, g->scale_eval, 1, g->scale);
    g->launches++;
    PGO_TRY(halo_exchange0(g, g->scale, 6));   // the off-diagonal blocks of a cut edge need the other owner's column scaling
    PGO_TRY(zero_scalars(g));
    PGO_TRY(linearize_full(g, g->poses, g->scale, opt->loss_type, opt->loss_a));
    summary->num_linearizations++;
  }
  CUDA_TRY(cudaEventRecord(g->ev1, g->stream));
  xnorm_kernel<<<(R + 255) / 256, 256, 0, g->stream>>>(R, g->poses, g->active, g->scale_eval, g->scalars);
  gradient_norm_kernel<<<(R + 255) / 256, 256, 0, g->stream>>>(R, g->poses, g->grad, g->scale, g->active, nullptr, g->scalars);
  PGO_TRY(reduce_scalars(g));
  lm_init_kernel<<<1, 32, 0, g->stream>>>(st, g->scalars, lo, g->lm_log, g->lm_log_cap);
  g->launches += 3;

  // ---- the loop ----
  const LinTarget alt = {g->Hdiag_alt, g->Hoff_alt, g->grad_alt};
  const LmDiagonal lmd = {0, opt->min_lm_diagonal, opt->max_lm_diagonal, opt->initial_trust_region_radius, g->diagonal, g->dlm};
  const int commit_ctas = std::max(1, std::min((int)(((long long)R * 36 + g->nnz_off * 36 + 511) / 512), 4 * g->num_sms));
  auto enqueue_iteration = [&]() -> int {
    // LevenbergMarquardtStrategy::ComputeStep: D = diag / radius, solve (H + D) y = g, step = -y
    PGO_TRY(linear_solve_device(g, opt, solver, g->grad, lmd, st));
    // candidate = Plus(x, -y .* scale), |step|, |x_cand|; the alternate system is zeroed on the way
    plus_kernel<<<(R + 255) / 256, 256, 0, g->stream>>>(R, g->poses, g->vx, g->scale, g->active, -1.0, g->poses_cand, g->scalars, st,
                                                        g->Hdiag_alt, g->grad_alt);
    g->launches++;
    if (g->has_dup_blocks && g->nnz_off) CUDA_TRY(cudaMemsetAsync(g->Hoff_alt, 0, (size_t)g->nnz_off * 36 * sizeof(double), g->stream));
    PGO_TRY(halo_exchange0(g, g->poses_cand, 8, &st->done));   // candidate poses of the cut edges' other endpoints
    PGO_TRY(launch_linearize(g, kLinFull, g->poses_cand, g->scale, opt->loss_type, opt->loss_a, nullptr, nullptr, &alt, st));
    gradient_norm_kernel<<<(R + 255) / 256, 256, 0, g->stream>>>(R, g->poses_cand, g->grad_alt, g->scale, g->active, nullptr, g->scalars, st);
    PGO_TRY(reduce_scalars(g));
    lm_decide_kernel<<<1, 32, 0, g->stream>>>(st, g->scalars, lo, g->lm_log, g->lm_log_cap);
    lm_commit_kernel<<<commit_ctas, 256, 0, g->stream>>>(st, g->scalars, (long long)N * 8, g->poses_cand, g->poses, (long long)R * 36, g->Hdiag_alt,
                                                         g->Hdiag, g->nnz_off * 36, g->Hoff_alt, g->Hoff, (long long)R * 6, g->grad_alt, g->grad);
    g->launches += 3;
    return PGO_OK;
  };
  const int depth = solver_is_host_polled(g, solver) ? 0 : 2;   // iterations the host may run ahead of what it has seen
  // The iteration is the same dozen launches every time (all that varies lives in device memory): capture it once per
  // (options, solver, log buffer) and replay it -- the kernels between the linear solves are a few microseconds each.
  static const bool graph_off = getenv("PGO_LM_GRAPH") && atoi(getenv("PGO_LM_GRAPH")) == 0;
  bool use_graph = depth > 0 && g->world == 1 && !graph_off && getenv("PGO_TIMELINE") == nullptr;
  // (captured lazily before the SECOND iteration of a solve: the first one runs as plain launches so that every
  // one-time initialisation -- function attributes, occupancy queries -- happens outside a capture)
  auto ensure_graph = [&]() -> int {
    if (use_graph) {
      // everything the captured launches carry BY VALUE (field by field: struct padding and the host pointers are not part of it)
      const double fields[] = {(double)opt->max_num_iterations, opt->function_tolerance, opt->gradient_tolerance, opt->parameter_tolerance,
                               opt->initial_trust_region_radius, opt->max_trust_region_radius, opt->min_trust_region_radius,
                               opt->min_relative_decrease, opt->min_lm_diagonal, opt->max_lm_diagonal,
                               (double)opt->max_num_consecutive_invalid_steps, (double)opt->jacobi_scaling, (double)opt->loss_type, opt->loss_a,
                               (double)opt->linear_solver_type, (double)opt->pcg_max_iterations, opt->pcg_tolerance, (double)opt->pcg_num_ctas,
                               opt->direct_residual_accept, (double)solver, (double)(g->edge_loss_set ? 1 : 0),
                               (double)(uintptr_t)g->lm_log, (double)(uintptr_t)g->stream};
      std::vector<unsigned char> key(sizeof fields);
      std::memcpy(key.data(), fields, sizeof fields);
      if (!g->lm_graph || key != g->lm_graph_key) {
        if (g->lm_graph) { cudaGraphExecDestroy(g->lm_graph); g->lm_graph = nullptr; }
        const long long l0 = g->launches;
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(g->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_iteration();
        const cudaError_t ce = cudaStreamEndCapture(g->stream, &graph);
        g->lm_graph_kernels = (int)(g->launches - l0);
        g->launches = l0;
        if (rc != PGO_OK || ce != cudaSuccess) {
          if (graph) cudaGraphDestroy(graph);
          cudaGetLastError();
          use_graph = false;                    // fall back to plain launches (same kernels, same order)
        } else {
          const cudaError_t ie = cudaGraphInstantiate(&g->lm_graph, graph, 0);
          cudaGraphDestroy(graph);
          if (ie != cudaSuccess) { cudaGetLastError(); g->lm_graph = nullptr; use_graph = false; }
          else g->lm_graph_key = key;
        }
      }
    }
    return PGO_OK;
  };
  LmState final_state;
  {
    // state after iteration zero (a non-finite initial cost or a zero gradient ends the solve before the loop)
    int k = 0;
    bool finished = false;
    auto publish = [&](int slot) -> int {
      CUDA_TRY(cudaMemcpyAsync(g->lm_ring + slot, st, sizeof(LmState), cudaMemcpyDeviceToHost, g->stream));
      CUDA_TRY(cudaEventRecord(g->lm_ev[slot], g->stream));
      return PGO_OK;
    };
    PGO_TRY(publish(3));
    if (depth == 0) {
      CUDA_TRY(cudaEventSynchronize(g->lm_ev[3]));
      finished = g->lm_ring[3].done != 0;
    }
    const int hard_cap = opt->max_num_iterations + 8;
    while (!finished && k < hard_cap) {
      if (use_graph && k > 0) PGO_TRY(ensure_graph());
      if (use_graph && k > 0 && g->lm_graph) { CUDA_TRY(cudaGraphLaunch(g->lm_graph, g->stream)); g->launches += g->lm_graph_kernels; }
      else PGO_TRY(enqueue_iteration());
      summary->num_linearizations++;
      const int slot = k % 3;
      PGO_TRY(publish(slot));
      if (depth == 0) {
        CUDA_TRY(cudaEventSynchronize(g->lm_ev[slot]));
        finished = g->lm_ring[slot].done != 0;
      } else if (k == 0) {
        CUDA_TRY(cudaEventSynchronize(g->lm_ev[3]));          // the state after iteration zero
        finished = g->lm_ring[3].done != 0;
      } else if (k >= depth - 1) {
        const int seen = (k - (depth - 1)) % 3;               // the iteration `depth - 1` before this one
        CUDA_TRY(cudaEventSynchronize(g->lm_ev[seen]));
        finished = g->lm_ring[seen].done != 0;
      }
      ++k;
    }
    CUDA_TRY(cudaMemcpyAsync(g->lm_ring + 3, st, sizeof(LmState), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    final_state = g->lm_ring[3];
    // linearisations that actually ran (iterations enqueued past the end returned at once)
    summary->num_linearizations = (opt->jacobi_scaling ? 2 : 1) + std::max(final_state.num_rows - 1, 0);
  }
  CUDA_TRY(cudaGetLastError());
  {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, g->ev0, g->ev1));
    summary->time_linearize_ms = ms + final_state.linearize_ns * 1e-6;
    summary->time_linear_solver_ms = final_state.solver_ns * 1e-6;
  }
  const int rows = std::min(final_state.num_rows, g->lm_log_cap);
  std::vector<pgo_iteration_summary> rows_h((size_t)std::max(rows, 1));
  if (rows > 0) CUDA_TRY(cudaMemcpy(rows_h.data(), g->lm_log, (size_t)rows * sizeof(pgo_iteration_summary), cudaMemcpyDeviceToHost));
  if (log) for (int k = 0; k < rows && k < log_cap; ++k) log[k] = rows_h[k];
  if (opt->verbose)
    for (int k = 1; k < rows; ++k)
      fprintf(stderr, "[pgo] it %d radius %.3e pcg %d (rel %.2e) cost %.6e %s\n", rows_h[k].iteration, rows_h[k].trust_region_radius,
              rows_h[k].linear_solver_iterations, rows_h[k].pcg_relative_residual, rows_h[k].cost,
              rows_h[k].step_is_successful ? "accepted" : (rows_h[k].step_is_valid ? "rejected" : "invalid"));
  summary->num_iterations = final_state.num_rows;
  summary->initial_cost = final_state.initial_cost;
  summary->final_cost = final_state.x_cost;
  summary->num_successful_steps = final_state.num_successful;
  summary->num_unsuccessful_steps = final_state.num_unsuccessful;
  summary->termination_type = final_state.done ? final_state.termination_type : PGO_NO_CONVERGENCE;
  summary->total_pcg_iterations = final_state.total_pcg_iterations;
  lm_message(final_state, opt, summary->message, sizeof summary->message);
  summary->comm_calls = g->comm_calls - comm_calls0;
  summary->comm_bytes = g->comm_bytes - comm_bytes0;
  if (g->amg) { summary->comm_bytes_per_pcg_iteration = g->amg->comm_bytes_per_iteration; summary->comm_calls_per_pcg_iteration = g->amg->comm_calls_per_iteration; }
  summary->peer_exchanges = amg_peer_pushes(g) - peer_x0;
  summary->peer_bytes = amg_peer_bytes(g) - peer_b0;
  if (g->amg && g->amg->peer) {
    summary->peer_exchanges_per_pcg_iteration = g->amg->comm_calls_per_iteration;
    summary->comm_calls_per_pcg_iteration = 0;
  }
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

// ------------------------------------------------------------------------------------------------
// One-shot entry point with a per-topology cache.  What ceres::Solve does for the reference on every call -- program
// preprocessing, symbolic factorisation -- depends only on the TOPOLOGY (edge endpoints, constant flags).  A repeat call
// on the same topology (a SLAM back end re-solving its graph with refreshed measurements; bench.py's end-to-end loop)
// reuses the device-resident graph: pattern, tiles' index part, elimination schedule, factor storage, LM buffers.  Only
// poses and measurements are re-uploaded.  Keyed by a 64-bit hash and confirmed by a full comparison of the arrays.
// PGO_NO_TOPOLOGY_CACHE=1 disables it; pgo_release_cached_memory() empties it.
// ------------------------------------------------------------------------------------------------
static std::mutex g_graph_cache_mu;
static std::vector<pgo_graph*> g_graph_cache;          // most recently used last
static std::atomic<bool> g_topology_cache_enabled{getenv("PGO_NO_TOPOLOGY_CACHE") == nullptr};
extern "C" int pgo_set_topology_cache(int enabled) {
  const bool was = g_topology_cache_enabled.exchange(enabled != 0);
  if (!enabled) graph_cache_clear(-1);
  return was ? 1 : 0;
}
constexpr size_t kGraphCacheEntries = 4;
constexpr int kGraphCacheMaxPoses = 200000;            // larger graphs are not worth pinning GBs of HBM for

static unsigned long long topology_hash(int n_poses, int n_edges, const int* edge_ids, const unsigned char* pose_const) {
  unsigned long long h = 1469598103934665603ull ^ ((unsigned long long)n_poses << 32) ^ (unsigned long long)n_edges;
  const unsigned long long* w = reinterpret_cast<const unsigned long long*>(edge_ids);   // one (a, b) pair per word
  for (int e = 0; e < n_edges; ++e) { h ^= w[e]; h *= 1099511628211ull; h ^= h >> 29; }
  if (pose_const) for (int i = 0; i < n_poses; ++i) if (pose_const[i]) { h ^= (unsigned long long)i * 0x9E3779B97F4A7C15ull + pose_const[i]; h *= 1099511628211ull; }
  return h;
}

static void graph_cache_clear(int device) {
  std::vector<pgo_graph*> drop;
  {
    std::lock_guard<std::mutex> lk(g_graph_cache_mu);
    for (auto it = g_graph_cache.begin(); it != g_graph_cache.end();)
      if (device < 0 || (*it)->device == device) { drop.push_back(*it); it = g_graph_cache.erase(it); } else ++it;
  }
  for (pgo_graph* g : drop) pgo_graph_destroy(g);
}

// refresh the values of a cached graph: poses, measurements, square-root information (same identity / full class)
static int graph_update_values(pgo_graph* g, const double* poses, const double* edge_meas, const double* edge_sqrt_info) {
  CUDA_TRY(cudaSetDevice(g->device));
  const int E = g->E, T = g->T;
  for (int e = 0; e < E; ++e) {
    const int pos = g->edge_pos[e];
    EdgeCoreTile& t = g->core_host[pos / kTile];
    const int l = pos % kTile;
    for (int k = 0; k < 7; ++k) t.meas[k][l] = edge_meas[7 * (size_t)e + k];
  }
  CUDA_TRY(cudaMemcpyAsync(g->core, g->core_host.data(), (size_t)std::max(T, 1) * sizeof(EdgeCoreTile), cudaMemcpyHostToDevice, g->stream));
  if (!g->identity_info) {
    for (int e = 0; e < E; ++e)
      for (int k = 0; k < 36; ++k) g->info_host[g->edge_pos[e] / kTile].S[k][g->edge_pos[e] % kTile] = edge_sqrt_info[36 * (size_t)e + k];
    CUDA_TRY(cudaMemcpyAsync(g->info, g->info_host.data(), (size_t)std::max(T, 1) * sizeof(EdgeInfoTile), cudaMemcpyHostToDevice, g->stream));
  }
  return pgo_graph_set_poses(g, poses);
}

static bool info_is_identity(int n_edges, const double* edge_sqrt_info) {
  if (!edge_sqrt_info) return true;
  static const double eye[36] = {1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1};
  for (int e = 0; e < n_edges; ++e)
    for (int k = 0; k < 36; ++k) if (edge_sqrt_info[36 * (size_t)e + k] != eye[k]) return false;
  return true;
}

extern "C" int pgo_solve_pose_graph(int device, int n_poses, double* poses, int n_edges, const int* edge_ids,
                                    const double* edge_meas, const double* edge_sqrt_info,
                                    const unsigned char* pose_const, const pgo_solver_options* options,
                                    pgo_solver_summary* summary, pgo_iteration_summary* iteration_log,
                                    int iteration_log_capacity) {
  pgo_solver_options defaults;
  if (!options) { pgo_default_options(&defaults); options = &defaults; }
  const bool cacheable = g_topology_cache_enabled.load() && n_poses > 0 && n_poses <= kGraphCacheMaxPoses && n_edges > 0 && edge_ids && poses && edge_meas;
  pgo_graph* g = nullptr;
  unsigned long long h = 0;
  const double t0 = wall_s();
  if (cacheable) {
    h = topology_hash(n_poses, n_edges, edge_ids, pose_const);
    const bool ident = info_is_identity(n_edges, edge_sqrt_info);
    std::lock_guard<std::mutex> lk(g_graph_cache_mu);
    for (auto it = g_graph_cache.begin(); it != g_graph_cache.end(); ++it) {
      pgo_graph* c = *it;
      if (c->device != device || c->topo_hash != h || c->N_global != n_poses || c->E_global != n_edges || c->identity_info != ident) continue;
      if (std::memcmp(c->topo_edge_ids.data(), edge_ids, (size_t)n_edges * 2 * sizeof(int)) != 0) continue;
      bool same_const = true;
      for (int i = 0; i < n_poses && same_const; ++i) same_const = c->topo_const[i] == (pose_const ? pose_const[i] : 0);
      if (!same_const) continue;
      g = c;
      g_graph_cache.erase(it);            // checked out: nobody else can pick it while it is in use
      break;
    }
  }
  int rc = PGO_OK;
  if (g) {
    rc = graph_update_values(g, poses, edge_meas, edge_sqrt_info);
    g->setup_s = wall_s() - t0;
  } else {
    g_symbolic_hint = options->linear_solver_type;     // pgo_graph_create starts the symbolic analysis on a helper thread
    g_keep_host_tiles = cacheable;
    rc = pgo_graph_create(&g, device, n_poses, n_edges, poses, edge_ids, edge_meas, edge_sqrt_info, pose_const);
    g_symbolic_hint = -1;
    g_keep_host_tiles = false;
    if (rc == PGO_OK && cacheable) {
      g->topo_hash = h;
      g->topo_edge_ids.assign(edge_ids, edge_ids + 2 * (size_t)n_edges);
      g->topo_const.assign((size_t)n_poses, 0);
      if (pose_const) g->topo_const.assign(pose_const, pose_const + n_poses);
    }
  }
  if (rc == PGO_OK) rc = pgo_graph_set_edge_losses(g, options->edge_loss_type, options->edge_loss_a);   // NULLs clear a cached graph's
  if (rc != PGO_OK) { if (g) pgo_graph_destroy(g); return rc; }
  rc = pgo_graph_solve(g, options, summary, iteration_log, iteration_log_capacity);
  if (rc == PGO_OK) rc = pgo_graph_get_poses(g, poses);
  if (rc == PGO_OK && cacheable) {
    pgo_graph* evict = nullptr;
    {
      std::lock_guard<std::mutex> lk(g_graph_cache_mu);
      g_graph_cache.push_back(g);
      if (g_graph_cache.size() > kGraphCacheEntries) { evict = g_graph_cache.front(); g_graph_cache.erase(g_graph_cache.begin()); }
    }
    if (evict) pgo_graph_destroy(evict);
  } else {
    pgo_graph_destroy(g);
  }
  return rc;
}
