/*
 * pgo_b200.h -- C-ABI of the B200-native pose-graph Levenberg-Marquardt solver.
 *
 * Drop-in boundary for the ONE hot path of TurtleZhong/PoseGraph-Ceres
 * (REF = /root/reference/src/POSE_GRAPH_CERES_PLUS):
 *
 *   BuildOptimizationProblem()   REF/test/pose_graph_ceres_plus_finial.cpp:461-497
 *       ceres::Problem::AddResidualBlock(PoseGraph3dErrorTerm, HuberLoss(1.0), p_a, q_a, p_b, q_b)
 *       ceres::Problem::SetParameterization(q, EigenQuaternionParameterization)
 *       ceres::Problem::SetParameterBlockConstant(first pose)
 *   SolveOptimizationProblem()   REF/test/pose_graph_ceres_plus_finial.cpp:500-514
 *       ceres::Solve(options{max_num_iterations = 1000, SPARSE_NORMAL_CHOLESKY}, problem, &summary)
 *   PoseGraph3dErrorTerm         REF/include/PoseGraph3dError.h:21-54   (the residual)
 *
 * Plain pointers and sizes only; all arrays are HOST memory unless a name ends in _dev.
 * Every function returns 0 on success, a negative pgo_status otherwise; pgo_last_error()
 * returns a message.  There is NO CPU fallback: without a CUDA device the calls fail.
 *
 * Array conventions (double precision):
 *   poses          [n_poses][7]  x y z qx qy qz qw   (Pose3d, REF/include/types.h:16-21; Eigen coeffs order)
 *   edge_ids       [n_edges][2]  id_begin, id_end    (Edge3d, REF/include/types.h:30-45)
 *   edge_meas      [n_edges][7]  t_be = T_begin^-1 * T_end, same layout as a pose
 *   edge_sqrt_info [n_edges][36] ROW-major 6x6 sqrt_information (residual = S * r); NULL = identity
 *   pose_const     [n_poses]     0 = variable, 1 = SetParameterBlockConstant(p) and (q) (REF :526-527 makes both calls for
 *                                the first pose), 2 = only p constant, 3 = only q constant
 *   tangent vectors (gradient, steps) are [n_poses][6]: (dx dy dz, d_rot[3]) in the local
 *   coordinates of EigenQuaternionParameterization: p+ = p + dp, q+ = Quat(cos|d|, sin|d|/|d| d) * q.
 */
#ifndef PGO_B200_H_
#define PGO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGO_B200_ABI_VERSION 6

typedef enum {
  PGO_OK = 0,
  PGO_ERR_INVALID_ARGUMENT = -1,
  PGO_ERR_CUDA = -2,
  PGO_ERR_NO_DEVICE = -3,
  PGO_ERR_NCCL = -4,
  PGO_ERR_NUMERICAL = -5
} pgo_status;

/* ceres::LossFunction the reference can pass to AddResidualBlock (NULL / HuberLoss / CauchyLoss). */
typedef enum { PGO_LOSS_TRIVIAL = 0, PGO_LOSS_HUBER = 1, PGO_LOSS_CAUCHY = 2 } pgo_loss_type;

/* ceres::TerminationType */
typedef enum { PGO_CONVERGENCE = 0, PGO_NO_CONVERGENCE = 1, PGO_FAILURE = 2 } pgo_termination_type;

/* Linear solver for the damped normal equations (replaces options.linear_solver_type =
 * SPARSE_NORMAL_CHOLESKY, REF/test/pose_graph_ceres_plus_finial.cpp:505). */
typedef enum {
  PGO_LINEAR_PCG_BLOCK_JACOBI = 0,   /* 6x6 block-Jacobi preconditioned CG (block-SpMV kernel)        */
  PGO_LINEAR_PCG_LEVEL_CHOLESKY = 1, /* PCG preconditioned by a level-scheduled block Cholesky factor */
  PGO_LINEAR_AUTO = 2,               /* LEVEL_CHOLESKY when the symbolic fill is small (chain-like graphs), else AMG
                                        (mesh-like graphs; tiny ones stay with BLOCK_JACOBI) */
  PGO_LINEAR_PCG_AMG = 3             /* PCG preconditioned by an aggregation-multigrid V-cycle on rigid-body modes of pose
                                        patches: every level is again a 6x6 block system swept by the block-SpMV kernel.
                                        The solver of every multi-GPU (row-partitioned) solve. */
} pgo_linear_solver_type;

/* Mirrors the ceres::Solver::Options fields that matter on this path; defaults = Ceres defaults
 * except max_num_iterations, which the reference sets to 1000. */
typedef struct {
  int max_num_iterations;
  double function_tolerance;
  double gradient_tolerance;
  double parameter_tolerance;
  double initial_trust_region_radius;
  double max_trust_region_radius;
  double min_trust_region_radius;
  double min_relative_decrease;
  double min_lm_diagonal;
  double max_lm_diagonal;
  int max_num_consecutive_invalid_steps;
  int jacobi_scaling;
  int loss_type;                 /* pgo_loss_type */
  double loss_a;
  int linear_solver_type;        /* pgo_linear_solver_type */
  int pcg_max_iterations;        /* per LM step */
  double pcg_tolerance;          /* stop when sqrt(r^T M^-1 r) <= tol * sqrt(b^T M^-1 b) */
  int pcg_num_ctas;              /* 0 = auto (persistent grid size of the PCG kernel) */
  double direct_residual_accept; /* LEVEL_CHOLESKY: the direct solve (first PCG iterate) is accepted without refinement when
                                    ||b - A x||_2 <= max(pcg_tolerance, this) * ||b||_2.  Default 1e-8: Ceres' own
                                    SPARSE_NORMAL_CHOLESKY never refines; PCG refinement stays the safety net. */
  int verbose;
  /* ceres::Problem::AddResidualBlock takes the loss function per residual block (REF :513-517).  NULL (default): every
   * edge uses loss_type / loss_a above.  Otherwise host arrays [n_edges] in the caller's edge order: pgo_loss_type and
   * scale a of each edge.  Read by pgo_solve_pose_graph; device-resident graphs use pgo_graph_set_edge_losses. */
  const int* edge_loss_type;
  const double* edge_loss_a;
} pgo_solver_options;

/* One row per minimizer iteration (ceres::IterationSummary). */
typedef struct {
  int iteration;
  int step_is_valid;
  int step_is_successful;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double gradient_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  int linear_solver_iterations;
  double pcg_relative_residual;
} pgo_iteration_summary;

/* ceres::Solver::Summary subset + device timings. */
typedef struct {
  double initial_cost;
  double final_cost;
  int num_successful_steps;
  int num_unsuccessful_steps;
  int num_iterations;            /* rows produced for the iteration log (incl. iteration 0) */
  int termination_type;          /* pgo_termination_type */
  char message[160];
  int num_linearizations;        /* residual + Jacobian + Hessian kernel launches (one per LM iteration: the candidate is
                                    linearised speculatively, so its cost, H and g arrive with the step statistics) */
  int num_cost_evaluations;      /* residual-only kernel launches (0 since the speculative linearisation) */
  long long total_pcg_iterations;
  long long kernel_launches;     /* kernels of this library launched by the call */
  double time_total_s;           /* host wall clock of the call */
  double time_setup_s;           /* structure analysis + uploads */
  double time_linearize_ms;      /* CUDA-event time in the linearize kernel(s) */
  double time_linear_solver_ms;  /* CUDA-event time in the linear solver */
  int linear_solver_used;        /* pgo_linear_solver_type actually used */
  long long hessian_blocks;      /* 6x6 blocks in the block-CSR Hessian (diag + off-diag) */
  long long factor_blocks;       /* 6x6 blocks in the level-Cholesky factor (0 if unused) */
  int factor_levels;
  int amg_levels;                /* levels of the multilevel preconditioner (0 if unused) */
  long long amg_blocks;          /* 6x6 blocks over all its levels stored on this rank */
  long long comm_calls;          /* multi-GPU: NCCL calls issued by this rank during the solve */
  long long comm_bytes;          /*            payload bytes this rank sent */
  long long comm_bytes_per_pcg_iteration;   /* halo + gather + scalar payload this rank sends per PCG iteration */
  int comm_calls_per_pcg_iteration;         /* NCCL calls per PCG iteration (0 when the exchanges run over peer memory) */
  long long peer_exchanges;      /* multi-GPU: exchanges done over NVLink peer memory (pgo_peer.cuh) during the solve ... */
  long long peer_bytes;          /*            ... and the bytes this rank stored into its peers' windows */
  int peer_exchanges_per_pcg_iteration;
} pgo_solver_summary;

/* Result of the host-side structure analysis (no GPU needed). */
typedef struct {
  int variable_poses;            /* poses that are used by an edge and not constant */
  long long hessian_blocks;      /* 6x6 blocks of the block-CSR Hessian (diag + off-diag) */
  int factor_usable;             /* 1 when the level-scheduled Cholesky fits max_fill_ratio */
  long long factor_blocks;       /* 6x6 blocks of L (diag + off-diag) */
  int factor_levels;             /* elimination levels = grid/cluster barriers per sweep */
  int factor_max_degree;
  long long factor_tasks;        /* 6x6 Schur update products per factorisation */
  double analysis_seconds;
} pgo_structure_info;

typedef struct pgo_graph pgo_graph;   /* opaque: a pose graph resident in HBM */

const char* pgo_last_error(void);
int pgo_abi_version(void);
int pgo_device_count(void);
void pgo_default_options(pgo_solver_options* options);

/* Host-only: variable poses, block-CSR Hessian pattern and the level-scheduled elimination order that
 * pgo_graph_create / pgo_graph_solve would use for this topology (what Ceres' program preprocessing and
 * symbolic factorisation do inside ceres::Solve).  max_fill_ratio <= 0: the full analysis, no limits (what an explicit
 * PGO_LINEAR_PCG_LEVEL_CHOLESKY request runs); > 0: the "cheap factor only" analysis of PGO_LINEAR_AUTO with this fill
 * limit (AUTO uses 8) plus its limits of 64 levels and node degree 16 -- it gives up early on mesh-like graphs. */
int pgo_analyze_structure(int n_poses, int n_edges, const int* edge_ids, const unsigned char* pose_const,
                          double max_fill_ratio, pgo_structure_info* info);

/* Device memory, streams, events and pinned buffers of destroyed graphs are cached per device and reused
 * by the next pgo_graph_create / pgo_solve_pose_graph; this returns them to the driver (device < 0: all). */
void pgo_release_cached_memory(int device);

/* Upload a pose graph (the content of ceres::Problem after BuildOptimizationProblem) to
 * device `device` and analyse its block structure (block-CSR Hessian pattern, edge->block map). */
int pgo_graph_create(pgo_graph** out, int device, int n_poses, int n_edges, const double* poses,
                     const int* edge_ids, const double* edge_meas, const double* edge_sqrt_info,
                     const unsigned char* pose_const);
void pgo_graph_destroy(pgo_graph* g);

/* Run all subsequent work of this graph on the caller's CUDA stream (a cudaStream_t passed as
 * void*, e.g. torch.cuda.current_stream().cuda_stream); NULL restores the graph's own stream. */
int pgo_graph_set_stream(pgo_graph* g, void* cuda_stream);

int pgo_graph_num_poses(const pgo_graph* g);
int pgo_graph_num_edges(const pgo_graph* g);
int pgo_graph_set_poses(pgo_graph* g, const double* poses);   /* host -> device */
int pgo_graph_get_poses(pgo_graph* g, double* poses);         /* device -> host */
/* keep / restore a device-side copy of the initial poses (bench: inputs already in HBM) */
int pgo_graph_snapshot_poses(pgo_graph* g);
int pgo_graph_restore_poses(pgo_graph* g);

/* Multi-GPU (one process per GPU): owner-computes row partition.  EVERY rank passes the same global graph; rank r keeps
 * the block rows of the poses [n r / W, n (r+1) / W) -- contiguous index ranges, pose graphs are trajectory-ordered --,
 * every edge that touches one of them (a cut edge is evaluated by both owners, each keeps its own rows of J^T J and J^T r)
 * and halo copies of the other endpoints.  Per PCG iteration only halo slices of the vectors move between neighbours plus
 * one all-reduce of two scalars; per LM iteration the halo poses and five scalars.  unique_id is the 128-byte ncclUniqueId
 * produced by rank 0 (pgo_nccl_unique_id) and distributed by the caller.  The calls below that take or return per-pose
 * arrays keep the GLOBAL layout on every rank and are collective: all ranks must make them in the same order. */
int pgo_nccl_unique_id(unsigned char unique_id[128]);
int pgo_graph_create_partitioned(pgo_graph** out, int device, int n_poses, int n_edges, const double* poses,
                                 const int* edge_ids, const double* edge_meas, const double* edge_sqrt_info,
                                 const unsigned char* pose_const, const unsigned char unique_id[128], int rank,
                                 int world_size);
int pgo_graph_rank(const pgo_graph* g);
int pgo_graph_world_size(const pgo_graph* g);
int pgo_graph_num_local_poses(const pgo_graph* g);   /* block rows owned by this rank */
int pgo_graph_num_halo_poses(const pgo_graph* g);    /* copies of other ranks' poses held for the cut edges */
int pgo_graph_num_local_edges(const pgo_graph* g);   /* edges evaluated here (cut edges count on both sides) */

/* Host-only: what pgo_graph_create_partitioned would set up for `rank` of `world_size` -- the partition, the halo
 * exchange plan and the multilevel hierarchy (no GPU needed; CPU tests cover the N > 1 host logic with it). */
typedef struct {
  int n_own, n_halo, n_local_edges, n_cut_edges;
  int n_neighbours;
  int send_total, recv_total;            /* level-0 halo plan */
  int send_to[64], recv_from[64];        /* per peer rank (world_size <= 64) */
  int amg_levels;
  int level_nodes[16];                   /* global nodes per level */
  int level_own[16], level_halo[16];     /* this rank's rows / halo columns per level */
  int level_replicated[16];
  long long level_blocks[16];            /* 6x6 blocks stored by this rank per level */
  int level_send[16], level_recv[16];
  unsigned long long plan_checksum;      /* order-dependent hash of send lists: the receiver-side hash must match */
  unsigned long long recv_checksum;
  int consistent;                        /* internal invariants hold (aggregates do not cross ranks, gather lists complete, ...) */
} pgo_partition_info;
int pgo_analyze_partition(int n_poses, int n_edges, const double* poses, const int* edge_ids,
                          const unsigned char* pose_const, int rank, int world_size, pgo_partition_info* info);
/* Host-only: the aggregation hierarchy itself.  level_nodes[16]; agg_out = concatenation over the levels 0..L-2 of the
 * aggregate (node id on the next level, -1 = not a variable) of every node; agg_out may be NULL to query the sizes. */
int pgo_amg_aggregates(int n_poses, int n_edges, const double* poses, const int* edge_ids, const unsigned char* pose_const,
                       int world_size, int* n_levels, int* level_nodes, int* agg_out, long long agg_capacity);

/* Per-edge loss functions of a device-resident graph ([n_edges] host arrays in the caller's edge order; NULL, NULL: back
 * to one loss for all edges).  While set they override the loss arguments of evaluate / linearize / solve. */
int pgo_graph_set_edge_losses(pgo_graph* g, const int* loss_type, const double* loss_a);

/* ceres::Problem::Evaluate on the device: cost, robustified residuals [n_edges][6], gradient
 * [n_poses][6] (unscaled, zero for constant poses), per-edge local Jacobians [n_edges][2][36]
 * (row-major 6x6; block 0 = d r / d pose_begin, block 1 = d r / d pose_end). Outputs may be NULL. */
int pgo_graph_evaluate(pgo_graph* g, int loss_type, double loss_a, double* cost, double* residuals,
                       double* gradient, double* jacobians);

/* One launch of the fused residual + Jacobian + J^T J / J^T r kernel at the current poses with
 * column scaling `scale` ([n_poses][6], NULL = ones). Leaves H and g in device memory.
 * Returns the cost; elapsed_ms (may be NULL) is the CUDA-event time of the kernel alone. */
int pgo_graph_linearize(pgo_graph* g, int loss_type, double loss_a, const double* scale,
                        double* cost, float* elapsed_ms);

/* Copy the assembled block-CSR Hessian to the host (tests): row_ptr [n_poses+1], col_idx [nnzb],
 * values [nnzb][36] ROW-major 6x6, diagonal block included (first in each row). Pass NULLs to
 * query nnzb. gradient [n_poses][6] = J^T r (scaled as H). */
int pgo_graph_get_hessian(pgo_graph* g, long long* nnzb, int* row_ptr, int* col_idx, double* values,
                          double* gradient);

/* y = (H + diag(d)) x with the block-SpMV kernel (tests / bench). x, y, d: [n_poses][6]; d may be NULL. */
int pgo_graph_spmv(pgo_graph* g, const double* x, const double* d, double* y, int repeats,
                   float* elapsed_ms);

/* Solve (H + diag(d)) y = b for the currently assembled H with the selected linear solver. */
int pgo_graph_linear_solve(pgo_graph* g, const pgo_solver_options* options, const double* d,
                           const double* b, double* y, int* iterations, double* relative_residual,
                           float* elapsed_ms);

/* ceres::Solve: Levenberg-Marquardt on the device-resident graph; poses are updated in HBM
 * (read them back with pgo_graph_get_poses). iteration_log may be NULL. */
int pgo_graph_solve(pgo_graph* g, const pgo_solver_options* options, pgo_solver_summary* summary,
                    pgo_iteration_summary* iteration_log, int iteration_log_capacity);

/* Convenience = what ceres::Solve(options, &problem, &summary) does for the reference: upload,
 * solve, write the optimised poses back into `poses` (in/out, host). */
int pgo_solve_pose_graph(int device, int n_poses, double* poses, int n_edges, const int* edge_ids,
                         const double* edge_meas, const double* edge_sqrt_info,
                         const unsigned char* pose_const, const pgo_solver_options* options,
                         pgo_solver_summary* summary, pgo_iteration_summary* iteration_log,
                         int iteration_log_capacity);

/* pgo_solve_pose_graph keeps up to four device-resident graphs per TOPOLOGY (edge endpoints + constant flags, hashed and
 * then compared in full): a repeat call on the same topology skips the structure analysis, the symbolic factorisation and
 * the index uploads and only refreshes poses and measurements.  enabled = 0 turns that off (and drops the kept graphs);
 * returns the previous setting.  Environment: PGO_NO_TOPOLOGY_CACHE=1.  pgo_release_cached_memory() also empties it. */
int pgo_set_topology_cache(int enabled);

/* Loop-edge candidate search, the caller side of the path: what getCandidatesIndex() / isInSearchRange() of
 * REF/test/generate_edges_from_trajectory_origion.cpp:58-110 compute into config/Edge_Candidates_index.txt (read back by
 * getEdegsCandidateIndex(), REF/include/ReadEdges.h:9-48).  positions [n_frames][3] camera centres (rounded to float like
 * the reference's CV_32F poses); for frame c = 1..n-1 the candidates are c-1 followed, in ascending order, by every
 * i < c - min_frame_gap with !(|p_i - p_c|^2 > radius^2) in float arithmetic -- bit-exact with the reference's file.
 * row_ptr [n_frames+1] (frame 0 has no entry); candidates may be NULL to query *total first. */
int pgo_edge_candidates(int device, int n_frames, const double* positions, double search_radius, int min_frame_gap,
                        long long* row_ptr, int* candidates, long long capacity, long long* total);

#ifdef __cplusplus
}
#endif
#endif /* PGO_B200_H_ */
