/*
 * pgo_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of what the reference does on its hot path:
 *   BuildOptimizationProblem()  /root/reference/src/POSE_GRAPH_CERES_PLUS/test/pose_graph_ceres_plus_finial.cpp:461-497
 *   SolveOptimizationProblem()  same file :500-514  (ceres::Solve, SPARSE_NORMAL_CHOLESKY, 1000 iterations)
 *   PoseGraph3dErrorTerm        /root/reference/src/POSE_GRAPH_CERES_PLUS/include/PoseGraph3dError.h:21-54
 * and of the parts of Ceres Solver (third-party, NOT vendored in /root/reference; the
 * reference was written against ceres-solver 1.12/1.13, Sept. 2017) those calls run:
 * AutoDiffCostFunction (forward-mode jets), HuberLoss + Corrector,
 * EigenQuaternionParameterization, TrustRegionMinimizer + LevenbergMarquardtStrategy,
 * SPARSE_NORMAL_CHOLESKY (here: block min-degree ordering + block up-looking Cholesky).
 *
 * PARITY PARTIALLY PINNED: Ceres is not installable in this image, so the LM
 * iterate sequence cannot be compared with a real Ceres run step by step.  The COST FUNCTION is pinned against the
 * reference's own source: oracle/_ref/libref_functor.so compiles PoseGraph3dErrorTerm::operator() unmodified from
 * /root/reference (ref_functor.cpp, minimal Eigen/Ceres stand-ins in ref_shim/) and oracle_evaluate's residuals and
 * Jacobians agree with it to rounding on random edges.  What is pinned against the
 * reference's own Ceres output (the result/trajectory text files): the cost function's stationarity at the reference's
 * optimised trajectory on every pose (free poses; loop-edge END poses with measurements recovered from the
 * BEGIN poses only), and the end result -- from the reference's initial trajectory the oracle's LM lands
 * within 2.3 cm (max) / 0.8 cm (mean) of the reference's optimised trajectory (tests/test_oracle_cpu.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 *
 * Data layout (all host pointers, double precision):
 *   poses      [n_poses][7]   x y z qx qy qz qw          (Pose3d, include/types.h:16-21;
 *                                                         q in Eigen coeffs() order x,y,z,w)
 *   edge_ids   [n_edges][2]   id_begin (a), id_end (b)   (Edge3d, include/types.h:30-45)
 *   edge_meas  [n_edges][7]   t_be: x y z qx qy qz qw
 *   edge_sqrt_info [n_edges][36]  6x6 ROW-major sqrt_information (residual = S * r)
 *   pose_const [n_poses]      1 = SetParameterBlockConstant on both p and q
 */
#ifndef PGO_ORACLE_H_
#define PGO_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_LOSS_TRIVIAL = 0, ORACLE_LOSS_HUBER = 1, ORACLE_LOSS_CAUCHY = 2 };

enum {
  ORACLE_CONVERGENCE = 0,
  ORACLE_NO_CONVERGENCE = 1,
  ORACLE_FAILURE = 2
};

typedef struct {
  int max_num_iterations;             /* ceres default 50; the reference sets 1000 */
  double function_tolerance;          /* 1e-6  */
  double gradient_tolerance;          /* 1e-10 */
  double parameter_tolerance;         /* 1e-8  */
  double initial_trust_region_radius; /* 1e4   */
  double max_trust_region_radius;     /* 1e16  */
  double min_trust_region_radius;     /* 1e-32 */
  double min_relative_decrease;       /* 1e-3  */
  double min_lm_diagonal;             /* 1e-6  */
  double max_lm_diagonal;             /* 1e32  */
  int max_num_consecutive_invalid_steps; /* 5 */
  int jacobi_scaling;                 /* 1 */
  int loss_type;                      /* ORACLE_LOSS_*; the reference uses HUBER */
  double loss_a;                      /* HuberLoss(1.0) */
  int ordering;                       /* 0 = natural, 1 = minimum degree */
} oracle_options;

/* one row per minimizer iteration, mirrors ceres::IterationSummary */
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
} oracle_iteration;

typedef struct {
  double initial_cost;
  double final_cost;
  int num_successful_steps;
  int num_unsuccessful_steps;
  int num_iterations;      /* rows written to the iteration log */
  int termination_type;    /* ORACLE_CONVERGENCE ... */
  char message[160];
  double time_total_s;
  double time_residual_s;  /* cost-only evaluations */
  double time_jacobian_s;  /* residual + jacobian evaluations */
  double time_linear_solver_s;
  long long factor_nnz_blocks;   /* 6x6 blocks in the Cholesky factor */
  int num_jacobian_evals;
  int num_residual_evals;
} oracle_summary;

void oracle_default_options(oracle_options* o);

/* threads of the per-edge evaluation loop (default 1 = Ceres' default num_threads) */
void oracle_set_num_threads(int n);
/* per-residual-block loss functions (NULL, NULL restores the options' single loss); arrays must outlive the calls */
void oracle_set_edge_losses(int n_edges, const int* type, const double* a);
int oracle_get_num_threads(void);

/* Problem::Evaluate: cost, robustified residuals [6E], gradient [6N] (local/tangent coordinates,
 * zero for constant poses) and the per-edge local Jacobian blocks jac[E][2][36] (row-major 6x6,
 * block 0 w.r.t. pose a, block 1 w.r.t. pose b; zero for constant poses). Any output may be NULL. */
int oracle_evaluate(int n_poses, const double* poses, const unsigned char* pose_const,
                    int n_edges, const int* edge_ids, const double* edge_meas,
                    const double* edge_sqrt_info, int loss_type, double loss_a,
                    double* cost, double* residuals, double* gradient, double* jac);

/* LocalParameterization::Plus over all poses: p += delta[0:3], q = dq(delta[3:6]) * q */
void oracle_plus(int n_poses, const double* poses, const double* delta, double* out);

/* ceres::Solve. poses is in/out. iter_log may be NULL; at most iter_log_cap rows are written. */
int oracle_solve(int n_poses, double* poses, const unsigned char* pose_const,
                 int n_edges, const int* edge_ids, const double* edge_meas,
                 const double* edge_sqrt_info, const oracle_options* opt,
                 oracle_summary* summary, oracle_iteration* iter_log, int iter_log_cap);

/* Solve (J^T J + diag(d)) y = rhs with the oracle's sparse Cholesky on the problem's structure;
 * jac as returned by oracle_evaluate; d, rhs, y are [6N]. For tests of the GPU PCG solver. */
int oracle_normal_solve(int n_poses, const unsigned char* pose_const, int n_edges,
                        const int* edge_ids, const double* jac, const double* d,
                        const double* rhs, double* y, int ordering);

/* Loop-edge candidate lists, the caller side of the path -- PINNED bit-exactly against the reference's own
 * config/Edge_Candidates_index.txt (tests/golden/kitti00_fixture.npz cand_*; tests/test_oracle_cpu.py).
 * getCandidatesIndex()  /root/reference/src/POSE_GRAPH_CERES_PLUS/test/generate_edges_from_trajectory_origion.cpp:58-82
 * isInSearchRange()     same file :84-110 (float arithmetic on CV_32F poses, radius from config "search_radius").
 * positions [n][3]; row_ptr [n+1]; candidates (capacity entries) may be NULL to count only.  Returns the total. */
long long oracle_edge_candidates(int n_frames, const double* positions, double search_radius, int min_frame_gap,
                                 long long* row_ptr, int* candidates, long long capacity);

#ifdef __cplusplus
}
#endif
#endif
