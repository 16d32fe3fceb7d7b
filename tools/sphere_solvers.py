#!/usr/bin/env python3
"""sphere2500 (BASELINE configs[2]) with each linear solver: time per LM iteration and agreement of the results."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import posegraph_ceres_b200 as P
g = P.datasets.sphere()
res = {}
for name, solver in (("block-jacobi PCG", P.LINEAR_PCG_BLOCK_JACOBI), ("level-Cholesky PCG", P.LINEAR_PCG_LEVEL_CHOLESKY)):
    o = P.default_options()
    o.linear_solver_type = solver
    for rep in range(2):
        t0 = time.perf_counter()
        poses, s, its = P.solve_pose_graph(g.poses, g.edge_ids, g.edge_meas, g.edge_sqrt_info, g.pose_const, o)
        dt = time.perf_counter() - t0
    res[name] = poses
    print(f"{name}: e2e {1e3 * dt:.1f} ms, {s.num_iterations - 1} LM iterations, pcg {s.total_pcg_iterations}, solver {s.time_linear_solver_ms:.1f} ms, "
          f"setup {1e3 * s.time_setup_s:.1f} ms, final cost {s.final_cost:.6f}", flush=True)
print("max |dp| between solvers", np.abs(res["block-jacobi PCG"][:, :3] - res["level-Cholesky PCG"][:, :3]).max())
