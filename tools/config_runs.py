#!/usr/bin/env python3
"""BASELINE.json configs 0-2 (and, with --large, 3-4 -- solved to convergence, default options) on one GPU: wall-clock of
the one-shot entry point, LM iterations, linear solver used, and parity of the converged poses against the CPU oracle
where it finishes in seconds."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import oracle_py as O  # noqa: E402
import posegraph_ceres_b200 as P  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rot_angle_between  # noqa: E402

O.set_num_threads(len(os.sched_getaffinity(0)))
cases = [("configs[0] manhattan 100/120", P.datasets.manhattan_loop(), True),
         ("configs[1] KITTI-00 4541/5179", P.datasets.kitti00(), True),
         ("configs[2] sphere 2500/9799", P.datasets.sphere(), True)]
if "--large" in sys.argv:
    cases += [("configs[4] torus 100k / 10% random loops", P.datasets.torus(100000), False),
              ("configs[3] grid 1M / 2.05M edges", P.datasets.manhattan_grid(1000, 1000, 50000), False)]
for name, g, full in cases:
    o = P.default_options()
    for rep in range(2):      # second call: pools warm
        t0 = time.perf_counter()
        poses, s, its = P.solve_pose_graph(g.poses, g.edge_ids, g.edge_meas, g.edge_sqrt_info, g.pose_const, o)
        dt = time.perf_counter() - t0
    line = (f"{name}: GPU {1e3 * dt:.2f} ms e2e, {s.num_iterations - 1} LM iterations, solver {['block-jacobi', 'level-chol', 'auto', 'multilevel'][s.linear_solver_used]} ({s.message.decode()[:28]}) "
            f"(levels {s.factor_levels}, factor blocks {s.factor_blocks}), pcg iterations {s.total_pcg_iterations}, cost {s.initial_cost:.4f} -> {s.final_cost:.6f}, "
            f"linearize {s.time_linearize_ms:.2f} ms, linear solver {s.time_linear_solver_ms:.2f} ms")
    if full:
        t0 = time.perf_counter()
        ref, rs, rits = O.solve(g)
        cdt = time.perf_counter() - t0
        dp = np.abs(poses[:, :3] - ref[:, :3]).max()
        dr = rot_angle_between(poses[:, 3:], ref[:, 3:]).max()
        line += (f" | CPU oracle {1e3 * cdt:.1f} ms, {rs.num_iterations - 1} iterations, final {rs.final_cost:.6f} | max |dp| {dp:.2e} m, max angle {dr:.2e} rad"
                 f" | speed-up {cdt / dt:.1f}x")
    print(line, flush=True)
