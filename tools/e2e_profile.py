#!/usr/bin/env python3
"""Wall-clock split of the one-shot entry point on KITTI-00 (PGO_PROFILE_HOST=1 prints the host-side laps)."""
import os, sys, time
os.environ["PGO_PROFILE_HOST"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import posegraph_ceres_b200 as P
g = P.datasets.kitti00()
o = P.default_options()
for k in range(4):
    t0 = time.perf_counter()
    poses, s, its = P.solve_pose_graph(g.poses, g.edge_ids, g.edge_meas, g.edge_sqrt_info, g.pose_const, o)
    dt = time.perf_counter() - t0
    print(f"e2e {1e3*dt:.3f} ms: summary total {1e3*s.time_total_s:.3f} ms setup {1e3*s.time_setup_s:.3f} ms, solver {s.time_linear_solver_ms:.3f} ms, linearize {s.time_linearize_ms:.3f} ms", file=sys.stderr, flush=True)
