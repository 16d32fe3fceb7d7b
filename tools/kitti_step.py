#!/usr/bin/env python3
"""A few device-resident KITTI-00 solves (for ncu launch lists / timelines): python tools/kitti_step.py [n_solves]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import posegraph_ceres_b200 as P  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = P.datasets.kitti00()
G = P.Graph.from_dataset(g)
G.snapshot_poses()
o = P.default_options()
for k in range(n):
    G.restore_poses()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s, its = G.solve(o)
    dt = time.perf_counter() - t0
    print(f"solve {k}: {1e3 * dt:.3f} ms wall, {s.num_iterations - 1} LM iterations, solver {s.time_linear_solver_ms:.3f} ms, linearize {s.time_linearize_ms:.3f} ms, "
          f"launches {s.kernel_launches}", flush=True)
G.close()
