#!/usr/bin/env python3
"""Debug aid: phase timeline of the level-Cholesky kernel (PGO_TIMELINE=1) and wall-clock split of the
one-shot entry point on KITTI-00."""
import os
import sys
import time

if os.environ.get("NO_TIMELINE") is None:
    os.environ["PGO_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import posegraph_ceres_b200 as P  # noqa: E402

g = P.datasets.kitti00() if len(sys.argv) < 2 or sys.argv[1] == "kitti" else P.datasets.sphere()
o = P.default_options()
o.max_num_iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 3
o.verbose = 2
G = P.Graph.from_dataset(g)
s, its = G.solve(o)
print("iterations", s.num_iterations, "pcg", s.total_pcg_iterations, "solver ms", s.time_linear_solver_ms, flush=True)
G.close()
for k in range(4):
    t0 = time.perf_counter()
    G = P.Graph.from_dataset(g)
    t1 = time.perf_counter()
    s, its = G.solve(o)
    t2 = time.perf_counter()
    p = G.get_poses()
    t3 = time.perf_counter()
    G.close()
    t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.3f} ms  solve({s.num_iterations - 1} it) {1e3*(t2-t1):.3f} ms  get {1e3*(t3-t2):.3f} ms  destroy {1e3*(t4-t3):.3f} ms",
          file=sys.stderr, flush=True)
