#!/usr/bin/env python3
"""Launch shapes of the level-Cholesky solver on the small configs: solve time, iterations, parity vs the CPU oracle.
One subprocess per shape (the shape override is read once per process).  python tools/small_graph_shapes.py [timeline]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "oracle")); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import posegraph_ceres_b200 as P
import oracle_py as oracle
from helpers import rot_angle_between
name = sys.argv[1]
g = {"manhattan": P.datasets.manhattan_loop, "kitti": P.datasets.kitti00, "sphere_small": lambda: P.datasets.sphere(10, 20, None)}[name]()
o = P.default_options()
if os.environ.get("PGO_TIMELINE"): o.max_num_iterations = 2
G = P.Graph.from_dataset(g); s, _ = G.solve(o); G.close()
best = 1e9
for k in range(5):
    G = P.Graph.from_dataset(g)
    t = time.perf_counter(); s, _ = G.solve(o); dt = time.perf_counter() - t
    poses = G.get_poses(); G.close(); best = min(best, dt)
ref, rs, _ = oracle.solve(g)
print(f"{name:13s} shape={os.environ.get('PGO_CHOL_SHAPE', 'auto'):8s} solve {best * 1e3:7.3f} ms  LM iterations {s.num_iterations:3d}  linear solver {s.time_linear_solver_ms:7.3f} ms  "
      f"pcg {s.total_pcg_iterations}  max|dp| {np.abs(poses[:, :3] - ref[:, :3]).max():.2e}  max angle {rot_angle_between(poses[:, 3:], ref[:, 3:]).max():.2e}  oracle its {rs.num_iterations}")
''' % (ROOT, ROOT, ROOT)

for name in ("manhattan", "sphere_small", "kitti"):
    for shape in (("auto",) if os.environ.get("SHAPES_AUTO_ONLY") else ("auto", "block", "cluster", "grid")):
        if name == "kitti" and shape == "block" and len(sys.argv) > 1:
            continue
        env = dict(os.environ)
        env.pop("PGO_CHOL_SHAPE", None)
        if shape != "auto":
            env["PGO_CHOL_SHAPE"] = shape
        if len(sys.argv) > 1 and sys.argv[1] == "timeline":
            env["PGO_TIMELINE"] = "1"
        r = subprocess.run([sys.executable, "-c", CHILD, name], env=env, capture_output=True, text=True, timeout=300)
        sys.stdout.write(r.stdout)
        if r.returncode != 0 or "timeline" in sys.argv[1:]:
            sys.stdout.write(r.stderr[-3000:])
        sys.stdout.flush()
