#!/usr/bin/env python3
"""Launch the two hot kernels (linearize, block-SpMV) a few times on a large Manhattan grid graph;
the target of the `ncu --set full` captures committed under profiles/.
   python tools/run_large_kernels.py [side=1000] [identity|info]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import posegraph_ceres_b200 as P  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
g = P.datasets.manhattan_grid(side, side, 50 * side)
if len(sys.argv) > 2 and sys.argv[2] == "identity":
    g.edge_sqrt_info[:] = np.eye(6).reshape(1, 36)
G = P.Graph.from_dataset(g)
bytes_per_edge = 1832 if not (len(sys.argv) > 2 and sys.argv[2] == "identity") else 1544
for _ in range(5):
    cost, ms = G.linearize()
    print(f"linearize[{'identity' if bytes_per_edge == 1544 else 'info'}, occ={os.environ.get('PGO_LIN_OCC', '2')}]: {g.n_edges} edges {ms:.3f} ms -> "
          f"{g.n_edges / ms / 1e3:.1f} M edges/s, {g.n_edges * bytes_per_edge / ms / 1e6:.0f} GB/s algorithmic", flush=True)
x = np.random.default_rng(0).normal(size=(g.n_poses, 6))
for _ in range(2):
    y, ms = G.spmv(x, None, 4)
    print(f"spmv: {G.hessian_blocks()} blocks {ms / 4:.3f} ms", flush=True)
G.close()
