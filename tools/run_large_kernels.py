#!/usr/bin/env python3
"""Launch the two hot kernels (linearize, block-SpMV) a few times on a large Manhattan grid graph;
the target of the `ncu --set full` captures committed under profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import posegraph_ceres_b200 as P  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
g = P.datasets.manhattan_grid(side, side, 50 * side)
G = P.Graph.from_dataset(g)
for _ in range(4):
    cost, ms = G.linearize()
    print(f"linearize: {g.n_edges} edges {ms:.3f} ms -> {g.n_edges / ms / 1e3:.1f} M edges/s", flush=True)
x = np.random.default_rng(0).normal(size=(g.n_poses, 6))
for _ in range(2):
    y, ms = G.spmv(x, None, 4)
    print(f"spmv: {G.hessian_blocks()} blocks {ms / 4:.3f} ms", flush=True)
G.close()
