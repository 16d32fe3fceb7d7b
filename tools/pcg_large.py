#!/usr/bin/env python3
"""Time the block-Jacobi PCG on the 1M-pose grid (2 LM iterations, PCG capped): persistent kernel vs stream-ordered path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import posegraph_ceres_b200 as P
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
g = P.datasets.manhattan_grid(side, side, 50 * side)
G = P.Graph.from_dataset(g)
o = P.default_options()
o.max_num_iterations = 2
o.pcg_max_iterations = 600
o.linear_solver_type = P.LINEAR_PCG_BLOCK_JACOBI
s, its = G.solve(o)
print(f"{'stream-ordered' if (os.environ.get('PGO_FORCE_STREAM_PCG') or g.n_poses >= 200000) else 'persistent'} PCG: {s.total_pcg_iterations} PCG iterations in {s.time_linear_solver_ms:.1f} ms -> "
      f"{s.time_linear_solver_ms / max(1, s.total_pcg_iterations):.4f} ms/iteration, final cost {s.final_cost:.6f}", flush=True)
G.close()
