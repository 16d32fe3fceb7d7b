#!/usr/bin/env python3
"""One GPU: the stream-ordered PCG of the multi-GPU path (PGO_FORCE_STREAM_PCG=1, all-reduce is a no-op at world 1)
against the persistent PCG kernel: same LM iterates, and how long each takes."""
import os
import sys
import time

os.environ["PGO_FORCE_STREAM_PCG"] = "1" if len(sys.argv) > 1 and sys.argv[1] == "stream" else ""
if not os.environ["PGO_FORCE_STREAM_PCG"]:
    del os.environ["PGO_FORCE_STREAM_PCG"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import posegraph_ceres_b200 as P  # noqa: E402

for name, g in (("sphere60", P.datasets.sphere(6, 10, None)), ("sphere200", P.datasets.sphere(10, 20, None))):
    G = P.Graph.from_dataset(g)
    o = P.default_options()
    o.linear_solver_type = P.LINEAR_PCG_BLOCK_JACOBI
    o.pcg_tolerance = 1e-12
    o.pcg_max_iterations = 100000
    t0 = time.perf_counter()
    s, its = G.solve(o)
    dt = time.perf_counter() - t0
    p = G.get_poses()
    print(f"{name}: {'stream' if 'PGO_FORCE_STREAM_PCG' in os.environ else 'persistent'} PCG: {s.num_iterations - 1} LM iterations, {s.total_pcg_iterations} PCG iterations, "
          f"{dt:.3f} s, final cost {s.final_cost:.9f}, checksum {np.abs(p).sum():.9f}", flush=True)
    G.close()
