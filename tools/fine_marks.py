#!/usr/bin/env python3
"""Debug aid: stage marks inside the staged factor step (library built with -DPGO_CHOL_FINE_MARKS, PGO_TIMELINE=1)."""
import os
import sys
os.environ["PGO_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import posegraph_ceres_b200 as P  # noqa: E402

g = P.datasets.manhattan_loop() if sys.argv[1] == "manhattan" else P.datasets.kitti00()
o = P.default_options()
o.max_num_iterations = 3
G = P.Graph.from_dataset(g)
s, _ = G.solve(o)
G.close()
