#!/usr/bin/env python3
"""Build tests/golden/path_plot_fixture.npz from the reference's other committed Ceres outputs.

Runs ONLY in the build container (needs /root/reference); the .npz it writes is committed so that tests never read
/root/reference at run time.

Sources (under /root/reference/src/POSE_GRAPH_CERES_PLUS/path_plot): pose_graph02_before.txt / pose_graph02.txt and
pose_graph08_before.txt / pose_graph08.txt -- the trajectories of KITTI sequences 02 and 08 written by OutputPoses()
(test/pose_graph_ceres_plus_finial.cpp:547-567, "id x y z qx qy qz qw") before and after ceres::Solve.  The loop edges
of those runs were not committed; the odometry edges are rebuilt from the "before" file exactly as the reference builds
them (:206-224).  The numbers are used as printed (see make_kitti00_fixture.py on the quaternion columns).
"""
import os

import numpy as np

REF = "/root/reference/src/POSE_GRAPH_CERES_PLUS/path_plot"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "path_plot_fixture.npz")

if __name__ == "__main__":
    out = {}
    for seq in ("02", "08"):
        b = np.loadtxt(os.path.join(REF, f"pose_graph{seq}_before.txt"))
        a = np.loadtxt(os.path.join(REF, f"pose_graph{seq}.txt"))
        assert np.array_equal(b[:, 0], np.arange(len(b))) and np.array_equal(a[:, 0], np.arange(len(a))) and len(a) == len(b)
        out[f"before{seq}"] = b[:, 1:8]
        out[f"after{seq}"] = a[:, 1:8]
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()}, os.path.getsize(OUT), "bytes")
