#!/usr/bin/env python3
"""Build tests/golden/kitti00_fixture.npz from the reference's own result files.

Runs ONLY in the build container (needs /root/reference); the .npz it writes is
committed so that tests / bench never read /root/reference at run time.

Sources (all under /root/reference/src/POSE_GRAPH_CERES_PLUS):
  result/trajectory/trajectory_origin.txt             poses BEFORE optimisation, written by
        OutputPoses() at test/pose_graph_ceres_plus_finial.cpp:137 ("id x y z qx qy qz qw")
  result/trajectory/trajectory_update_y_not_constant.txt  poses AFTER ceres::Solve (same file :147)
  result/Edges/edges_for_loop.txt                      accepted loop edges "begin end" (:250-251)
  config/Edge_Candidates_index.txt                     candidate topology read by
        getEdegsCandidateIndex() (include/ReadEdges.h:9-48)

Quaternion columns.  The files are used AS PRINTED: columns 5-8 are what OutputPoses() streams
as q.x() q.y() q.z() q.w(), i.e. the values ceres::Solve saw in the Eigen coeffs() slots x,y,z,w.
Physically those four numbers are (w, x, y, z) of the camera rotation -- row 0 reads "1 0 0 0" and
only that reading gives a body-frame step R_i^T (p_{i+1} - p_i) = (0.00, -0.01, +0.82) m, a car
driving along the camera's +z axis -- because the revision of the reference that produced the
committed result files filled Eigen::Quaterniond through its coefficient-pointer constructor with a
{w,x,y,z} array (the same slip is still visible in src/GroundTruth.cc:62-63, loadPoses1).  For
solver parity only the numbers Ceres optimised matter, and tests/test_oracle_cpu.py::
test_reference_ceres_output_is_stationary_for_the_oracle confirms the as-printed reading: with it
the oracle's gradient at the reference's optimised trajectory vanishes (<= 5e-3, file rounding) on
every pose without a loop edge, while the physically "corrected" w-first reading leaves a median
gradient of 5.7e-2 and a maximum of 0.45 on the same poses.

What is NOT in the reference tree: the PnP-estimated loop measurements (they were computed
from KITTI images at run time, test/pose_graph_ceres_plus_finial.cpp:203-255, and never
written out).  The fixture therefore stores topology + before/after poses only; loop
measurements are synthesised deterministically by posegraph_ceres_b200.datasets.kitti00().
"""
import os
import numpy as np

REF = "/root/reference/src/POSE_GRAPH_CERES_PLUS"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kitti00_fixture.npz")


def main():
    before = np.loadtxt(f"{REF}/result/trajectory/trajectory_origin.txt")
    after = np.loadtxt(f"{REF}/result/trajectory/trajectory_update_y_not_constant.txt")
    loops = np.loadtxt(f"{REF}/result/Edges/edges_for_loop.txt", dtype=np.int32)
    assert before.shape == (4541, 8) and after.shape == (4541, 8)
    assert (before[:, 0] == np.arange(4541)).all()
    # candidate lists: "cur prev cand..." one line per frame 1..4540
    cand_ptr = [0]
    cand_idx = []
    cand_cur = []
    with open(f"{REF}/config/Edge_Candidates_index.txt") as f:
        for line in f:
            tok = [int(t) for t in line.split()]
            if not tok:
                continue
            cand_cur.append(tok[0])
            cand_idx.extend(tok[1:])
            cand_ptr.append(len(cand_idx))
    np.savez_compressed(
        OUT,
        poses_before=before[:, 1:8].astype(np.float64),   # x y z qx qy qz qw
        poses_after=after[:, 1:8].astype(np.float64),
        loop_edges=loops,                                  # (639, 2) begin, end
        cand_cur=np.asarray(cand_cur, np.int32),
        cand_ptr=np.asarray(cand_ptr, np.int32),
        cand_idx=np.asarray(cand_idx, np.int32),
    )
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
