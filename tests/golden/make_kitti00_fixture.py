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
written out).  They are RECOVERED here from the reference's own Ceres output: the optimised
trajectory is a stationary point of the cost, every loop edge's begin frame carries exactly one
loop edge (a frame adds at most one, :238-289) and is never the end frame of another, so the six
gradient equations at the begin pose determine the edge's six measurement parameters (Newton on a
two-pose sub-problem of the CPU oracle).  The six equations at the END pose are not used in the fit:
tests/test_oracle_cpu.py checks that they hold too (to the files' rounding), and that the oracle's LM
run from trajectory_origin with these measurements lands on trajectory_update_y_not_constant.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

REF = "/root/reference/src/POSE_GRAPH_CERES_PLUS"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kitti00_fixture.npz")


def recover_loop_measurements(before, after, loops):
    """Loop-edge measurements t_be from the stationarity of the reference's optimised trajectory (docstring)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    import posegraph_ceres_b200.datasets as D
    n = before.shape[0]
    odo_ids = np.stack([np.arange(1, n), np.arange(0, n - 1)], axis=1).astype(np.int32)
    odo_meas = D.relative_pose(before[odo_ids[:, 0]], before[odo_ids[:, 1]])
    eye = np.eye(6).reshape(1, 36)
    const = np.zeros(n, np.uint8)
    const[0] = 1
    go = D.PoseGraph("odo", before, odo_ids, odo_meas, np.tile(eye, (len(odo_ids), 1)), const)
    _, _, godo, _ = O.evaluate(go, poses=after)        # odometry-only gradient at the reference optimum
    assert len(set(loops[:, 0].tolist())) == len(loops) and not (set(loops[:, 0].tolist()) & set(loops[:, 1].tolist()))

    def meas_from(base, d):
        return D.compose(base[None], np.concatenate([d[:3], D.qexp(d[3:])])[None])[0]

    def begin_gradient(a, b, m):
        sub = D.PoseGraph("e", after[[a, b]], np.array([[0, 1]], np.int32), m[None], eye, np.zeros(2, np.uint8))
        return O.evaluate(sub)[2][0]

    out = np.zeros((len(loops), 7))
    worst = 0.0
    for e, (a, b) in enumerate(loops):
        base = D.relative_pose(after[a][None], after[b][None])[0]
        base[3:] /= np.linalg.norm(base[3:])
        target = -godo[a]
        d = np.zeros(6)
        for _ in range(20):
            f = begin_gradient(a, b, meas_from(base, d)) - target
            if np.linalg.norm(f) < 1e-13:
                break
            J = np.zeros((6, 6))
            h = 1e-7
            for k in range(6):
                dd = d.copy()
                dd[k] += h
                J[:, k] = (begin_gradient(a, b, meas_from(base, dd)) - (f + target)) / h
            d = d + np.linalg.lstsq(J, -f, rcond=1e-10)[0]
        out[e] = meas_from(base, d)
        worst = max(worst, float(np.linalg.norm(begin_gradient(a, b, out[e]) - target)))
    return out, worst


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
    loop_meas, fit = recover_loop_measurements(before[:, 1:8].astype(np.float64), after[:, 1:8].astype(np.float64), loops)
    print("recovered %d loop measurements, worst begin-pose gradient residual %.2e" % (len(loop_meas), fit))
    np.savez_compressed(
        OUT,
        loop_meas=loop_meas,                               # (639, 7) t_be of the loop edges, recovered (see docstring)
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
