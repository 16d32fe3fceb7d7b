"""Pose-graph workloads (host side, numpy): the configurations BASELINE.json names.

Conventions follow the reference's types (REF = /root/reference/src/POSE_GRAPH_CERES_PLUS):
  pose  = (x, y, z, qx, qy, qz, qw)            Pose3d, REF/include/types.h:16-21
  edge  = (id_begin a, id_end b, t_be, info)   Edge3d, REF/include/types.h:30-45; t_be = T_a^-1 T_b
  sqrt_information = information.llt().matrixL()   REF/test/pose_graph_ceres_plus_finial.cpp:475
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


@dataclass
class PoseGraph:
    name: str
    poses: np.ndarray        # (N, 7) float64 initial estimate
    edge_ids: np.ndarray     # (E, 2) int32  (a = id_begin, b = id_end)
    edge_meas: np.ndarray    # (E, 7) float64
    edge_sqrt_info: np.ndarray  # (E, 36) float64, row-major 6x6
    pose_const: np.ndarray   # (N,) uint8
    truth: np.ndarray | None = None

    @property
    def n_poses(self) -> int:
        return int(self.poses.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.edge_ids.shape[0])


# ---------------------------------------------------------------- quaternion helpers (x,y,z,w)
def qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def qconj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def qrot(q, v):
    u = q[..., :3]
    w = q[..., 3:4]
    uv = 2.0 * np.cross(u, v)
    return v + w * uv + np.cross(u, uv)


def qexp(rv):
    """rotation vector -> unit quaternion"""
    rv = np.asarray(rv, dtype=np.float64)
    th = np.linalg.norm(rv, axis=-1, keepdims=True)
    half = 0.5 * th
    k = np.where(th > 1e-12, np.sin(half) / np.maximum(th, 1e-300), 0.5)
    return np.concatenate([k * rv, np.cos(half)], axis=-1)


def relative_pose(pa, pb):
    """t_ab = T_a^-1 T_b for (...,7) pose arrays."""
    qa_inv = qconj(pa[..., 3:7])
    p = qrot(qa_inv, pb[..., 0:3] - pa[..., 0:3])
    q = qmul(qa_inv, pb[..., 3:7])
    return np.concatenate([p, q], axis=-1)


def compose(pa, t):
    """T_a * t"""
    p = pa[..., 0:3] + qrot(pa[..., 3:7], t[..., 0:3])
    q = qmul(pa[..., 3:7], t[..., 3:7])
    return np.concatenate([p, q], axis=-1)


def perturb(t, rng, sigma_t, sigma_r):
    """right-multiply a small random motion"""
    n = t.shape[0]
    noise = np.concatenate([rng.normal(0.0, sigma_t, (n, 3)), qexp(rng.normal(0.0, sigma_r, (n, 3)))], axis=-1)
    out = compose(t, noise)
    out[:, 3:7] /= np.linalg.norm(out[:, 3:7], axis=1, keepdims=True)
    return out


def _identity_sqrt_info(n):
    return np.tile(np.eye(6).reshape(1, 36), (n, 1))


def _sqrt_info_from_sigmas(n, sigma_t, sigma_r, rng=None, correlated=False):
    """information.llt().matrixL() for diag (or mildly correlated) information matrices."""
    d = np.array([1.0 / sigma_t] * 3 + [1.0 / sigma_r] * 3)
    if not correlated:
        return np.tile(np.diag(d).reshape(1, 36), (n, 1))
    out = np.empty((n, 36))
    for i in range(n):
        a = rng.normal(0.0, 0.15, (6, 6))
        info = np.diag(d) @ (np.eye(6) + a @ a.T) @ np.diag(d)
        out[i] = np.linalg.cholesky(info).reshape(36)
    return out


# ---------------------------------------------------------------- configs[1]: KITTI-00
def kitti00(synthetic_loops: bool = False, loop_sigma_t: float = 0.02, loop_sigma_r: float = 0.002, seed: int = 0) -> PoseGraph:
    """KITTI-00 pose graph as the reference builds it (4541 poses, 4540 odometry + 639 loop edges).

    Vertices: trajectory_origin.txt (the reference's own dump of the initial poses).
    Odometry edge for frame i>0: begin=i, end=i-1, t_be = Tcw_i * Twc_{i-1}, information = I
      (REF/test/pose_graph_ceres_plus_finial.cpp:206-224).
    Loop edges: the 639 (begin, end) pairs the reference accepted (edges_for_loop.txt), all of them
      members of Edge_Candidates_index.txt.  Their PnP measurements are not in the reference tree; the
      fixture holds the values RECOVERED from the reference's own Ceres output (stationarity of
      trajectory_update_y_not_constant.txt, see tests/golden/make_kitti00_fixture.py).
      synthetic_loops=True instead draws them as the optimised relative pose plus seeded noise.
    Edge order follows the reference: per frame, the odometry edge then that frame's loop edge.
    """
    fx = np.load(os.path.join(_GOLDEN, "kitti00_fixture.npz"))
    before = fx["poses_before"].copy()
    after = fx["poses_after"]
    loops = fx["loop_edges"]
    n = before.shape[0]
    odo_ids = np.stack([np.arange(1, n), np.arange(0, n - 1)], axis=1).astype(np.int32)
    odo_meas = relative_pose(before[odo_ids[:, 0]], before[odo_ids[:, 1]])
    if synthetic_loops:
        rng = np.random.default_rng(seed)
        loop_meas = relative_pose(after[loops[:, 0]], after[loops[:, 1]])
        loop_meas[:, 3:7] /= np.linalg.norm(loop_meas[:, 3:7], axis=1, keepdims=True)
        loop_meas = perturb(loop_meas, rng, loop_sigma_t, loop_sigma_r)
    else:
        loop_meas = fx["loop_meas"].copy()
    ids = np.concatenate([odo_ids, loops.astype(np.int32)], axis=0)
    meas = np.concatenate([odo_meas, loop_meas], axis=0)
    # reference order: sort by begin frame, odometry (end = begin-1) first
    key = ids[:, 0].astype(np.int64) * 2 + (ids[:, 0] - ids[:, 1] != 1)
    order = np.argsort(key, kind="stable")
    ids, meas = ids[order], meas[order]
    const = np.zeros(n, np.uint8)
    const[0] = 1   # SetParameterBlockConstant(poses->begin()), :525-527
    return PoseGraph("kitti00", before, ids, meas, _identity_sqrt_info(len(ids)), const, truth=after.copy())


# ---------------------------------------------------------------- configs[0]: Manhattan loop
def manhattan_loop(n_poses: int = 100, n_edges: int = 120, seed: int = 1) -> PoseGraph:
    """Square loop in the plane, unit steps, 90-degree turns at the corners; n_edges - (n_poses-1)
    loop closures between poses that revisit the same place on the second lap."""
    rng = np.random.default_rng(seed)
    side = max(n_poses // 8, 2)          # two laps of a square with `side` steps per side
    truth = np.zeros((n_poses, 7))
    pos = np.zeros(3)
    yaw = 0.0
    for i in range(n_poses):
        truth[i, 0:3] = pos
        truth[i, 3:7] = qexp(np.array([0.0, 0.0, yaw]))
        if (i + 1) % side == 0:
            yaw += np.pi / 2
        pos = pos + np.array([np.cos(yaw), np.sin(yaw), 0.0])
    lap = 4 * side
    odo = np.stack([np.arange(1, n_poses), np.arange(0, n_poses - 1)], axis=1)
    n_loop = n_edges - (n_poses - 1)
    cand = np.arange(lap, n_poses)
    pick = cand[np.linspace(0, len(cand) - 1, n_loop).astype(int)]
    loops = np.stack([pick, pick - lap], axis=1)
    ids = np.concatenate([odo, loops], axis=0).astype(np.int32)
    meas = perturb(relative_pose(truth[ids[:, 0]], truth[ids[:, 1]]), rng, 0.02, 0.01)
    # initial guess: integrate the noisy odometry (begin = i, end = i-1  =>  T_i = T_{i-1} * t^-1)
    init = np.zeros_like(truth)
    init[0] = truth[0]
    for i in range(1, n_poses):
        t = meas[i - 1]
        tinv_q = qconj(t[3:7])
        tinv = np.concatenate([-qrot(tinv_q, t[0:3]), tinv_q])
        init[i] = compose(init[i - 1], tinv)
    const = np.zeros(n_poses, np.uint8)
    const[0] = 1
    sqrt_info = _sqrt_info_from_sigmas(len(ids), 0.02, 0.01)
    return PoseGraph(f"manhattan{n_poses}", init, ids, meas, sqrt_info, const, truth=truth)


# ---------------------------------------------------------------- configs[2]: sphere2500-like
def sphere(n_rings: int = 50, per_ring: int = 50, n_edges: int | None = 9799, radius: float = 50.0,
           seed: int = 2, sigma_t: float = 0.05, sigma_r: float = 0.01, init_sigma_t: float = 0.5,
           init_sigma_r: float = 0.05, correlated_info: bool = True) -> PoseGraph:
    """Spiral over a sphere (the g2o sphere generator's topology): odometry along the spiral plus
    constraints to the three nearest poses of the previous ring (i-per_ring, i-per_ring+-1)."""
    rng = np.random.default_rng(seed)
    n = n_rings * per_ring
    k = np.arange(n)
    az = 2 * np.pi * (k % per_ring) / per_ring
    el = -np.pi / 2 * 0.9 + (np.pi * 0.9) * (k / (n - 1))
    pos = radius * np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    q = qmul(qexp(np.stack([np.zeros(n), np.zeros(n), az], axis=1)),
             qexp(np.stack([np.zeros(n), -el, np.zeros(n)], axis=1)))
    truth = np.concatenate([pos, q], axis=1)
    e = [np.stack([k[1:], k[:-1]], axis=1)]
    for off in (per_ring, per_ring - 1, per_ring + 1):
        e.append(np.stack([k[off:], k[:-off]], axis=1))
    ids = np.concatenate(e, axis=0).astype(np.int32)
    if n_edges is not None:
        ids = ids[:n_edges]
    meas = perturb(relative_pose(truth[ids[:, 0]], truth[ids[:, 1]]), rng, sigma_t, sigma_r)
    init = perturb(truth, rng, init_sigma_t, init_sigma_r)
    init[0] = truth[0]
    const = np.zeros(n, np.uint8)
    const[0] = 1
    sqrt_info = _sqrt_info_from_sigmas(len(ids), sigma_t, sigma_r, rng, correlated=correlated_info)
    return PoseGraph(f"sphere{n}", init, ids, meas, sqrt_info, const, truth=truth)


# ---------------------------------------------------------------- configs[3]: Manhattan grid
def manhattan_grid(rows: int = 1000, cols: int = 1000, n_loops: int = 50000, seed: int = 3,
                   sigma_t: float = 0.02, sigma_r: float = 0.005, init_sigma_t: float = 0.1,
                   init_sigma_r: float = 0.02) -> PoseGraph:
    """Boustrophedon walk over a rows x cols grid (rows*cols poses): rows*cols-1 odometry edges,
    (rows-1)*cols edges to the cell one row below (=> ~2 edges per pose) and n_loops random
    long-range closures.  1000 x 1000 + 50k is the 1M-pose / 2M-edge configuration."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    k = np.arange(n)
    r = k // cols
    c = np.where(r % 2 == 0, k % cols, cols - 1 - (k % cols))
    yaw = np.where(r % 2 == 0, 0.0, np.pi)
    truth = np.concatenate([np.stack([c.astype(np.float64), r.astype(np.float64), np.zeros(n)], axis=1),
                            qexp(np.stack([np.zeros(n), np.zeros(n), yaw], axis=1))], axis=1)
    odo = np.stack([k[1:], k[:-1]], axis=1)
    # cell (r, c) is pose r*cols + (c or cols-1-c); the cell below it is in row r-1
    idx_of = lambda rr, cc: rr * cols + np.where(rr % 2 == 0, cc, cols - 1 - cc)
    rr, cc = np.meshgrid(np.arange(1, rows), np.arange(cols), indexing="ij")
    up = np.stack([idx_of(rr, cc).ravel(), idx_of(rr - 1, cc).ravel()], axis=1)
    up = up[np.abs(up[:, 0] - up[:, 1]) > 1]
    la = rng.integers(0, n, n_loops)
    lb = rng.integers(0, n, n_loops)
    keep = np.abs(la - lb) > 1
    loops = np.stack([np.maximum(la, lb)[keep], np.minimum(la, lb)[keep]], axis=1)
    ids = np.concatenate([odo, up, loops], axis=0).astype(np.int32)
    meas = perturb(relative_pose(truth[ids[:, 0]], truth[ids[:, 1]]), rng, sigma_t, sigma_r)
    init = perturb(truth, rng, init_sigma_t, init_sigma_r)
    init[0] = truth[0]
    const = np.zeros(n, np.uint8)
    const[0] = 1
    return PoseGraph(f"grid{rows}x{cols}", init, ids, meas,
                     _sqrt_info_from_sigmas(len(ids), sigma_t, sigma_r), const, truth=truth)


# ---------------------------------------------------------------- configs[4]: torus
def torus(n_poses: int = 100000, loop_fraction: float = 0.10, seed: int = 4, R: float = 40.0,
          r: float = 10.0, winds: int = 200, sigma_t: float = 0.02, sigma_r: float = 0.005,
          init_sigma_t: float = 0.1, init_sigma_r: float = 0.02) -> PoseGraph:
    """Toroidal spiral with loop_fraction*n_poses uniformly random (long-range, 'dense') loop
    edges: high fill in J^T J, a stress test for the preconditioner."""
    rng = np.random.default_rng(seed)
    n = n_poses
    k = np.arange(n)
    u = 2 * np.pi * k / n            # around the big circle
    v = 2 * np.pi * winds * k / n    # around the tube
    pos = np.stack([(R + r * np.cos(v)) * np.cos(u), (R + r * np.cos(v)) * np.sin(u), r * np.sin(v)], axis=1)
    q = qmul(qexp(np.stack([np.zeros(n), np.zeros(n), u], axis=1)),
             qexp(np.stack([np.zeros(n), v, np.zeros(n)], axis=1)))
    truth = np.concatenate([pos, q], axis=1)
    odo = np.stack([k[1:], k[:-1]], axis=1)
    per_wind = n // winds
    ring = np.stack([k[per_wind:], k[:-per_wind]], axis=1)
    n_loops = int(loop_fraction * n)
    la = rng.integers(0, n, n_loops)
    lb = rng.integers(0, n, n_loops)
    keep = np.abs(la - lb) > 1
    loops = np.stack([np.maximum(la, lb)[keep], np.minimum(la, lb)[keep]], axis=1)
    ids = np.concatenate([odo, ring, loops], axis=0).astype(np.int32)
    meas = perturb(relative_pose(truth[ids[:, 0]], truth[ids[:, 1]]), rng, sigma_t, sigma_r)
    init = perturb(truth, rng, init_sigma_t, init_sigma_r)
    init[0] = truth[0]
    const = np.zeros(n, np.uint8)
    const[0] = 1
    return PoseGraph(f"torus{n}", init, ids, meas,
                     _sqrt_info_from_sigmas(len(ids), sigma_t, sigma_r), const, truth=truth)


# ---------------------------------------------------------------- g2o files (VERTEX_SE3:QUAT / EDGE_SE3:QUAT)
def write_g2o(g: PoseGraph, path: str) -> None:
    """Write the graph as a g2o file.  g2o stores the INFORMATION matrix (upper triangle, row-major);
    the reference turns it into sqrt_information = information.llt().matrixL() (REF/test/
    pose_graph_ceres_plus_finial.cpp:508), so information = S S^T is written for a lower-triangular S."""
    with open(path, "w") as f:
        for i in range(g.n_poses):
            f.write("VERTEX_SE3:QUAT %d %s\n" % (i, " ".join(repr(float(v)) for v in g.poses[i])))
        for e in range(g.n_edges):
            S = g.edge_sqrt_info[e].reshape(6, 6)
            if np.abs(np.triu(S, 1)).max() != 0.0:
                raise ValueError("write_g2o: sqrt_information must be lower triangular (an llt().matrixL())")
            info = S @ S.T
            up = [repr(float(info[r, c])) for r in range(6) for c in range(r, 6)]
            f.write("EDGE_SE3:QUAT %d %d %s %s\n" % (g.edge_ids[e, 0], g.edge_ids[e, 1],
                                                     " ".join(repr(float(v)) for v in g.edge_meas[e]), " ".join(up)))


def read_g2o(path: str, name: str | None = None) -> PoseGraph:
    """Read VERTEX_SE3:QUAT / EDGE_SE3:QUAT records; vertex ids are compacted in ascending order (the
    reference keeps poses in a std::map<int, Pose3d>) and the first vertex is held constant."""
    vid, vpose, eids, emeas, einfo = [], [], [], [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "VERTEX_SE3:QUAT":
                vid.append(int(t[1]))
                vpose.append([float(x) for x in t[2:9]])
            elif t[0] == "EDGE_SE3:QUAT":
                eids.append((int(t[1]), int(t[2])))
                emeas.append([float(x) for x in t[3:10]])
                up = [float(x) for x in t[10:31]]
                info = np.zeros((6, 6))
                k = 0
                for r in range(6):
                    for c in range(r, 6):
                        info[r, c] = info[c, r] = up[k]
                        k += 1
                einfo.append(np.linalg.cholesky(info).reshape(36))
    order = np.argsort(np.asarray(vid, np.int64), kind="stable")
    remap = {int(vid[k]): i for i, k in enumerate(order)}
    poses = np.asarray(vpose, np.float64)[order]
    ids = np.asarray([(remap[a], remap[b]) for a, b in eids], np.int32).reshape(-1, 2)
    const = np.zeros(len(vid), np.uint8)
    const[0] = 1
    return PoseGraph(name or os.path.basename(path), poses, ids, np.asarray(emeas, np.float64).reshape(-1, 7),
                     np.asarray(einfo, np.float64).reshape(-1, 36), const)


def shard_edges(g: PoseGraph, rank: int, world: int) -> PoseGraph:
    """Contiguous edge shard for rank `rank` of `world` (poses are replicated)."""
    e = g.n_edges
    lo = (e * rank) // world
    hi = (e * (rank + 1)) // world
    return PoseGraph(f"{g.name}[{rank}/{world}]", g.poses.copy(), g.edge_ids[lo:hi], g.edge_meas[lo:hi],
                     g.edge_sqrt_info[lo:hi], g.pose_const, g.truth)


@dataclass
class RowPartition:
    """One rank's slice under the owner-computes row partition of the multi-GPU path (host-side mirror of
    build_partition in csrc/pgo_b200.cu): the block rows of poses [lo, hi), every edge touching one of them, and halo
    copies of the other endpoints of cut edges.  Local ids: owned poses first (global id - lo), then the halo in
    ascending global id."""
    lo: int
    hi: int
    halo_gid: np.ndarray       # (n_halo,) global ids, ascending
    edge_sel: np.ndarray       # global ids of the local edges, ascending
    local: PoseGraph           # poses = owned + halo, edge_ids in LOCAL ids
    cost_edges: np.ndarray     # bool per local edge: its cost is accounted here (id_begin is owned)

    @property
    def n_own(self) -> int:
        return self.hi - self.lo


def partition_rows(g: PoseGraph, rank: int, world: int) -> RowPartition:
    n = g.n_poses
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    a, b = g.edge_ids[:, 0], g.edge_ids[:, 1]
    oa, ob = (a >= lo) & (a < hi), (b >= lo) & (b < hi)
    sel = np.nonzero(oa | ob)[0]
    halo = np.unique(np.concatenate([a[sel][~oa[sel]], b[sel][~ob[sel]]])).astype(np.int64)

    def local_of(gid):
        own = (gid >= lo) & (gid < hi)
        return np.where(own, gid - lo, (hi - lo) + np.searchsorted(halo, gid))

    ids = np.stack([local_of(a[sel]), local_of(b[sel])], axis=1).astype(np.int32)
    gids = np.concatenate([np.arange(lo, hi), halo])
    loc = PoseGraph(f"{g.name}[rows {rank}/{world}]", g.poses[gids].copy(), ids, g.edge_meas[sel].copy(),
                    g.edge_sqrt_info[sel].copy(), g.pose_const[gids].copy(), None if g.truth is None else g.truth[gids].copy())
    return RowPartition(lo, hi, halo, sel, loc, oa[sel])


# ---------------------------------------------------------------------------------------------------------------
# config/Edge_Candidates_index.txt: "cur cand cand ..." per line for frames 1..n-1, written by
# REF/src/POSE_GRAPH_CERES_PLUS/test/generate_edges_from_trajectory_origion.cpp:36-52 and read back by
# getEdegsCandidateIndex(), REF/src/POSE_GRAPH_CERES_PLUS/include/ReadEdges.h:9-48.
# ---------------------------------------------------------------------------------------------------------------
def write_edge_candidates(row_ptr, candidates, path: str):
    with open(path, "w") as f:
        for c in range(1, len(row_ptr) - 1):
            # every token is followed by one blank, like the reference's "outFile << v << \" \"" (:43-48)
            f.write("".join(f"{int(v)} " for v in [c, *candidates[row_ptr[c]:row_ptr[c + 1]]]) + "\n")


def read_edge_candidates(path: str):
    """-> (row_ptr[n+1], candidates) in the layout pgo_edge_candidates returns (frame 0 has no line)."""
    row_ptr, idx = [0, 0], []
    with open(path) as f:
        for line in f:
            tok = [int(t) for t in line.split()]
            if not tok:
                continue
            if tok[0] != len(row_ptr) - 1:
                raise ValueError(f"{path}: expected frame {len(row_ptr) - 1}, found {tok[0]}")
            idx.extend(tok[1:])
            row_ptr.append(len(idx))
    return np.asarray(row_ptr, np.int64), np.asarray(idx, np.int32)
