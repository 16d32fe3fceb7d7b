"""ctypes binding of the CPU oracle (oracle/pgo_oracle.c). TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libpgo_oracle.so")

LOSS_TRIVIAL, LOSS_HUBER, LOSS_CAUCHY = 0, 1, 2


class Options(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int), ("jacobi_scaling", C.c_int),
                ("loss_type", C.c_int), ("loss_a", C.c_double), ("ordering", C.c_int)]


class Iteration(C.Structure):
    _fields_ = [("iteration", C.c_int), ("step_is_valid", C.c_int), ("step_is_successful", C.c_int),
                ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double),
                ("gradient_norm", C.c_double), ("step_norm", C.c_double),
                ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double)]


class Summary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("num_successful_steps", C.c_int), ("num_unsuccessful_steps", C.c_int),
                ("num_iterations", C.c_int), ("termination_type", C.c_int), ("message", C.c_char * 160),
                ("time_total_s", C.c_double), ("time_residual_s", C.c_double),
                ("time_jacobian_s", C.c_double), ("time_linear_solver_s", C.c_double),
                ("factor_nnz_blocks", C.c_longlong), ("num_jacobian_evals", C.c_int),
                ("num_residual_evals", C.c_int)]


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "pgo_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.oracle_default_options.argtypes = [C.POINTER(Options)]
        _lib.oracle_evaluate.restype = C.c_int
        _lib.oracle_solve.restype = C.c_int
        _lib.oracle_normal_solve.restype = C.c_int
        _lib.oracle_edge_candidates.restype = C.c_longlong
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def set_num_threads(n: int):
    lib().oracle_set_num_threads(C.c_int(n))


class edge_losses:
    """with oracle_py.edge_losses(types, scales): ...  -- per-residual-block loss functions for the calls inside"""

    def __init__(self, types, scales):
        self.t = np.ascontiguousarray(types, np.int32)
        self.a = np.ascontiguousarray(scales, np.float64)

    def __enter__(self):
        lib().oracle_set_edge_losses(C.c_int(self.t.size), _p(self.t, C.c_int), _p(self.a))
        return self

    def __exit__(self, *exc):
        lib().oracle_set_edge_losses(C.c_int(0), None, None)
        return False


def default_options() -> Options:
    o = Options()
    lib().oracle_default_options(C.byref(o))
    return o


def evaluate(g, poses=None, loss_type=LOSS_HUBER, loss_a=1.0, want_jac=True):
    poses = np.ascontiguousarray(g.poses if poses is None else poses, np.float64)
    n, e = g.n_poses, g.n_edges
    cost = C.c_double()
    res = np.zeros(6 * e)
    grad = np.zeros(6 * n)
    jac = np.zeros((e, 2, 36)) if want_jac else None
    ids = np.ascontiguousarray(g.edge_ids, np.int32)
    rc = lib().oracle_evaluate(C.c_int(n), _p(poses), _p(np.ascontiguousarray(g.pose_const, np.uint8), C.c_ubyte),
                               C.c_int(e), _p(ids, C.c_int), _p(np.ascontiguousarray(g.edge_meas)),
                               _p(np.ascontiguousarray(g.edge_sqrt_info)), C.c_int(loss_type), C.c_double(loss_a),
                               C.byref(cost), _p(res), _p(grad), _p(jac))
    assert rc == 0
    return cost.value, res.reshape(e, 6), grad.reshape(n, 6), jac


def plus(poses, delta):
    poses = np.ascontiguousarray(poses, np.float64)
    delta = np.ascontiguousarray(delta, np.float64)
    out = np.empty_like(poses)
    lib().oracle_plus(C.c_int(poses.shape[0]), _p(poses), _p(delta), _p(out))
    return out


def solve(g, options: Options | None = None, max_log=2048):
    o = options or default_options()
    poses = np.ascontiguousarray(g.poses, np.float64).copy()
    s = Summary()
    log = (Iteration * max_log)()
    ids = np.ascontiguousarray(g.edge_ids, np.int32)
    rc = lib().oracle_solve(C.c_int(g.n_poses), _p(poses), _p(np.ascontiguousarray(g.pose_const, np.uint8), C.c_ubyte),
                            C.c_int(g.n_edges), _p(ids, C.c_int), _p(np.ascontiguousarray(g.edge_meas)),
                            _p(np.ascontiguousarray(g.edge_sqrt_info)), C.byref(o), C.byref(s), log, C.c_int(max_log))
    assert rc == 0
    its = [log[i] for i in range(min(s.num_iterations, max_log))]
    return poses, s, its


def normal_solve(g, jac, d, rhs, ordering=1):
    y = np.zeros(6 * g.n_poses)
    ids = np.ascontiguousarray(g.edge_ids, np.int32)
    rc = lib().oracle_normal_solve(C.c_int(g.n_poses), _p(np.ascontiguousarray(g.pose_const, np.uint8), C.c_ubyte),
                                   C.c_int(g.n_edges), _p(ids, C.c_int), _p(np.ascontiguousarray(jac, np.float64)),
                                   _p(np.ascontiguousarray(d, np.float64)), _p(np.ascontiguousarray(rhs, np.float64)),
                                   _p(y), C.c_int(ordering))
    return rc, y.reshape(-1, 6)


def edge_candidates(positions, search_radius=6.0, min_frame_gap=100):
    """(row_ptr[n+1], candidates): the lists generate_edges_from_trajectory_origion.cpp writes per frame."""
    pos = np.ascontiguousarray(positions, np.float64)
    n = int(pos.shape[0])
    row_ptr = np.zeros(n + 1, np.int64)
    args = (C.c_int(n), _p(pos), C.c_double(search_radius), C.c_int(min_frame_gap), _p(row_ptr, C.c_longlong))
    total = lib().oracle_edge_candidates(*args, None, C.c_longlong(0))
    idx = np.zeros(max(total, 1), np.int32)
    lib().oracle_edge_candidates(*args, _p(idx, C.c_int), C.c_longlong(idx.size))
    return row_ptr, idx[:total]


# ---------------------------------------------------------------------------------------------------------------
# oracle/_ref: the REFERENCE's own cost functor (PoseGraph3dError.h compiled from /root/reference, see ref_functor.cpp)
# ---------------------------------------------------------------------------------------------------------------
_REF_LIB = os.path.join(_HERE, "_ref", "libref_functor.so")
_REF_HEADER = "/root/reference/src/POSE_GRAPH_CERES_PLUS/include/PoseGraph3dError.h"
_ref = None


def ref_functor():
    """ctypes handle of oracle/_ref/libref_functor.so, built on demand when the reference tree is present; None otherwise."""
    global _ref
    if _ref is None:
        if os.path.exists(_REF_HEADER):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        if not os.path.exists(_REF_LIB):
            return None
        _ref = C.CDLL(_REF_LIB)
    return _ref


def ref_edge_jacobian(pose_a, pose_b, meas, sqrt_info):
    """(residual[6], jacobian[6][14]) of one edge from the reference's functor; the Jacobian is w.r.t. the ambient
    parameters (p_a[3], q_a[4], p_b[3], q_b[4]), as AutoDiffCostFunction<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4> sees them."""
    a = [np.ascontiguousarray(v, np.float64) for v in (pose_a, pose_b, meas, sqrt_info)]
    res, jac = np.zeros(6), np.zeros((6, 14))
    ref_functor().ref_edge_jacobian(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(res), _p(jac))
    return res, jac


def ref_edge_residual(pose_a, pose_b, meas, sqrt_info):
    a = [np.ascontiguousarray(v, np.float64) for v in (pose_a, pose_b, meas, sqrt_info)]
    res = np.zeros(6)
    ref_functor().ref_edge_residual(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(res))
    return res


_REF_READER = os.path.join(_HERE, "_ref", "ref_read_edges")


def ref_read_edge_candidates(candidate_file):
    """Parse `candidate_file` with the REFERENCE's own reader (ReadEdges.h:9-48 compiled into oracle/_ref/ref_read_edges):
    returns {key: [candidates]} exactly as getEdegsCandidateIndex() does, or None when oracle/_ref is unavailable."""
    import shutil
    import tempfile
    if ref_functor() is None or not os.path.exists(_REF_READER):
        return None
    with tempfile.TemporaryDirectory() as d:            # the reader opens ../config/Edge_Candidates_index.txt
        os.makedirs(os.path.join(d, "bin"))
        os.makedirs(os.path.join(d, "config"))
        shutil.copy(candidate_file, os.path.join(d, "config", "Edge_Candidates_index.txt"))
        out = subprocess.run([_REF_READER], cwd=os.path.join(d, "bin"), capture_output=True, text=True, check=True).stdout
    result = {}
    for line in out.splitlines():
        key, _, rest = line.partition(":")
        result[int(key)] = [int(t) for t in rest.split()]
    return result


_REF_GENERATOR = os.path.join(_HERE, "_ref", "ref_generate_edges")


def ref_generate_edge_candidates(positions, search_radius=6, want_file=False):
    """Run the REFERENCE's own candidate generator (test/generate_edges_from_trajectory_origion.cpp compiled into
    oracle/_ref/ref_generate_edges, stand-ins for OpenCV / yaml under ref_shim/gen) on `positions` [n][3].
    Returns (row_ptr[n+1], candidates) in pgo_edge_candidates' layout (and the raw file bytes with want_file), or None
    when oracle/_ref is unavailable.  The radius is an integer (Config::get<int>), the frame gap is the reference's 100."""
    import tempfile
    if ref_functor() is None or not os.path.exists(_REF_GENERATOR):
        return None
    pos = np.ascontiguousarray(positions, np.float64)
    n = int(pos.shape[0])
    with tempfile.TemporaryDirectory() as d:            # the program writes ../config/Edge_Candidates_index.txt
        os.makedirs(os.path.join(d, "bin"))
        os.makedirs(os.path.join(d, "config"))
        traj = os.path.join(d, "trajectory.txt")        # format 1: x y z q_x q_y q_z q_w
        np.savetxt(traj, np.hstack([pos, np.tile([0.0, 0.0, 0.0, 1.0], (n, 1))]), fmt="%.17g")
        env = dict(os.environ, REF_TRAJECTORY=traj, REF_SEQUENCE_LENGTH=str(n), REF_SEARCH_RADIUS=str(int(search_radius)))
        subprocess.run([_REF_GENERATOR], cwd=os.path.join(d, "bin"), env=env, stdout=subprocess.DEVNULL, check=True)
        with open(os.path.join(d, "config", "Edge_Candidates_index.txt"), "rb") as f:
            raw = f.read()
    row_ptr, idx = [0, 0], []
    for line in raw.decode().splitlines():
        tok = [int(t) for t in line.split()]
        if not tok:
            continue
        assert tok[0] == len(row_ptr) - 1
        idx.extend(tok[1:])
        row_ptr.append(len(idx))
    while len(row_ptr) < n + 1:
        row_ptr.append(len(idx))
    out = (np.asarray(row_ptr, np.int64), np.asarray(idx, np.int32))
    return out + (raw,) if want_file else out
