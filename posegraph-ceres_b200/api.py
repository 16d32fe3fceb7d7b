"""ctypes binding of libpgo_b200.so (include/pgo_b200.h) -- the C-ABI drop-in for the reference's
ceres::Problem / ceres::Solve pose-graph path.  Used by tests/ and bench.py; the C++ host-side
mirror of the Ceres surface is include/ceres_b200/ceres.h.

There is no CPU fallback: loading fails loudly when the CUDA library has not been built, and every
compute entry point returns PGO_ERR_NO_DEVICE without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "libpgo_b200.so")

LOSS_TRIVIAL, LOSS_HUBER, LOSS_CAUCHY = 0, 1, 2
LINEAR_PCG_BLOCK_JACOBI, LINEAR_PCG_LEVEL_CHOLESKY, LINEAR_AUTO, LINEAR_PCG_AMG = 0, 1, 2, 3
CONVERGENCE, NO_CONVERGENCE, FAILURE = 0, 1, 2

EXPORTED_SYMBOLS = [
    "pgo_last_error", "pgo_abi_version", "pgo_device_count", "pgo_default_options", "pgo_graph_create",
    "pgo_graph_destroy", "pgo_graph_set_stream", "pgo_graph_num_poses", "pgo_graph_num_edges", "pgo_graph_set_poses",
    "pgo_graph_get_poses", "pgo_graph_snapshot_poses", "pgo_graph_restore_poses", "pgo_nccl_unique_id",
    "pgo_graph_create_partitioned", "pgo_graph_rank", "pgo_graph_world_size", "pgo_graph_num_local_poses",
    "pgo_graph_num_halo_poses", "pgo_graph_num_local_edges", "pgo_analyze_partition", "pgo_amg_aggregates", "pgo_graph_evaluate", "pgo_graph_linearize", "pgo_graph_get_hessian",
    "pgo_graph_spmv", "pgo_graph_linear_solve", "pgo_graph_solve", "pgo_solve_pose_graph",
    "pgo_analyze_structure", "pgo_release_cached_memory", "pgo_edge_candidates", "pgo_set_topology_cache", "pgo_graph_set_edge_losses",
]


class SolverOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int), ("jacobi_scaling", C.c_int),
                ("loss_type", C.c_int), ("loss_a", C.c_double), ("linear_solver_type", C.c_int),
                ("pcg_max_iterations", C.c_int), ("pcg_tolerance", C.c_double), ("pcg_num_ctas", C.c_int),
                ("direct_residual_accept", C.c_double), ("verbose", C.c_int),
                ("edge_loss_type", C.POINTER(C.c_int)), ("edge_loss_a", C.POINTER(C.c_double))]


class IterationSummary(C.Structure):
    _fields_ = [("iteration", C.c_int), ("step_is_valid", C.c_int), ("step_is_successful", C.c_int),
                ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double),
                ("gradient_norm", C.c_double), ("step_norm", C.c_double), ("relative_decrease", C.c_double),
                ("trust_region_radius", C.c_double), ("linear_solver_iterations", C.c_int),
                ("pcg_relative_residual", C.c_double)]


class SolverSummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("num_successful_steps", C.c_int),
                ("num_unsuccessful_steps", C.c_int), ("num_iterations", C.c_int), ("termination_type", C.c_int),
                ("message", C.c_char * 160), ("num_linearizations", C.c_int), ("num_cost_evaluations", C.c_int),
                ("total_pcg_iterations", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("time_total_s", C.c_double), ("time_setup_s", C.c_double), ("time_linearize_ms", C.c_double),
                ("time_linear_solver_ms", C.c_double), ("linear_solver_used", C.c_int),
                ("hessian_blocks", C.c_longlong), ("factor_blocks", C.c_longlong), ("factor_levels", C.c_int),
                ("amg_levels", C.c_int), ("amg_blocks", C.c_longlong), ("comm_calls", C.c_longlong),
                ("comm_bytes", C.c_longlong), ("comm_bytes_per_pcg_iteration", C.c_longlong),
                ("comm_calls_per_pcg_iteration", C.c_int), ("peer_exchanges", C.c_longlong), ("peer_bytes", C.c_longlong),
                ("peer_exchanges_per_pcg_iteration", C.c_int)]


class StructureInfo(C.Structure):
    _fields_ = [("variable_poses", C.c_int), ("hessian_blocks", C.c_longlong), ("factor_usable", C.c_int),
                ("factor_blocks", C.c_longlong), ("factor_levels", C.c_int), ("factor_max_degree", C.c_int),
                ("factor_tasks", C.c_longlong), ("analysis_seconds", C.c_double)]


class PartitionInfo(C.Structure):
    _fields_ = [("n_own", C.c_int), ("n_halo", C.c_int), ("n_local_edges", C.c_int), ("n_cut_edges", C.c_int),
                ("n_neighbours", C.c_int), ("send_total", C.c_int), ("recv_total", C.c_int),
                ("send_to", C.c_int * 64), ("recv_from", C.c_int * 64), ("amg_levels", C.c_int),
                ("level_nodes", C.c_int * 16), ("level_own", C.c_int * 16), ("level_halo", C.c_int * 16),
                ("level_replicated", C.c_int * 16), ("level_blocks", C.c_longlong * 16),
                ("level_send", C.c_int * 16), ("level_recv", C.c_int * 16),
                ("plan_checksum", C.c_ulonglong), ("recv_checksum", C.c_ulonglong), ("consistent", C.c_int)]


class PgoError(RuntimeError):
    pass


def build_library(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/build.sh); cross-compiles without a GPU."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_CSRC), "..", "include", "pgo_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["bash", os.path.join(_CSRC, "build.sh")])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PgoError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                           "(posegraph-ceres_b200/csrc/build.sh). There is no CPU fallback.")
        try:                      # torch ships the libnccl.so.2 this library is linked against
            import torch  # noqa: F401
        except Exception:
            pass
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.pgo_last_error.restype = C.c_char_p
        _lib.pgo_graph_destroy.restype = None
        _lib.pgo_default_options.restype = None
        _lib.pgo_release_cached_memory.restype = None
        for name in EXPORTED_SYMBOLS:
            getattr(_lib, name)
    return _lib


def _check(rc: int):
    if rc != 0:
        raise PgoError(f"pgo_b200 error {rc}: {lib().pgo_last_error().decode()}")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def default_options() -> SolverOptions:
    o = SolverOptions()
    lib().pgo_default_options(C.byref(o))
    return o


def device_count() -> int:
    return int(lib().pgo_device_count())


def analyze_structure(n_poses, edge_ids, pose_const=None, max_fill_ratio: float = 0.0) -> StructureInfo:
    """Host-only structure analysis (pgo_analyze_structure); works without a GPU."""
    edge_ids = np.ascontiguousarray(edge_ids, np.int32)
    pc = None if pose_const is None else np.ascontiguousarray(pose_const, np.uint8)
    info = StructureInfo()
    _check(lib().pgo_analyze_structure(C.c_int(n_poses), C.c_int(edge_ids.shape[0]), edge_ids.ctypes.data_as(C.POINTER(C.c_int)),
                                       pc.ctypes.data_as(C.POINTER(C.c_ubyte)) if pc is not None else None,
                                       C.c_double(max_fill_ratio), C.byref(info)))
    return info


def analyze_partition(g, rank: int, world: int) -> PartitionInfo:
    """Host-only: the row partition, halo plan and multilevel hierarchy rank `rank` of `world` would build (no GPU)."""
    poses = np.ascontiguousarray(g.poses, np.float64)
    ids = np.ascontiguousarray(g.edge_ids, np.int32)
    pc = np.ascontiguousarray(g.pose_const, np.uint8)
    info = PartitionInfo()
    _check(lib().pgo_analyze_partition(C.c_int(poses.shape[0]), C.c_int(ids.shape[0]), _dp(poses), ids.ctypes.data_as(C.POINTER(C.c_int)),
                                       pc.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(rank), C.c_int(world), C.byref(info)))
    return info


def amg_aggregates(g, world: int = 1):
    """Host-only: [agg_0, agg_1, ...] -- for every level but the last, the aggregate (next-level node) of every node."""
    poses = np.ascontiguousarray(g.poses, np.float64)
    ids = np.ascontiguousarray(g.edge_ids, np.int32)
    pc = np.ascontiguousarray(g.pose_const, np.uint8)
    nl = C.c_int()
    nodes = (C.c_int * 16)()
    args = (C.c_int(poses.shape[0]), C.c_int(ids.shape[0]), _dp(poses), ids.ctypes.data_as(C.POINTER(C.c_int)),
            pc.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int(world), C.byref(nl), nodes)
    _check(lib().pgo_amg_aggregates(*args, None, C.c_longlong(0)))
    sizes = [nodes[l] for l in range(nl.value)]
    out = np.zeros(max(1, sum(sizes[:-1])), np.int32)
    _check(lib().pgo_amg_aggregates(*args, out.ctypes.data_as(C.POINTER(C.c_int)), C.c_longlong(out.size)))
    res, k = [], 0
    for n in sizes[:-1]:
        res.append(out[k:k + n].copy())
        k += n
    return sizes, res


def edge_candidates(positions, search_radius: float = 6.0, min_frame_gap: int = 100, device: int = 0):
    """Loop-edge candidates per frame (pgo_edge_candidates): returns (row_ptr[n+1], candidates) -- frame c's list is
    candidates[row_ptr[c]:row_ptr[c+1]] = [c-1, then every i < c - min_frame_gap within the search radius, ascending]."""
    pos = np.ascontiguousarray(positions, np.float64)
    n = int(pos.shape[0])
    row_ptr = np.zeros(n + 1, np.int64)
    total = C.c_longlong()
    _check(lib().pgo_edge_candidates(C.c_int(device), C.c_int(n), _dp(pos), C.c_double(search_radius), C.c_int(min_frame_gap),
                                     row_ptr.ctypes.data_as(C.POINTER(C.c_longlong)), None, C.c_longlong(0), C.byref(total)))
    idx = np.zeros(max(total.value, 1), np.int32)
    _check(lib().pgo_edge_candidates(C.c_int(device), C.c_int(n), _dp(pos), C.c_double(search_radius), C.c_int(min_frame_gap),
                                     row_ptr.ctypes.data_as(C.POINTER(C.c_longlong)), idx.ctypes.data_as(C.POINTER(C.c_int)),
                                     C.c_longlong(idx.size), C.byref(total)))
    return row_ptr, idx[:total.value]


def set_topology_cache(enabled: bool) -> bool:
    """Per-topology graph cache of solve_pose_graph (pgo_set_topology_cache); returns the previous setting."""
    return bool(lib().pgo_set_topology_cache(C.c_int(1 if enabled else 0)))


def release_cached_memory(device: int = -1):
    lib().pgo_release_cached_memory(C.c_int(device))


def nccl_unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    _check(lib().pgo_nccl_unique_id(buf))
    return bytes(buf)


class Graph:
    """A pose graph resident in HBM (pgo_graph)."""

    def __init__(self, poses, edge_ids, edge_meas, edge_sqrt_info=None, pose_const=None, device: int = 0,
                 unique_id: bytes | None = None, rank: int = 0, world: int = 1):
        """world > 1: every rank passes the same global graph and keeps its row slice (pgo_graph_create_partitioned)."""
        self._h = C.c_void_p()
        poses = np.ascontiguousarray(poses, np.float64)
        edge_ids = np.ascontiguousarray(edge_ids, np.int32)
        edge_meas = np.ascontiguousarray(edge_meas, np.float64)
        self.n_poses, self.n_edges = int(poses.shape[0]), int(edge_ids.shape[0])
        si = None if edge_sqrt_info is None else np.ascontiguousarray(edge_sqrt_info, np.float64)
        pc = None if pose_const is None else np.ascontiguousarray(pose_const, np.uint8)
        args = (C.byref(self._h), C.c_int(device), C.c_int(self.n_poses), C.c_int(self.n_edges),
                _dp(poses), edge_ids.ctypes.data_as(C.POINTER(C.c_int)), _dp(edge_meas), _dp(si),
                pc.ctypes.data_as(C.POINTER(C.c_ubyte)) if pc is not None else None)
        if world > 1:
            buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
            _check(lib().pgo_graph_create_partitioned(*args, buf, C.c_int(rank), C.c_int(world)))
        else:
            _check(lib().pgo_graph_create(*args))

    @classmethod
    def from_dataset(cls, g, device: int = 0, unique_id: bytes | None = None, rank: int = 0, world: int = 1):
        return cls(g.poses, g.edge_ids, g.edge_meas, g.edge_sqrt_info, g.pose_const, device, unique_id, rank, world)

    def local_sizes(self):
        """(owned poses, halo poses, local edges) of this rank's slice"""
        L = lib()
        return int(L.pgo_graph_num_local_poses(self._h)), int(L.pgo_graph_num_halo_poses(self._h)), int(L.pgo_graph_num_local_edges(self._h))

    def close(self):
        if self._h:
            lib().pgo_graph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        _check(lib().pgo_graph_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_poses(self, poses):
        _check(lib().pgo_graph_set_poses(self._h, _dp(np.ascontiguousarray(poses, np.float64))))

    def get_poses(self):
        out = np.empty((self.n_poses, 7))
        _check(lib().pgo_graph_get_poses(self._h, _dp(out)))
        return out

    def snapshot_poses(self):
        _check(lib().pgo_graph_snapshot_poses(self._h))

    def restore_poses(self):
        _check(lib().pgo_graph_restore_poses(self._h))

    def set_edge_losses(self, loss_types=None, loss_scales=None):
        """per-edge ceres::LossFunction (pgo_graph_set_edge_losses); None, None: back to one loss for every edge"""
        if loss_types is None:
            _check(lib().pgo_graph_set_edge_losses(self._h, None, None))
            return
        t = np.ascontiguousarray(loss_types, np.int32)
        s = np.ascontiguousarray(loss_scales, np.float64)
        assert t.size == self.n_edges and s.size == self.n_edges
        _check(lib().pgo_graph_set_edge_losses(self._h, t.ctypes.data_as(C.POINTER(C.c_int)), _dp(s)))

    def evaluate(self, loss_type=LOSS_HUBER, loss_a=1.0, want_jacobians=True):
        cost = C.c_double()
        res = np.zeros((self.n_edges, 6))
        grad = np.zeros((self.n_poses, 6))
        jac = np.zeros((self.n_edges, 2, 36)) if want_jacobians else None
        _check(lib().pgo_graph_evaluate(self._h, C.c_int(loss_type), C.c_double(loss_a), C.byref(cost), _dp(res),
                                        _dp(grad), _dp(jac)))
        return cost.value, res, grad, jac

    def linearize(self, loss_type=LOSS_HUBER, loss_a=1.0, scale=None):
        cost = C.c_double()
        ms = C.c_float()
        sc = None if scale is None else np.ascontiguousarray(scale, np.float64)
        _check(lib().pgo_graph_linearize(self._h, C.c_int(loss_type), C.c_double(loss_a), _dp(sc), C.byref(cost), C.byref(ms)))
        return cost.value, ms.value

    def hessian(self):
        """(row_ptr, col_idx, values[nnzb,6,6], gradient[N,6]) of the last linearisation."""
        nnzb = C.c_longlong()
        _check(lib().pgo_graph_get_hessian(self._h, C.byref(nnzb), None, None, None, None))
        rp = np.zeros(self.n_poses + 1, np.int32)
        ci = np.zeros(nnzb.value, np.int32)
        vals = np.zeros((nnzb.value, 36))
        grad = np.zeros((self.n_poses, 6))
        _check(lib().pgo_graph_get_hessian(self._h, C.byref(nnzb), rp.ctypes.data_as(C.POINTER(C.c_int)),
                                           ci.ctypes.data_as(C.POINTER(C.c_int)), _dp(vals), _dp(grad)))
        return rp, ci, vals.reshape(-1, 6, 6), grad

    def hessian_blocks(self) -> int:
        nnzb = C.c_longlong()
        _check(lib().pgo_graph_get_hessian(self._h, C.byref(nnzb), None, None, None, None))
        return int(nnzb.value)

    def spmv(self, x, d=None, repeats: int = 1):
        x = np.ascontiguousarray(x, np.float64)
        dd = None if d is None else np.ascontiguousarray(d, np.float64)
        y = np.zeros((self.n_poses, 6))
        ms = C.c_float()
        _check(lib().pgo_graph_spmv(self._h, _dp(x), _dp(dd), _dp(y), C.c_int(repeats), C.byref(ms)))
        return y, ms.value

    def linear_solve(self, d, b, options: SolverOptions | None = None):
        o = options or default_options()
        y = np.zeros((self.n_poses, 6))
        it = C.c_int()
        rel = C.c_double()
        ms = C.c_float()
        _check(lib().pgo_graph_linear_solve(self._h, C.byref(o), _dp(np.ascontiguousarray(d, np.float64)),
                                            _dp(np.ascontiguousarray(b, np.float64)), _dp(y), C.byref(it), C.byref(rel),
                                            C.byref(ms)))
        return y, it.value, rel.value, ms.value

    def solve(self, options: SolverOptions | None = None, max_log: int = 2048):
        o = options or default_options()
        s = SolverSummary()
        log = (IterationSummary * max_log)()
        _check(lib().pgo_graph_solve(self._h, C.byref(o), C.byref(s), log, C.c_int(max_log)))
        return s, [log[i] for i in range(min(s.num_iterations, max_log))]


def solve_pose_graph(poses, edge_ids, edge_meas, edge_sqrt_info=None, pose_const=None,
                     options: SolverOptions | None = None, device: int = 0, max_log: int = 2048,
                     edge_loss_types=None, edge_loss_scales=None):
    """ceres::Solve for the reference's pose graph, host buffers in and out (pgo_solve_pose_graph).
    edge_loss_types / edge_loss_scales: per-edge loss functions ([n_edges]); default: options.loss_type / loss_a."""
    o = options or default_options()
    keep = None
    if edge_loss_types is not None:
        keep = (np.ascontiguousarray(edge_loss_types, np.int32), np.ascontiguousarray(edge_loss_scales, np.float64))
        o.edge_loss_type = keep[0].ctypes.data_as(C.POINTER(C.c_int))
        o.edge_loss_a = keep[1].ctypes.data_as(C.POINTER(C.c_double))
    else:
        o.edge_loss_type = None
        o.edge_loss_a = None
    poses = np.ascontiguousarray(poses, np.float64).copy()
    edge_ids = np.ascontiguousarray(edge_ids, np.int32)
    edge_meas = np.ascontiguousarray(edge_meas, np.float64)
    si = None if edge_sqrt_info is None else np.ascontiguousarray(edge_sqrt_info, np.float64)
    pc = None if pose_const is None else np.ascontiguousarray(pose_const, np.uint8)
    s = SolverSummary()
    log = (IterationSummary * max_log)()
    _check(lib().pgo_solve_pose_graph(C.c_int(device), C.c_int(poses.shape[0]), _dp(poses), C.c_int(edge_ids.shape[0]),
                                      edge_ids.ctypes.data_as(C.POINTER(C.c_int)), _dp(edge_meas), _dp(si),
                                      pc.ctypes.data_as(C.POINTER(C.c_ubyte)) if pc is not None else None,
                                      C.byref(o), C.byref(s), log, C.c_int(max_log)))
    return poses, s, [log[i] for i in range(min(s.num_iterations, max_log))]
