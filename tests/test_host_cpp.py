"""The C++ host-side mirror of the Ceres surface (include/ceres_b200/ceres.h) and the example that
restates the reference's BuildOptimizationProblem / SolveOptimizationProblem / OutputPoses on it."""
import os
import subprocess

import numpy as np
import pytest

from helpers import rot_angle_between

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "examples", "pose_graph_b200")
CANDIDATES_EXAMPLE = os.path.join(ROOT, "examples", "edge_candidates_b200")


@pytest.fixture(scope="module")
def example(pgo):
    pgo.lib()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s"])
    return EXAMPLE


@pytest.fixture(scope="module")
def D():
    import posegraph_ceres_b200.datasets as d
    return d


def test_g2o_round_trip(D, tmp_path):
    g = D.sphere(6, 10, None)
    path = str(tmp_path / "s.g2o")
    D.write_g2o(g, path)
    h = D.read_g2o(path)
    assert np.array_equal(h.poses, g.poses) and np.array_equal(h.edge_ids, g.edge_ids)
    assert np.array_equal(h.edge_meas, g.edge_meas) and h.pose_const[0] == 1 and h.pose_const.sum() == 1
    assert np.abs(h.edge_sqrt_info - g.edge_sqrt_info).max() <= 1e-12 * np.abs(g.edge_sqrt_info).max()


def test_example_fails_loudly_without_gpu(pgo, D, example, tmp_path):
    """no CPU fallback behind the Ceres-surface mirror either"""
    if pgo.device_count() > 0:
        pytest.skip("GPU present")
    path = str(tmp_path / "m.g2o")
    D.write_g2o(D.manhattan_loop(), path)
    r = subprocess.run([example, path], capture_output=True, text=True)
    assert r.returncode == 1
    assert "FAILURE" in r.stdout and "no CUDA device" in r.stdout and "May be some problems!" in r.stdout


def test_mirror_rejects_what_the_device_path_cannot_do(tmp_path, pgo):
    """error behaviour of ceres_b200::Problem: compile a tiny program that violates each contract."""
    pgo.lib()
    src = tmp_path / "contracts.cpp"
    src.write_text(r'''
#include <cstdio>
#include "ceres_b200/ceres.h"
namespace ceres = ceres_b200;
template <typename F> static int throws(F f) { try { f(); } catch (const std::invalid_argument&) { return 1; } return 0; }
int main() {
  double t[7] = {1, 0, 0, 0, 0, 0, 1};
  double p0[3] = {0, 0, 0}, q0[4] = {0, 0, 0, 1}, p1[3] = {1, 0, 0}, q1[4] = {0, 0, 0, 1}, q2[4] = {0, 0, 0, 1};
  int ok = 1;
  { ceres::Problem pr;   // unknown block
    ok &= throws([&] { pr.SetParameterBlockConstant(p0); });
    ok &= throws([&] { pr.SetParameterization(q0, new ceres::EigenQuaternionParameterization); }); }
  { ceres::Problem pr;   // p paired with two different q blocks
    pr.AddResidualBlock(ceres::PoseGraph3dErrorTerm::Create(t, nullptr), nullptr, p0, q0, p1, q1);
    ok &= throws([&] { pr.AddResidualBlock(ceres::PoseGraph3dErrorTerm::Create(t, nullptr), nullptr, p0, q2, p1, q1); });
    ok &= throws([&] { pr.AddResidualBlock(ceres::PoseGraph3dErrorTerm::Create(t, nullptr), nullptr, p0, q0, p0, q0); });
    ok &= (pr.NumResidualBlocks() == 1 && pr.NumParameterBlocks() == 5 && pr.NumResiduals() == 6); }
  { ceres::Problem pr;   // q without EigenQuaternionParameterization / half-constant pose / mixed losses -> Solve rejects
    pr.AddResidualBlock(ceres::PoseGraph3dErrorTerm::Create(t, nullptr), new ceres::HuberLoss(1.0), p0, q0, p1, q1);
    ceres::Solver::Options o; ceres::Solver::Summary s;
    ok &= throws([&] { ceres::Solve(o, &pr, &s); });
    ceres::LocalParameterization* lp = new ceres::EigenQuaternionParameterization;
    pr.SetParameterization(q0, lp); pr.SetParameterization(q1, lp);
    pr.SetParameterBlockConstant(p0);
    ok &= throws([&] { ceres::Solve(o, &pr, &s); });
    pr.SetParameterBlockConstant(q0);
    pr.AddResidualBlock(ceres::PoseGraph3dErrorTerm::Create(t, nullptr), new ceres::CauchyLoss(1.0), p1, q1, p0, q0);
    ok &= throws([&] { ceres::Solve(o, &pr, &s); });
    ok &= pr.IsParameterBlockConstant(p0) && !pr.IsParameterBlockConstant(p1); }
  { double rho[3]; ceres::HuberLoss h(1.0); h.Evaluate(4.0, rho); ok &= (rho[0] == 3.0 && rho[1] == 0.5 && rho[2] == -0.0625); }
  std::printf("%s\n", ok ? "CONTRACTS_OK" : "CONTRACTS_BROKEN");
  return ok ? 0 : 1;
}
''')
    exe = str(tmp_path / "contracts")
    libdir = os.path.join(ROOT, "posegraph-ceres_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src),
                           "-L", libdir, "-lpgo_b200", f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"])
    env = dict(os.environ)
    import torch  # noqa: F401  (its bundled libnccl / libcudart directories are on the loader path below)
    extra = []
    for mod in ("nvidia.nccl", "nvidia.cuda_runtime"):
        try:
            m = __import__(mod, fromlist=["x"])
            extra.append(os.path.join(list(m.__path__)[0], "lib"))
        except Exception:
            pass
    env["LD_LIBRARY_PATH"] = ":".join(extra + [env.get("LD_LIBRARY_PATH", "")])
    r = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "CONTRACTS_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["manhattan", "sphere"])
def test_example_matches_oracle(pgo, oracle, D, example, tmp_path, name):
    """g2o file -> C++ Problem/Solve mirror -> OutputPoses; converged poses vs the CPU oracle (1e-4 m / 1e-4 rad)."""
    g = D.manhattan_loop() if name == "manhattan" else D.sphere(10, 20, None)
    path, out = str(tmp_path / "g.g2o"), str(tmp_path / "after.txt")
    D.write_g2o(g, path)
    r = subprocess.run([example, path, str(tmp_path / "before.txt"), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Optimizing Suscessfully!" in r.stdout and "Termination:                     CONVERGENCE" in r.stdout
    got = np.loadtxt(out)
    assert np.array_equal(got[:, 0], np.arange(g.n_poses))
    ref, rs, _ = oracle.solve(D.read_g2o(path))
    assert np.abs(got[:, 1:4] - ref[:, :3]).max() <= 1e-4
    assert rot_angle_between(got[:, 4:8], ref[:, 3:]).max() <= 1e-4
    before = np.loadtxt(str(tmp_path / "before.txt"))
    assert np.abs(before[:, 1:] - g.poses).max() <= 1e-15


@pytest.mark.gpu
def test_example_reproduces_the_reference_trajectory(D, example, tmp_path):
    """The reference-facing C++ surface end to end on the reference's own problem: KITTI-00 as a g2o file ->
    ceres_b200::Problem / Solve (the patched BuildOptimizationProblem / SolveOptimizationProblem) -> OutputPoses;
    the written trajectory is within 2.5 cm / 1e-3 rad of the reference's result/trajectory/trajectory_update_y_not_constant.txt."""
    g = D.kitti00()
    path, out = str(tmp_path / "kitti00.g2o"), str(tmp_path / "trajectory_update.txt")
    D.write_g2o(g, path)
    r = subprocess.run([example, path, str(tmp_path / "trajectory_origin.txt"), out], capture_output=True, text=True)
    assert r.returncode == 0 and "Optimizing Suscessfully!" in r.stdout, r.stdout + r.stderr
    got = np.loadtxt(out)
    err = np.linalg.norm(got[:, 1:4] - g.truth[:, :3], axis=1)
    assert err.max() <= 0.025 and err.mean() <= 0.010
    assert rot_angle_between(got[:, 4:8], g.truth[:, 3:]).max() <= 1e-3


def _write_trajectory(path, with_ids=False):
    f = np.load(os.path.join(ROOT, "tests", "golden", "kitti00_fixture.npz"))
    poses = f["poses_before"]
    if with_ids:
        poses = np.column_stack([np.arange(len(poses)), poses])
    np.savetxt(path, poses, fmt="%.17g")


def test_candidates_example_fails_loudly_without_gpu(pgo, example, tmp_path):
    if pgo.device_count() > 0:
        pytest.skip("GPU present")
    _write_trajectory(str(tmp_path / "trajectory.txt"))
    r = subprocess.run([CANDIDATES_EXAMPLE, str(tmp_path / "trajectory.txt"), str(tmp_path / "out.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    r = subprocess.run([CANDIDATES_EXAMPLE, str(tmp_path / "missing.txt"), str(tmp_path / "out.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("with_ids", [False, True])
def test_candidates_example_writes_the_reference_file(example, tmp_path, with_ids):
    """examples/edge_candidates_b200 (the reference's generate_edges_from_trajectory_origion with the CUDA search) on the
    reference's trajectory_origin poses: the written Edge_Candidates_index.txt is byte-identical to the reference's."""
    import hashlib
    from test_oracle_cpu import REF_CANDIDATE_FILE_SHA256
    traj, out = str(tmp_path / "trajectory.txt"), str(tmp_path / "Edge_Candidates_index.txt")
    _write_trajectory(traj, with_ids)
    r = subprocess.run([CANDIDATES_EXAMPLE, traj, out], capture_output=True, text=True)
    assert r.returncode == 0 and "4541 frames, 20499 candidates" in r.stdout, r.stdout + r.stderr
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == REF_CANDIDATE_FILE_SHA256
