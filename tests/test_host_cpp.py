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
    pr.AddResidualBlock(ceres::pgo::MakePoseGraph3dCost(t, nullptr), nullptr, p0, q0, p1, q1);
    ok &= throws([&] { pr.AddResidualBlock(ceres::pgo::MakePoseGraph3dCost(t, nullptr), nullptr, p0, q2, p1, q1); });
    ok &= throws([&] { pr.AddResidualBlock(ceres::pgo::MakePoseGraph3dCost(t, nullptr), nullptr, p0, q0, p0, q0); });
    // a rejected call registers nothing (q2 is not a parameter block) and the problem owns the rejected cost functions
    ok &= (pr.NumResidualBlocks() == 1 && pr.NumParameterBlocks() == 4 && pr.NumResiduals() == 6 && !pr.HasParameterBlock(q2)); }
  { ceres::Problem pr;   // q without EigenQuaternionParameterization -> Solve rejects
    pr.AddResidualBlock(ceres::pgo::MakePoseGraph3dCost(t, nullptr), new ceres::HuberLoss(1.0), p0, q0, p1, q1);
    ceres::Solver::Options o; ceres::Solver::Summary s;
    ok &= throws([&] { ceres::Solve(o, &pr, &s); });
    ceres::LocalParameterization* lp = new ceres::EigenQuaternionParameterization;
    pr.SetParameterization(q0, lp); pr.SetParameterization(q1, lp);
    // what Ceres allows per block is accepted: p constant without q, a different loss on another residual block
    pr.SetParameterBlockConstant(p0);
    pr.AddResidualBlock(ceres::pgo::MakePoseGraph3dCost(t, nullptr), new ceres::CauchyLoss(1.0), p1, q1, p0, q0);
    ok &= !throws([&] { ceres::Solve(o, &pr, &s); });
    ok &= pr.IsParameterBlockConstant(p0) && !pr.IsParameterBlockConstant(q0) && !pr.IsParameterBlockConstant(p1); }
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


def test_mirror_compiles_in_a_reference_shaped_translation_unit_and_recovers_the_functor(tmp_path, pgo):
    """The reference TU has `using namespace ceres; using namespace POSE_GRAPH;` (REF test/pose_graph_ceres_plus_finial.cpp:20-22)
    and its own POSE_GRAPH::Pose3d / PoseGraph3dErrorTerm: the mirror must not export colliding names, and
    ceres::AutoDiffCostFunction<Functor, 6, 3, 4, 3, 4> must recover (t_ab, sqrt_information) from ANY functor that is the
    SE(3) relative-pose residual -- here a restatement on plain arrays with random full sqrt-information -- and reject
    everything else.  Host-only: no GPU needed."""
    src = tmp_path / "ref_shaped.cpp"
    src.write_text(r"""
#include <cmath>
#include <cstdio>
#include <random>
#include <ceres/ceres.h>          // resolves to include/ceres_b200/compat/ceres/ceres.h
namespace POSE_GRAPH {
struct Pose3d { double p[3]; double q[4]; };
// the residual of PoseGraph3dError.h:21-54 on plain arrays (templated like the reference's functor)
class PoseGraph3dErrorTerm {
 public:
  PoseGraph3dErrorTerm(const Pose3d& t, const double* S, double bias = 0.0) : t_(t), bias_(bias) { for (int k = 0; k < 36; ++k) S_[k] = S[k]; }
  template <typename T> static void qmul(const T* a, const T* b, T* o) {
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  }
  template <typename T>
  bool operator()(const T* const pa, const T* const qa, const T* const pb, const T* const qb, T* r) const {
    const T qi[4] = {-qa[0], -qa[1], -qa[2], qa[3]};
    const T d[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    const T tx = T(2.0) * (qi[1] * d[2] - qi[2] * d[1]), ty = T(2.0) * (qi[2] * d[0] - qi[0] * d[2]), tz = T(2.0) * (qi[0] * d[1] - qi[1] * d[0]);
    T e[6];
    e[0] = d[0] + qi[3] * tx + (qi[1] * tz - qi[2] * ty) - T(t_.p[0]);
    e[1] = d[1] + qi[3] * ty + (qi[2] * tx - qi[0] * tz) - T(t_.p[1]);
    e[2] = d[2] + qi[3] * tz + (qi[0] * ty - qi[1] * tx) - T(t_.p[2]);
    T qab[4], dq[4];
    qmul(qi, qb, qab);
    const T qabc[4] = {-qab[0], -qab[1], -qab[2], qab[3]};
    const T qm[4] = {T(t_.q[0]), T(t_.q[1]), T(t_.q[2]), T(t_.q[3])};
    qmul(qm, qabc, dq);
    e[3] = T(2.0) * dq[0]; e[4] = T(2.0) * dq[1]; e[5] = T(2.0) * dq[2];
    for (int i = 0; i < 6; ++i) { T s = T(bias_); for (int k = 0; k < 6; ++k) s = s + T(S_[i * 6 + k]) * e[k]; r[i] = s; }
    return true;
  }
  static ceres::CostFunction* Create(const Pose3d& t, const double* S) {
    return new ceres::AutoDiffCostFunction<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4>(new PoseGraph3dErrorTerm(t, S));
  }
 private:
  Pose3d t_; double S_[36]; double bias_;
};
}  // namespace POSE_GRAPH
using namespace std;
using namespace ceres;
using namespace POSE_GRAPH;       // Pose3d / PoseGraph3dErrorTerm must stay unambiguous
int main() {
  std::mt19937_64 rng(7);
  std::normal_distribution<double> nrm;
  double worst = 0.0;
  for (int trial = 0; trial < 200; ++trial) {
    Pose3d t;
    double S[36], n = 0.0;
    for (double& v : t.p) v = 3.0 * nrm(rng);
    for (double& v : t.q) { v = nrm(rng); n += v * v; }
    for (double& v : t.q) v /= std::sqrt(n);
    for (int k = 0; k < 36; ++k) S[k] = (k / 6 == k % 6 ? 5.0 : 0.0) + nrm(rng) * (trial % 2 ? 1.0 : 0.2);
    if (trial % 3 == 0) for (int i = 0; i < 6; ++i) for (int j = i + 1; j < 6; ++j) S[i * 6 + j] = 0.0;   // llt().matrixL()
    CostFunction* c = PoseGraph3dErrorTerm::Create(t, S);
    ceres::pgo::PoseGraph3dCost* pc = dynamic_cast<ceres::pgo::PoseGraph3dCost*>(c);
    if (!pc || pc->num_residuals() != 6 || pc->parameter_block_sizes().size() != 4) { std::printf("BAD_TYPE\n"); return 1; }
    PoseGraph3dErrorTerm f(t, S);
    for (int k = 0; k < 8; ++k) {
      double pa[3], qa[4], pb[3], qb[4], na = 0, nb = 0, want[6], got[6];
      for (double& v : pa) v = nrm(rng);
      for (double& v : pb) v = nrm(rng);
      for (double& v : qa) { v = nrm(rng); na += v * v; }
      for (double& v : qb) { v = nrm(rng); nb += v * v; }
      for (double& v : qa) v /= std::sqrt(na);
      for (double& v : qb) v /= std::sqrt(nb);
      f(pa, qa, pb, qb, want);
      ceres::pgo::pose_graph_residual(pc->t_ab(), pc->sqrt_information(), pa, qa, pb, qb, got);
      double sc = 1.0;
      for (int i = 0; i < 6; ++i) sc = std::max(sc, std::fabs(want[i]));
      for (int i = 0; i < 6; ++i) worst = std::max(worst, std::fabs(want[i] - got[i]) / sc);
    }
    delete c;
  }
  int rejected = 0;
  { Pose3d t = {{1, 2, 3}, {0, 0, 0, 1}}; double S[36]; for (int k = 0; k < 36; ++k) S[k] = k / 6 == k % 6;
    try { delete new ceres::AutoDiffCostFunction<PoseGraph3dErrorTerm, 6, 3, 4, 3, 4>(new PoseGraph3dErrorTerm(t, S, 0.25)); }
    catch (const std::invalid_argument&) { rejected = 1; } }
  std::printf("worst %.3e rejected %d\n", worst, rejected);
  std::printf("%s\n", worst <= 1e-12 && rejected ? "PROBE_OK" : "PROBE_BROKEN");
  return worst <= 1e-12 && rejected ? 0 : 1;
}
""")
    exe = str(tmp_path / "ref_shaped")
    libdir = os.path.join(ROOT, "posegraph-ceres_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include", "ceres_b200", "compat"), "-o", exe, str(src),
                           "-L", libdir, "-lpgo_b200", f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"])
    r = subprocess.run([exe], capture_output=True, text=True, env=_loader_env())
    assert r.returncode == 0 and "PROBE_OK" in r.stdout, r.stdout + r.stderr


def _loader_env():
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
    return env


def _write_drop_in_graph(g, path):
    """graph file of oracle/ref_drop_in.cpp: poses, then edges with the 6x6 information = S S^T (row-major)"""
    with open(path, "w") as f:
        f.write(f"{g.n_poses} {g.n_edges}\n")
        for i in range(g.n_poses):
            f.write(f"{i} " + " ".join(repr(float(v)) for v in g.poses[i]) + "\n")
        for e in range(g.n_edges):
            S = g.edge_sqrt_info[e].reshape(6, 6)
            info = S @ S.T
            f.write(f"{g.edge_ids[e, 0]} {g.edge_ids[e, 1]} " + " ".join(repr(float(v)) for v in g.edge_meas[e]) + " "
                    + " ".join(repr(float(v)) for v in info.ravel()) + "\n")


REF_DROP_IN = os.path.join(ROOT, "oracle", "_ref", "ref_drop_in")


def _need_drop_in(oracle):
    oracle.ref_functor()          # runs `make -C oracle ref` when /root/reference is present
    if not os.path.exists(REF_DROP_IN):
        pytest.skip("oracle/_ref/ref_drop_in not built (needs /root/reference at build time)")


def test_reference_text_drop_in_fails_loudly_without_gpu(pgo, oracle, D, tmp_path):
    """The reference's own BuildOptimizationProblem / SolveOptimizationProblem / OutputPoses text (REF :491-567, extracted
    at build time) compiled against the mirror: builds, runs its problem setup on the host, and -- without a GPU --
    reports FAILURE instead of falling back to anything."""
    _need_drop_in(oracle)
    if pgo.device_count() > 0:
        pytest.skip("GPU present")
    path = str(tmp_path / "g.txt")
    _write_drop_in_graph(D.manhattan_loop(), path)
    r = subprocess.run([REF_DROP_IN, path, str(tmp_path / "out.txt")], capture_output=True, text=True, env=_loader_env())
    assert r.returncode == 1, r.stdout + r.stderr
    assert "Number of poses: 100" in r.stdout and "no CUDA device" in r.stdout and "May be some problems!" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kitti00", "sphere"])
def test_reference_text_drop_in_matches_oracle(pgo, oracle, D, tmp_path, name):
    """The reference's own optimisation-stage text + its unmodified PoseGraph3dError.h, linked against libpgo_b200.so
    (oracle/ref_drop_in.cpp): the poses it writes with its own OutputPoses equal the CPU oracle's."""
    if not os.path.exists(REF_DROP_IN):
        pytest.skip("oracle/_ref/ref_drop_in not built (needs /root/reference at build time)")
    g = D.kitti00() if name == "kitti00" else D.sphere(10, 20, None)
    path, out = str(tmp_path / "g.txt"), str(tmp_path / "out.txt")
    _write_drop_in_graph(g, path)
    r = subprocess.run([REF_DROP_IN, path, out], capture_output=True, text=True, env=_loader_env())
    assert r.returncode == 0 and "Optimizing Suscessfully!" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    got = np.loadtxt(out)
    assert np.array_equal(got[:, 0], np.arange(g.n_poses))
    ref, rs, _ = oracle.solve(g)
    # OutputPoses prints 6 significant digits (the stream's default precision): that is the reference's own text
    assert np.abs(got[:, 1:4] - ref[:, :3]).max() <= 1e-4 + 1e-5 * np.abs(ref[:, :3]).max()
    from helpers import rot_angle_between
    assert rot_angle_between(got[:, 4:8], ref[:, 3:]).max() <= 1e-4 + 2e-5


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


@pytest.mark.gpu
def test_mirror_per_block_loss_and_half_constant_pose_match_oracle(tmp_path, pgo, oracle, D):
    """ceres_b200::Problem with a different LossFunction per residual block and SetParameterBlockConstant on p alone / q
    alone (what Ceres allows per block): the C++ mirror on the GPU vs the oracle with the same per-block settings."""
    import dataclasses
    g = D.sphere(8, 12, None)
    types = (np.arange(g.n_edges) % 3).astype(np.int32)
    scales = np.where(types == 1, 0.5, 2.0)
    pc = g.pose_const.copy()
    pc[3], pc[5] = 2, 3
    g2 = dataclasses.replace(g, pose_const=pc)
    with oracle.edge_losses(types, scales):
        ref, rs, _ = oracle.solve(g2)
    gfile, out = str(tmp_path / "g.txt"), str(tmp_path / "out.txt")
    with open(gfile, "w") as f:
        f.write(f"{g.n_poses} {g.n_edges}\n")
        for i in range(g.n_poses):
            f.write(" ".join(repr(float(v)) for v in g.poses[i]) + f" {int(pc[i])}\n")
        for e in range(g.n_edges):
            f.write(f"{g.edge_ids[e, 0]} {g.edge_ids[e, 1]} {int(types[e])} {float(scales[e])!r} " + " ".join(repr(float(v)) for v in g.edge_meas[e]) + " "
                    + " ".join(repr(float(v)) for v in g.edge_sqrt_info[e]) + "\n")
    src = tmp_path / "perblock.cpp"
    src.write_text(r"""
#include <cstdio>
#include <fstream>
#include <vector>
#include <ceres/ceres.h>
int main(int argc, char** argv) {
  std::ifstream in(argv[1]);
  int n, m;
  in >> n >> m;
  std::vector<double> poses(7 * n), meas(7 * m), S(36 * m), la(m);
  std::vector<int> pc(n), ids(2 * m), lt(m);
  for (int i = 0; i < n; ++i) { for (int k = 0; k < 7; ++k) in >> poses[7 * i + k]; in >> pc[i]; }
  for (int e = 0; e < m; ++e) {
    in >> ids[2 * e] >> ids[2 * e + 1] >> lt[e] >> la[e];
    for (int k = 0; k < 7; ++k) in >> meas[7 * e + k];
    for (int k = 0; k < 36; ++k) in >> S[36 * e + k];
  }
  ceres::Problem problem;
  ceres::LocalParameterization* lp = new ceres::EigenQuaternionParameterization;
  for (int e = 0; e < m; ++e) {
    ceres::LossFunction* loss = lt[e] == 0 ? nullptr : lt[e] == 1 ? (ceres::LossFunction*)new ceres::HuberLoss(la[e]) : new ceres::CauchyLoss(la[e]);
    double* pa = &poses[7 * ids[2 * e]];
    double* pb = &poses[7 * ids[2 * e + 1]];
    problem.AddResidualBlock(ceres::pgo::MakePoseGraph3dCost(&meas[7 * e], &S[36 * e]), loss, pa, pa + 3, pb, pb + 3);
    problem.SetParameterization(pa + 3, lp);
    problem.SetParameterization(pb + 3, lp);
  }
  for (int i = 0; i < n; ++i) {
    if (pc[i] == 1 || pc[i] == 2) problem.SetParameterBlockConstant(&poses[7 * i]);
    if (pc[i] == 1 || pc[i] == 3) problem.SetParameterBlockConstant(&poses[7 * i + 3]);
  }
  ceres::Solver::Options options;
  options.max_num_iterations = 1000;
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  std::printf("%s\n", summary.BriefReport().c_str());
  FILE* f = std::fopen(argv[2], "w");
  for (int i = 0; i < n; ++i) { for (int k = 0; k < 7; ++k) std::fprintf(f, "%.17g ", poses[7 * i + k]); std::fprintf(f, "\n"); }
  std::fclose(f);
  return summary.IsSolutionUsable() ? 0 : 1;
}
""")
    exe = str(tmp_path / "perblock")
    libdir = os.path.join(ROOT, "posegraph-ceres_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include", "ceres_b200", "compat"), "-o", exe, str(src),
                           "-L", libdir, "-lpgo_b200", f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"])
    r = subprocess.run([exe, gfile, out], capture_output=True, text=True, env=_loader_env())
    assert r.returncode == 0 and "CONVERGENCE" in r.stdout, r.stdout + r.stderr
    got = np.loadtxt(out)
    assert np.abs(got[:, :3] - ref[:, :3]).max() <= 1e-4 and rot_angle_between(got[:, 3:], ref[:, 3:]).max() <= 1e-4
    assert np.array_equal(got[3, :3], g.poses[3, :3]) and np.array_equal(got[5, 3:], g.poses[5, 3:])
