// pose_graph_b200 -- the optimisation stage of the reference's pose_graph_ceres_plus_finial program
// (REF = /root/reference/src/POSE_GRAPH_CERES_PLUS/test/pose_graph_ceres_plus_finial.cpp), written against
// include/ceres_b200/ceres.h instead of <ceres/ceres.h>:
//
//   BuildOptimizationProblem()  REF:491-528   HuberLoss(1.0), EigenQuaternionParameterization on every q block,
//                                             first pose (poses->begin()) held constant
//   SolveOptimizationProblem()  REF:531-544   max_num_iterations = 1000, SPARSE_NORMAL_CHOLESKY, FullReport()
//   OutputPoses()               REF:547-567   "id x y z q_x q_y q_z q_w" per line
//
// The reference produces its vertices and edges from KITTI images (ORB matching + PnP, REF:74-130); that
// front end is out of scope, so the graph is read from a g2o file (VERTEX_SE3:QUAT / EDGE_SE3:QUAT) instead.
//
//   usage: pose_graph_b200 input.g2o [poses_before.txt] [poses_after.txt] [--no-loss] [--progress]
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "ceres_b200/ceres.h"

namespace ceres = ceres_b200;

namespace POSE_GRAPH {
typedef ceres::pgo::PosePod Pose3d;                   // REF/include/types.h:15-20 (Eigen-free: p[3], q[4] = x y z w)
typedef std::map<int, Pose3d> MapOfPoses;             // REF/include/types.h:22-24
struct Matrix6d {                                     // stand-in for Eigen::Matrix<double, 6, 6>
  double m[6][6];
  double operator()(int i, int j) const { return m[i][j]; }
  double& operator()(int i, int j) { return m[i][j]; }
  static Matrix6d Identity() { Matrix6d I; for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) I.m[i][j] = i == j; return I; }
  // information.llt().matrixL()   (REF:508)
  bool lltMatrixL(Matrix6d* L) const {
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) L->m[i][j] = 0.0;
    for (int j = 0; j < 6; ++j) {
      double s = m[j][j];
      for (int k = 0; k < j; ++k) s -= L->m[j][k] * L->m[j][k];
      if (!(s > 0.0)) return false;
      L->m[j][j] = std::sqrt(s);
      for (int i = j + 1; i < 6; ++i) {
        double t = m[i][j];
        for (int k = 0; k < j; ++k) t -= L->m[i][k] * L->m[j][k];
        L->m[i][j] = t / L->m[j][j];
      }
    }
    return true;
  }
};
struct Edge3d {                                       // REF/include/types.h:28-45
  int id_begin, id_end;
  Pose3d t_be;
  Matrix6d information;
};
typedef std::vector<Edge3d> VectorOfEdges;
}  // namespace POSE_GRAPH
using namespace POSE_GRAPH;

static bool ReadG2o(const std::string& filename, MapOfPoses* poses, VectorOfEdges* edges) {
  std::ifstream in(filename.c_str());
  if (!in) return false;
  std::string line, tag;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    if (!(ss >> tag)) continue;
    if (tag == "VERTEX_SE3:QUAT") {
      int id; Pose3d p;
      ss >> id >> p.p[0] >> p.p[1] >> p.p[2] >> p.q[0] >> p.q[1] >> p.q[2] >> p.q[3];
      if (!ss) return false;
      (*poses)[id] = p;
    } else if (tag == "EDGE_SE3:QUAT") {
      Edge3d e;
      ss >> e.id_begin >> e.id_end >> e.t_be.p[0] >> e.t_be.p[1] >> e.t_be.p[2] >> e.t_be.q[0] >> e.t_be.q[1] >> e.t_be.q[2] >> e.t_be.q[3];
      for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) { ss >> e.information(i, j); e.information(j, i) = e.information(i, j); }
      if (!ss) return false;
      edges->push_back(e);
    }
  }
  return true;
}

// REF:491-528
static void BuildOptimizationProblem(const VectorOfEdges& Edges, MapOfPoses* poses, ceres::Problem* problem, bool use_loss) {
  ceres::LossFunction* loss_function = use_loss ? new ceres::HuberLoss(1.0) : NULL;
  ceres::LocalParameterization* quaternion_local_parameterization = new ceres::EigenQuaternionParameterization;
  for (VectorOfEdges::const_iterator it = Edges.begin(); it != Edges.end(); ++it) {
    const Edge3d& edge = *it;
    MapOfPoses::iterator pose_begin_iter = poses->find(edge.id_begin);
    MapOfPoses::iterator pose_end_iter = poses->find(edge.id_end);
    if (pose_begin_iter == poses->end() || pose_end_iter == poses->end()) {
      std::cerr << "edge " << edge.id_begin << " -> " << edge.id_end << " references a missing vertex\n";
      std::exit(2);
    }
    Matrix6d sqrt_information;
    if (!edge.information.lltMatrixL(&sqrt_information)) { std::cerr << "information matrix is not positive definite\n"; std::exit(2); }
    const double t_be[7] = {edge.t_be.p[0], edge.t_be.p[1], edge.t_be.p[2], edge.t_be.q[0], edge.t_be.q[1], edge.t_be.q[2], edge.t_be.q[3]};
    ceres::CostFunction* cost_function = ceres::pgo::MakePoseGraph3dCost(t_be, &sqrt_information.m[0][0]);
    problem->AddResidualBlock(cost_function, loss_function, pose_begin_iter->second.p, pose_begin_iter->second.q,
                              pose_end_iter->second.p, pose_end_iter->second.q);
    problem->SetParameterization(pose_begin_iter->second.q, quaternion_local_parameterization);
    problem->SetParameterization(pose_end_iter->second.q, quaternion_local_parameterization);
  }
  MapOfPoses::iterator pose_start_iter = poses->begin();
  problem->SetParameterBlockConstant(pose_start_iter->second.p);
  problem->SetParameterBlockConstant(pose_start_iter->second.q);
}

// REF:531-544
static bool SolveOptimizationProblem(ceres::Problem* problem, bool progress) {
  ceres::Solver::Options options;
  options.max_num_iterations = 1000;
  options.linear_solver_type = ceres::SPARSE_NORMAL_CHOLESKY;
  options.minimizer_progress_to_stdout = progress;
  ceres::Solver::Summary summary;
  ceres::Solve(options, problem, &summary);
  std::cout << summary.FullReport() << '\n';
  return summary.IsSolutionUsable();
}

// REF:547-567
static bool OutputPoses(const std::string& filename, const MapOfPoses& poses) {
  std::ofstream outfile(filename.c_str());
  if (!outfile) { std::cout << "Error opening the file: " << filename; return false; }
  outfile.precision(17);
  for (MapOfPoses::const_iterator it = poses.begin(); it != poses.end(); ++it)
    outfile << it->first << " " << it->second.p[0] << " " << it->second.p[1] << " " << it->second.p[2] << " " << it->second.q[0]
            << " " << it->second.q[1] << " " << it->second.q[2] << " " << it->second.q[3] << '\n';
  return true;
}

int main(int argc, char** argv) {
  std::vector<std::string> pos;
  bool use_loss = true, progress = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--no-loss") use_loss = false;
    else if (a == "--progress") progress = true;
    else pos.push_back(a);
  }
  if (pos.empty()) { std::cerr << "usage: pose_graph_b200 input.g2o [before.txt] [after.txt] [--no-loss] [--progress]\n"; return 2; }
  MapOfPoses poses;
  VectorOfEdges Edges;
  if (!ReadG2o(pos[0], &poses, &Edges)) { std::cerr << "cannot read " << pos[0] << "\n"; return 2; }
  std::cout << "Number of poses: " << poses.size() << "\nNumber of edges: " << Edges.size() << '\n';
  if (pos.size() > 1) OutputPoses(pos[1], poses);
  ceres::Problem problem;
  try {
    BuildOptimizationProblem(Edges, &poses, &problem, use_loss);
    const bool ok = SolveOptimizationProblem(&problem, progress);
    std::cout << (ok ? "Optimizing Suscessfully!" : "May be some problems!") << std::endl;   // REF:136-139
    if (pos.size() > 2) OutputPoses(pos[2], poses);
    return ok ? 0 : 1;
  } catch (const std::exception& ex) {
    std::cerr << "error: " << ex.what() << "\n";
    return 3;
  }
}
