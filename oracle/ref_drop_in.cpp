// ref_drop_in.cpp -- TEST INFRASTRUCTURE: the reference's OWN optimisation-stage text compiled against the B200 mirror.
//
// `make -C oracle ref` extracts BuildOptimizationProblem / SolveOptimizationProblem / OutputPoses verbatim from where they
// lie (REF/test/pose_graph_ceres_plus_finial.cpp:491-567) into _ref/ref_solve_text.inc (git-ignored, never copied into the
// repo) and compiles them here, together with the reference's unmodified PoseGraph3dError.h and types.h, with
//   -I include/ceres_b200/compat   (its <ceres/ceres.h> IS include/ceres_b200/ceres.h, `namespace ceres = ceres_b200`)
//   -I oracle/ref_shim             (Eigen stand-in: Eigen itself is not in this image)
// and links libpgo_b200.so: no source change in the reference -- the drop-in is an include path and a library.  The
// using-directives below are the reference's own (REF :20-22); they are what made an earlier version of the mirror
// ambiguous (it exported its own Pose3d / PoseGraph3dErrorTerm next to POSE_GRAPH's).
//
// usage: ref_drop_in graph.txt poses_out.txt
//   graph.txt: "n_poses n_edges", then per pose "id x y z qx qy qz qw", then per edge
//              "id_begin id_end x y z qx qy qz qw  i00 i01 ... i55" (information, 36 values row-major)
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "PoseGraph3dError.h"
#include "types.h"

/*ceres parts*/
#include <ceres/ceres.h>

#define BOLDCYAN ""     // REF/include/common_include.h terminal colour (the OpenCV-laden header is not included)

using namespace std;
using namespace ceres;
using namespace POSE_GRAPH;

void BuildOptimizationProblem(const VectorOfEdges& Edges, MapOfPoses* poses, ceres::Problem* problem);   // REF :39-42
bool SolveOptimizationProblem(ceres::Problem* problem);
bool OutputPoses(const std::string& filename, const MapOfPoses& poses);

#include "_ref/ref_solve_text.inc"

int main(int argc, char** argv) {
  if (argc < 3) { cerr << "usage: ref_drop_in graph.txt poses_out.txt\n"; return 2; }
  ifstream in(argv[1]);
  int n = 0, m = 0;
  if (!(in >> n >> m)) { cerr << "cannot read " << argv[1] << "\n"; return 2; }
  /*new add 2017.09.16 --> ceres*/           // REF :57-60
  Problem problem;
  MapOfPoses poses;
  VectorOfEdges Edges;
  for (int i = 0; i < n; ++i) {
    int id; double v[7];
    in >> id;
    for (double& x : v) in >> x;
    Pose3d p;
    p.p = Eigen::Vector3d(v[0], v[1], v[2]);
    p.q = Eigen::Quaterniond(v[6], v[3], v[4], v[5]);
    poses[id] = p;
  }
  for (int e = 0; e < m; ++e) {
    Edge3d edge;
    double v[7];
    in >> edge.id_begin >> edge.id_end;
    for (double& x : v) in >> x;
    edge.t_be.p = Eigen::Vector3d(v[0], v[1], v[2]);
    edge.t_be.q = Eigen::Quaterniond(v[6], v[3], v[4], v[5]);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) in >> edge.information(i, j);
    Edges.push_back(edge);
  }
  if (!in) { cerr << "truncated graph file\n"; return 2; }
  cout << "Number of poses: " << poses.size() << '\n';      // REF :128-130
  cout << "Number of edges: " << Edges.size() << '\n';
  try {
    BuildOptimizationProblem(Edges, &poses, &problem);      // REF :131
    if (SolveOptimizationProblem(&problem)) cout << "Optimizing Suscessfully!" << endl;   // REF :136-139
    else { cout << "May be some problems!" << endl; return 1; }
  } catch (const std::exception& ex) {
    cerr << "error: " << ex.what() << "\n";
    return 3;
  }
  cout.precision(17);
  return OutputPoses(argv[2], poses) ? 0 : 1;
}
