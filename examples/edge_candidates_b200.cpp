// edge_candidates_b200 -- the reference's generate_edges_from_trajectory_origion program
// (REF = /root/reference/src/POSE_GRAPH_CERES_PLUS/test/generate_edges_from_trajectory_origion.cpp) with the
// getCandidatesIndex() / isInSearchRange() double loop (REF:58-110) replaced by one pgo_edge_candidates call:
//
//   main()  REF:25-55   read the trajectory, one candidate line per frame id = 1 .. n-1 into Edge_Candidates_index.txt,
//                       every token followed by a blank (REF:43-48)
//
// The trajectory file is the reference's format 1 (REF/config/config.yaml:9-12, GroundTruth::loadPoses1,
// REF/src/GroundTruth.cc:48-73): "x y z q_x q_y q_z q_w" per line; an 8-column file with a leading frame id
// (OutputPoses' format) is accepted too.  Only x y z are used (REF:10-14).
//
//   usage: edge_candidates_b200 trajectory.txt Edge_Candidates_index.txt [search_radius = 6] [min_frame_gap = 100]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "pgo_b200.h"

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: " << argv[0] << " trajectory.txt Edge_Candidates_index.txt [search_radius] [min_frame_gap]\n";
    return 2;
  }
  const double search_radius = argc > 3 ? std::atof(argv[3]) : 6.0;     // config.yaml: search_radius: 6
  const int min_frame_gap = argc > 4 ? std::atoi(argv[4]) : 100;        // REF:63-70
  std::ifstream in(argv[1]);
  if (!in) { std::cerr << "cannot open " << argv[1] << "\n"; return 1; }
  std::vector<double> centres;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::vector<double> v;
    double x;
    while (ls >> x) v.push_back(x);
    if (v.empty()) continue;
    if (v.size() != 7 && v.size() != 8) { std::cerr << argv[1] << ": expected 7 or 8 columns, found " << v.size() << "\n"; return 1; }
    const size_t o = v.size() - 7;
    centres.insert(centres.end(), v.begin() + o, v.begin() + o + 3);
  }
  const int n = (int)(centres.size() / 3);
  if (n == 0) { std::cerr << argv[1] << ": no poses\n"; return 1; }
  std::vector<long long> row_ptr((size_t)n + 1);
  long long total = 0;
  if (pgo_edge_candidates(0, n, centres.data(), search_radius, min_frame_gap, row_ptr.data(), nullptr, 0, &total) != PGO_OK) {
    std::cerr << "pgo_edge_candidates failed: " << pgo_last_error() << "\n";
    return 1;
  }
  std::vector<int> cand((size_t)(total > 0 ? total : 1));
  if (pgo_edge_candidates(0, n, centres.data(), search_radius, min_frame_gap, row_ptr.data(), cand.data(), total, &total) != PGO_OK) {
    std::cerr << "pgo_edge_candidates failed: " << pgo_last_error() << "\n";
    return 1;
  }
  std::ofstream outFile(argv[2]);
  if (!outFile) { std::cerr << "cannot open " << argv[2] << "\n"; return 1; }
  for (int id = 1; id < n; ++id) {
    outFile << id << " ";
    for (long long k = row_ptr[id]; k < row_ptr[id + 1]; ++k) outFile << cand[(size_t)k] << " ";
    outFile << std::endl;
  }
  std::cout << n << " frames, " << total << " candidates -> " << argv[2] << std::endl;
  return 0;
}
