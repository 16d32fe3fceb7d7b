// Stand-in for the reference's include/GroundTruth.h + src/GroundTruth.cc:48-73 (loadPoses1) + Converter::toCvMat:
// reads "x y z q_x q_y q_z q_w" lines from $REF_TRAJECTORY into 4x4 CV_32F-like matrices (the translation column is
// the double rounded to float, which is all the candidate search looks at; the rotation block is left as identity).
#ifndef REF_SHIM_GROUNDTRUTH_H
#define REF_SHIM_GROUNDTRUTH_H
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "common_include.h"
namespace POSE_GRAPH {
class GroundTruth {
 public:
  GroundTruth() {
    const char* path = std::getenv("REF_TRAJECTORY");
    std::ifstream in(path ? path : "");
    std::string line;
    while (std::getline(in, line)) {
      std::istringstream ls(line);
      double v[7];
      int n = 0;
      while (n < 7 && (ls >> v[n])) ++n;
      if (n < 7) continue;
      cv::Mat P(4, 4);
      for (int i = 0; i < 4; ++i) P.at<float>(i, i) = 1.0f;
      for (int i = 0; i < 3; ++i) P.at<float>(i, 3) = (float)v[i];
      poses.push_back(P);
    }
  }
  std::vector<cv::Mat> getPoses() { return poses; }
 private:
  std::vector<cv::Mat> poses;
};
}  // namespace POSE_GRAPH
#endif
