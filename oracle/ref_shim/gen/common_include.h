// Stand-in for the reference's umbrella header include/common_include.h (TEST INFRASTRUCTURE): shadows it on the include
// path so that test/generate_edges_from_trajectory_origion.cpp compiles UNMODIFIED without OpenCV / Eigen / the ORB front
// end.  Provides only what that file uses: the std names, a minimal cv::Mat (float matrix views with rowRange / col /
// at<float>), and the colour macros (empty).
#ifndef REF_SHIM_COMMON_INCLUDE_H
#define REF_SHIM_COMMON_INCLUDE_H

#include <stdint.h>

#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace cv {
// A view (row / column range) of a shared row-major float matrix -- the subset of cv::Mat semantics the file relies on:
// copies share storage, rowRange(a, b) / col(c) return sub-views, at<float>(i) indexes a column (or row) vector,
// at<float>(i, j) a matrix element.
class Mat {
 public:
  Mat() : r0_(0), r1_(0), c0_(0), c1_(0), stride_(0) {}
  Mat(int rows, int cols) : d_(new std::vector<float>((size_t)rows * cols, 0.0f)), r0_(0), r1_(rows), c0_(0), c1_(cols), stride_(cols) {}
  Mat rowRange(int a, int b) const { Mat m(*this); m.r0_ = r0_ + a; m.r1_ = r0_ + b; return m; }
  Mat col(int c) const { Mat m(*this); m.c0_ = c0_ + c; m.c1_ = c0_ + c + 1; return m; }
  int rows() const { return r1_ - r0_; }
  int cols() const { return c1_ - c0_; }
  template <typename T> T& at(int i, int j) { return (*d_)[(size_t)(r0_ + i) * stride_ + c0_ + j]; }
  template <typename T> T& at(int i) { return cols() == 1 ? at<T>(i, 0) : at<T>(0, i); }
 private:
  std::shared_ptr<std::vector<float> > d_;
  int r0_, r1_, c0_, c1_, stride_;
};
}  // namespace cv

using cv::Mat;
using namespace std;

#define RESET ""
#define BOLDGREEN ""
#define BOLDCYAN ""
#define BOLDRED ""
#define BOLDYELLOW ""
#define BOLDBLUE ""
#define BOLDMAGENTA ""
#define BOLDWHITE ""

#endif
