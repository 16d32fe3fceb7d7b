// Stand-in for the reference's include/config.h (cv::FileStorage-backed YAML reader): the two integers
// generate_edges_from_trajectory_origion.cpp asks for come from the environment instead of ../config/config.yaml
// (REF_SEQUENCE_LENGTH, REF_SEARCH_RADIUS; the reference's yaml has sequence_length: 4541, search_radius: 6).
#ifndef REF_SHIM_CONFIG_H
#define REF_SHIM_CONFIG_H
#include <cstdlib>
#include <string>
namespace POSE_GRAPH {
class Config {
 public:
  static void setParameterFile(const std::string&) {}
  template <typename T> static T get(const std::string& key) {
    const char* v = key == "sequence_length" ? std::getenv("REF_SEQUENCE_LENGTH") : key == "search_radius" ? std::getenv("REF_SEARCH_RADIUS") : 0;
    return T(v ? std::atof(v) : 0.0);
  }
};
}  // namespace POSE_GRAPH
#endif
