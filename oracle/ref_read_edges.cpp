// ref_read_edges.cpp -- TEST INFRASTRUCTURE (oracle/_ref): runs the REFERENCE's own candidate-file reader,
//   POSE_GRAPH::getEdegsCandidateIndex()   /root/reference/src/POSE_GRAPH_CERES_PLUS/include/ReadEdges.h:9-48
// compiled unmodified from where it lies, and prints the map it returns ("key: v v ...").  The reader opens
// "../config/Edge_Candidates_index.txt" relative to the working directory, as pose_graph_ceres_plus_finial.cpp:72 does.
// Its umbrella header common_include.h pulls in OpenCV and the ORB front end, none of which the reader uses: the include
// guard is pre-defined so that only the standard headers below are needed.
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
using namespace std;   // common_include.h:62
#define COMMON_INCLUDE_H
#include "ReadEdges.h"

int main() {
  const map<int, vector<int> > m = POSE_GRAPH::getEdegsCandidateIndex();
  for (map<int, vector<int> >::const_iterator it = m.begin(); it != m.end(); ++it) {
    cout << it->first << ":";
    for (size_t k = 0; k < it->second.size(); ++k) cout << " " << it->second[k];
    cout << "\n";
  }
  return 0;
}
