// oracle/shim/visualization_msgs/Marker.h — the fields of visualization_msgs::Marker that SDFMap's RViz exporters
// write (sdf_map.cpp:122-153, :370-420).  The exporters are out of scope; the type only has to exist
// for src/sdf_map.cpp to compile unmodified.  TEST INFRASTRUCTURE ONLY (see oracle/shim/Eigen/Eigen).
#ifndef GTOP_ORACLE_SHIM_MARKER
#define GTOP_ORACLE_SHIM_MARKER
#include <string>
#include <vector>

namespace geometry_msgs {
struct Point {
  double x = 0, y = 0, z = 0;
};
}  // namespace geometry_msgs

namespace visualization_msgs {
struct Marker {
  enum { CUBE_LIST = 6 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2 };
  struct {
    std::string frame_id;
  } header;
  int id = 0, type = 0, action = 0;
  struct {
    double x = 0, y = 0, z = 0;
  } scale;
  struct {
    float r = 0, g = 0, b = 0, a = 0;
  } color;
  std::vector<geometry_msgs::Point> points;
};
}  // namespace visualization_msgs
#endif
