// oracle/shim/ros/ros.h — stand-in for the three things the reference's hot-path sources take from roscpp:
// ros::param::get (grad_traj_optimizer.cpp:5-32), ros::Time (:135,194,285,435) and the ROS_* log macros
// (qp_generator.cpp:152,353,404).  TEST INFRASTRUCTURE ONLY (see oracle/shim/Eigen/Eigen).
// Parameters come from a process-wide table that the wrapper (oracle/ref_capi.cpp) fills before it constructs a
// GradTrajOptimizer; as with a real parameter server, a name that is not set leaves the caller's variable untouched.
#ifndef GTOP_ORACLE_SHIM_ROS
#define GTOP_ORACLE_SHIM_ROS
#include <chrono>
#include <cstdint>
#include <map>
#include <string>

namespace ros {

namespace param {
inline std::map<std::string, double>& table() {
  static std::map<std::string, double> t;
  return t;
}
template <typename T>
inline bool get(const std::string& key, T& value) {
  auto it = table().find(key);
  if (it == table().end()) return false;
  value = static_cast<T>(it->second);
  return true;
}
}  // namespace param

class Duration {
 public:
  explicit Duration(int64_t ns = 0) : ns_(ns) {}
  double toSec() const { return 1e-9 * (double)ns_; }

 private:
  int64_t ns_;
};

class Time {
 public:
  Time() : ns_(0) {}
  static Time now() {
    Time t;
    t.ns_ = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return t;
  }
  uint64_t toNSec() const { return (uint64_t)ns_; }
  double toSec() const { return 1e-9 * (double)ns_; }
  Duration operator-(const Time& o) const { return Duration(ns_ - o.ns_); }

 private:
  int64_t ns_;
};

}  // namespace ros

#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#endif
