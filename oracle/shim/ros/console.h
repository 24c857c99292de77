// oracle/shim/ros/console.h — the ROS_* macros live in ros/ros.h of this shim (TEST INFRASTRUCTURE ONLY).
#include <ros/ros.h>
