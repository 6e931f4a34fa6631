// TEST INFRASTRUCTURE — stand-ins for the LIO headers that lio/src/liw/lio_utils.h includes (they pull in Sophus, PCL, glog and
// tsl::robin_map, none of which is installed here). lio_utils.h itself and lidarFactor.{h,cpp} are compiled UNMODIFIED from
// /root/reference; these stubs only declare the names lio_utils.h's data structures mention (point3D, IMUPtr, cv::Mat).
#pragma once
#include <deque>
#include <memory>
#include <vector>
#include <Eigen/Dense>
namespace cv { class Mat {}; }
namespace zjloc {
struct point3D { Eigen::Vector3d raw_point, point; double intensity = 0, alpha_time = 0, relative_time = 0, timespan = 0; int ring = 0; };
struct IMU { double timestamp_ = 0; Eigen::Vector3d gyro_, acce_; };
using IMUPtr = std::shared_ptr<IMU>;
}
