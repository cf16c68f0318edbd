// Same role as third_party/nano_gicp/include/nano_gicp/point_type_nano_gicp.hpp:7 (global PointType typedef).
#pragma once
#include "../b200reg_compat.hpp"
using PointType = pcl::PointXYZI;
