// API-shaped stand-in for <pcl/point_types.h> (TEST ONLY): the two point types the reference instantiates
#pragma once
namespace pcl {
struct alignas(16) PointXYZ {
  union {
    float data[4];
    struct {
      float x, y, z;
    };
  };
  PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
};
struct alignas(16) PointXYZI {  // 32 bytes like PCL's (EIGEN_ALIGN16, intensity in its own 16-byte lane)
  union {
    float data[4];
    struct {
      float x, y, z;
    };
  };
  union {
    struct {
      float intensity;
    };
    float data_c[4];
  };
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
}  // namespace pcl
