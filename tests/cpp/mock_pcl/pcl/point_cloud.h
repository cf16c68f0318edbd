// API-shaped stand-in for <pcl/point_cloud.h> (TEST ONLY): members the facades and LoopClosure's call sequence use
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;  // boost::shared_ptr up to PCL 1.10, std::shared_ptr from 1.11
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() {
    points.clear();
    width = 0;
  }
  void resize(std::size_t n) {
    points.resize(n);
    width = static_cast<std::uint32_t>(n);
  }
  void push_back(const PointT& p) {
    points.push_back(p);
    width = static_cast<std::uint32_t>(points.size());
  }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl
