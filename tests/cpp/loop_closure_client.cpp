// A client that drives the facade classes with the call sequence LoopClosure uses
// (constructor: fast_lio_sam_qn/src/loop_closure.cpp:9-27; icpAlignment :110-136; coarseToFineAlignment :138-159),
// written against the same member names so that it doubles as the "does the drop-in link" test.
// usage: loop_closure_client src.bin dst.bin mode(gicp|quatro)   (bin = n x 4 float32: x y z intensity)
#define B200REG_HOST_IMPLEMENTATION
#include <nano_gicp/nano_gicp.hpp>
#include <nano_gicp/point_type_nano_gicp.hpp>
#include <quatro/quatro_module.h>

#if defined(B200REG_EXPECT_PCL) && !defined(B200REG_HAVE_PCL)
#error "the PCL/Eigen configuration of the facade was requested but b200reg_compat.hpp did not select it"
#endif

#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <string>

struct RegistrationOutput {  // fast_lio_sam_qn/include/loop_closure.h:64-70
  bool is_valid_ = false;
  bool is_converged_ = false;
  double score_ = std::numeric_limits<double>::max();
  Eigen::Matrix4d pose_between_eig_ = Eigen::Matrix4d::Identity();
};

static pcl::PointCloud<PointType> load(const char* path) {
  pcl::PointCloud<PointType> c;
  std::ifstream f(path, std::ios::binary);
  float v[4];
  while (f.read(reinterpret_cast<char*>(v), sizeof(v))) {
    PointType p;
    p.x = v[0];
    p.y = v[1];
    p.z = v[2];
    p.intensity = v[3];
    c.push_back(p);
  }
  return c;
}

static pcl::PointCloud<PointType> transformPcd(const pcl::PointCloud<PointType>& in, const Eigen::Matrix4d& T) {
  // utilities.hpp:164-175 (pcl::transformPointCloud with a Matrix4d: double math, float result)
  pcl::PointCloud<PointType> out = in;
  for (size_t i = 0; i < in.size(); i++) {
    const double x = in[i].x, y = in[i].y, z = in[i].z;
    out[i].x = static_cast<float>(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3));
    out[i].y = static_cast<float>(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3));
    out[i].z = static_cast<float>(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3));
  }
  return out;
}

struct Client {
  nano_gicp::NanoGICP<PointType, PointType> nano_gicp_;
  std::shared_ptr<quatro<PointType>> quatro_handler_;
  pcl::PointCloud<PointType> aligned_, coarse_aligned_;
  double icp_score_thr_ = 1.5;

  Client() {
    nano_gicp_.setNumThreads(0);
    nano_gicp_.setCorrespondenceRandomness(15);
    nano_gicp_.setMaximumIterations(32);
    nano_gicp_.setRANSACIterations(5);
    nano_gicp_.setMaxCorrespondenceDistance(52.5);
    nano_gicp_.setTransformationEpsilon(0.01);
    nano_gicp_.setEuclideanFitnessEpsilon(0.01);
    nano_gicp_.setRANSACOutlierRejectionThreshold(1.0);
    quatro_handler_ = std::make_shared<quatro<PointType>>(0.9, 1.5, 0.3, 1.4, 0.0001, 50, false, true, 35.0, 200);
  }
  RegistrationOutput icpAlignment(const pcl::PointCloud<PointType>& src, const pcl::PointCloud<PointType>& dst) {
    RegistrationOutput reg_output;
    aligned_.clear();
    pcl::PointCloud<PointType>::Ptr src_cloud(new pcl::PointCloud<PointType>());
    pcl::PointCloud<PointType>::Ptr dst_cloud(new pcl::PointCloud<PointType>());
    *src_cloud = src;
    *dst_cloud = dst;
    nano_gicp_.setInputSource(src_cloud);
    nano_gicp_.calculateSourceCovariances();
    nano_gicp_.setInputTarget(dst_cloud);
    nano_gicp_.calculateTargetCovariances();
    nano_gicp_.align(aligned_);
    reg_output.score_ = nano_gicp_.getFitnessScore();
    if (nano_gicp_.hasConverged() && reg_output.score_ < icp_score_thr_) {
      reg_output.is_valid_ = true;
      reg_output.is_converged_ = true;
      reg_output.pose_between_eig_ = nano_gicp_.getFinalTransformation().cast<double>();
    }
    return reg_output;
  }
  RegistrationOutput coarseToFineAlignment(const pcl::PointCloud<PointType>& src, const pcl::PointCloud<PointType>& dst) {
    RegistrationOutput reg_output;
    coarse_aligned_.clear();
    reg_output.pose_between_eig_ = quatro_handler_->align(src, dst, reg_output.is_converged_);
    if (!reg_output.is_converged_) return reg_output;
    coarse_aligned_ = transformPcd(src, reg_output.pose_between_eig_);
    const auto fine_output = icpAlignment(coarse_aligned_, dst);
    const auto quatro_tf_ = reg_output.pose_between_eig_;
    reg_output = fine_output;
    reg_output.pose_between_eig_ = fine_output.pose_between_eig_ * quatro_tf_;
    return reg_output;
  }
};

// The rest of the class surface (nano_gicp.hpp:84-106, lsq_registration.hpp:84-92) that fast_lio_sam_qn/src does not call:
// covariance getters / setters with the reference's container type, getFinalHessian, swapSourceAndTarget, timing of a
// second align() on unchanged inputs (the single-pair latency a caller of NanoGICP::align sees).
static int surface(const pcl::PointCloud<PointType>& src, const pcl::PointCloud<PointType>& dst) {
  using Gicp = nano_gicp::NanoGICP<PointType, PointType>;
  Client c;
  const RegistrationOutput r1 = c.icpAlignment(src, dst);
  const Eigen::Matrix<double, 6, 6> H1 = c.nano_gicp_.getFinalHessian();
  // round trip: the computed covariances handed back through the setters must reproduce the registration bit for bit
  std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> cs = c.nano_gicp_.getSourceCovariances();
  std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> ct = c.nano_gicp_.getTargetCovariances();
  Gicp g2;
  g2.setCorrespondenceRandomness(15);
  g2.setMaximumIterations(32);
  g2.setMaxCorrespondenceDistance(52.5);
  g2.setTransformationEpsilon(0.01);
  pcl::PointCloud<PointType>::Ptr sp(new pcl::PointCloud<PointType>(src)), dp(new pcl::PointCloud<PointType>(dst));
  g2.setInputSource(sp);
  g2.setInputTarget(dp);
  g2.setSourceCovariances(cs);
  g2.setTargetCovariances(ct);
  pcl::PointCloud<PointType> out;
  g2.align(out);
  const Eigen::Matrix4f T1 = c.nano_gicp_.getFinalTransformation(), T2 = g2.getFinalTransformation();
  int same = 1;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) same &= T1(i, j) == T2(i, j);
  // scaled covariances change the weighting, hence (slightly) the optimum: the setters are really used
  for (auto& m : cs)
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) m(i, j) *= (i == 2 || j == 2) ? 4.0 : 1.0;
  g2.setSourceCovariances(cs);
  g2.align(out);
  const Eigen::Matrix4f T3 = g2.getFinalTransformation();
  int differs = 0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) differs |= T1(i, j) != T3(i, j);
  int cov_struct = cs.size() == src.size();
  for (size_t i = 0; i < ct.size(); i += 97) cov_struct &= ct[i](3, 3) == 0.0 && ct[i](0, 3) == 0.0 && ct[i](0, 1) == ct[i](1, 0);
  std::printf("{\"valid\": %d, \"setters_roundtrip_same\": %d, \"scaled_covs_differ\": %d, \"cov_struct\": %d, \"score\": %.17g, \"H\": [",
              r1.is_valid_ ? 1 : 0, same, differs, cov_struct, r1.score_);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) std::printf("%.17g%s", H1(i, j), (i == 5 && j == 5) ? "" : ", ");
  std::printf("]}\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const auto src = load(argv[1]), dst = load(argv[2]);
  if (std::string(argv[3]) == "surface") return surface(src, dst);
  Client c;
  const RegistrationOutput r = std::string(argv[3]) == "quatro" ? c.coarseToFineAlignment(src, dst) : c.icpAlignment(src, dst);
  std::printf("{\"valid\": %d, \"converged\": %d, \"score\": %.17g, \"aligned\": %zu, \"T\": [", r.is_valid_ ? 1 : 0, r.is_converged_ ? 1 : 0,
              r.score_, c.aligned_.size());
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) std::printf("%.17g%s", r.pose_between_eig_(i, j), (i == 3 && j == 3) ? "" : ", ");
  std::printf("], \"aligned0\": [%.9g, %.9g, %.9g, %.9g]}\n", c.aligned_.size() ? c.aligned_[0].x : 0.f, c.aligned_.size() ? c.aligned_[0].y : 0.f,
              c.aligned_.size() ? c.aligned_[0].z : 0.f, c.aligned_.size() ? c.aligned_[0].intensity : 0.f);
  return 0;
}
