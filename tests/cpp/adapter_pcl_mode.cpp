// Type-checks include/b200reg_pcl.hpp in PCL mode (-DB200REG_WITH_PCL) against a minimal PCL-1.12-shaped stub
// (tests/cpp/fake_pcl): the two-line patch of INTEGRATION.md section 2, as the nodes would write it — and checks that
// align() through the base-class pointer does NOT trigger pcl::Registration's lazy host kd-tree build over the target
// (the stub counts them), which would put hundreds of milliseconds of CPU work in front of the GPU solve.
// Exit codes: 3 = no GPU (engine refuses to construct), 0 = ok, anything else = failure.
#include <cstdio>
#include <memory>
#include <stdexcept>

#include "b200reg_pcl.hpp"

int main() {
  try {
    using Reg = pcl::Registration<pcl::PointXYZI, pcl::PointXYZI>;
    std::shared_ptr<Reg> registration;  // scanmatcher_component.h:93
    auto ndt = std::make_shared<b200reg::NormalDistributionsTransform<pcl::PointXYZI, pcl::PointXYZI>>();
    ndt->setResolution(2.0f);
    ndt->setTransformationEpsilon(0.01);
    ndt->setNeighborhoodSearchMethod(b200reg::DIRECT7);
    registration = ndt;  // :111
    auto target = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
    auto source = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
    for (int i = 0; i < 3000; i++) {
      pcl::PointXYZI p;
      p.x = 0.05f * (i % 60) - 1.5f + 0.37f * (i % 7);
      p.y = 0.05f * (i / 60) + 0.11f * (i % 5);
      p.z = (i % 3 == 0) ? 0.0f : 0.02f * (i % 11);
      target->points.push_back(p);
      p.x -= 0.1f;
      source->points.push_back(p);
    }
    registration->setInputTarget(target);
    registration->setInputSource(source);
    pcl::PointCloud<pcl::PointXYZI> output;
    registration->align(output, Eigen::Matrix4f::Identity());  // :353
    Eigen::Matrix4f T = registration->getFinalTransformation();
    const double fit = b200reg::getFitnessScore(*registration);  // gbs.cpp:231 with the one-line patch
    std::printf("converged=%d tx=%.3f fitness=%.5f host_tree_builds=%d host_fitness_calls=%d\n", (int)registration->hasConverged(),
                T.data()[12], fit, registration->host_tree_builds(), registration->host_fitness_calls());
    if (!registration->hasConverged()) return 10;
    if (registration->host_tree_builds() != 0) return 11;   // the adapter armed pcl::Registration's lazy kd-tree
    if (registration->host_fitness_calls() != 0) return 12; // fitness came from the host tree, not from the GPU
    if (output.size() != source->size() || output.points[5].x != source->points[5].x) return 13;  // output D2H is opt-in
    // a new target (map update) still does not build a host tree; the aligned cloud on request
    registration->setInputTarget(target);
    ndt->setComputeOutputCloud(true);
    registration->align(output);
    if (registration->host_tree_builds() != 0) return 14;
    if (output.points[5].x == source->points[5].x) return 15;
    // PCL's own behaviour back on request
    ndt->setKeepHostSearchTree(true);
    registration->setInputTarget(target);
    registration->align(output);
    if (registration->host_tree_builds() != 1) return 16;
    auto gicp = std::make_shared<b200reg::GeneralizedIterativeClosestPoint<pcl::PointXYZI, pcl::PointXYZI>>();
    gicp->setMaxCorrespondenceDistance(5.0);
    registration = gicp;
    return registration->hasConverged() ? 2 : 0;
  } catch (const std::exception& e) {
    std::printf("no GPU: %s\n", e.what());
    return 3;
  }
}
