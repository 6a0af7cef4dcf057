// amcl3d.h -- the localisation steps that call the weighting path, B200 edition.
//
// Source-compatible with the part of amcl3d/src/amcl3d.h of the reference that sits directly on the path:
//   RelocPose / PFMove / PFResample / getParticle / globalLocalizationDone, and (private there) addParticleWeight /
//   computeSampleNum / getMapHeight / updateSampleHeight.  processFrame itself (input voxel filter, Eigen odometry
//   increment, global/local mode switching, amcl3d.cpp:45-163) is host control logic around these calls and is not
//   rebuilt; it keeps working on top of this class.
#pragma once

#include "amcl3d/ParticleFilter.h"

namespace amcl3d
{
class MonteCarloLocalization : public ParticleFilter
{
public:
  explicit MonteCarloLocalization(VSCOMMON::LoggerPtr& log, int device = 0) : ParticleFilter(log, device) {}
  explicit MonteCarloLocalization(int device = 0) : ParticleFilter(device) {}

  /*! Reference: amcl3d.cpp:165-181 -- snaps z to the nearest keypose, then ParticleFilter::relocPose with the
   *  init deviations of localization_params_, and selects the global particle-count window. */
  void RelocPose(const int num_particles, const float x_init, const float y_init, const float z_init, const float roll_init,
                 const float pitch_init, const float yaw_init);

  /*! Reference: amcl3d.cpp:183-190. */
  void PFMove(const Particle& odom_increment, const Particle& odom_increment_noise);

  /*! Reference: amcl3d.cpp:192-203: [addParticleWeight] -> computeSampleNum -> uniformSample -> updateSampleHeight.
   *  One C-ABI call; the particle set stays on the device between the four steps. */
  void PFResample(const bool weight_multiply = false);

  /*! Reference: ds_.setLeafSize(voxel_size x3) (amcl3d.cpp:37). */
  void setVoxelSize(const double voxel_size) { voxel_size_ = static_cast<float>(voxel_size); }
  /*! Reference: ds_.setInputCloud(input_cloud); ds_.filter(*cloud_ds) (amcl3d.cpp:51-52) -- pcl::VoxelGrid's
   *  centroid filter on the device.  The filtered cloud stays resident as the input of updateDownsampled();
   *  cloud_ds (optional) receives a host copy.  Returns the number of points kept. */
  int downsampleInput(const pcl::PointCloud<PointType>::Ptr& input_cloud, pcl::PointCloud<PointType>* cloud_ds = nullptr);
  /*! ParticleFilter::update on the cloud downsampleInput left on the device (no second upload). */
  void updateDownsampled();

  std::vector<Particle> getParticle() { return p_; }
  bool globalLocalizationDone() const { return global_loc_done_; }

  /*! Reference: amcl3d.cpp:242-288 (default cnt_per_grid = 3 as declared in amcl3d.h). */
  int computeSampleNum(const int cnt_per_grid = 3);
  /*! Reference: amcl3d.cpp:295-308 -- nearest keypose to (x, y, 0). */
  PointType getMapHeight(const float& x, const float& y);
  /*! Reference: amcl3d.cpp:310-321. */
  void updateSampleHeight();

private:
  std::vector<float> keyposeArray() const;
  float voxel_size_{ 0.1f };
  bool global_loc_done_{ false };
  Particle set_pose_;
};

}  // namespace amcl3d
