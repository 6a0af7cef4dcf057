// amcl3d.cpp -- host shim of the localisation steps around the weighting path over the C-ABI.
#include "amcl3d/amcl3d.h"

#include <cstddef>
#include <limits>
#include <stdexcept>
#include <string>

namespace amcl3d
{
namespace
{
void check(int rc, const char* what)
{
  if (rc != AMCL3D_B200_OK)
    throw std::runtime_error(std::string(what) + ": " + amcl3d_b200_last_error());
}
}  // namespace

std::vector<float> MonteCarloLocalization::keyposeArray() const
{
  std::vector<float> kp;
  if (grid3d_ && grid3d_->keyposes_3d)
  {
    kp.reserve(grid3d_->keyposes_3d->size() * 3);
    for (const auto& p : grid3d_->keyposes_3d->points)
    {
      kp.push_back(p.x);
      kp.push_back(p.y);
      kp.push_back(p.z);
    }
  }
  return kp;
}

// exact 1-NN over the keyposes with FLANN's L2_Simple<float> distance (what the reference's kd-tree returns)
PointType MonteCarloLocalization::getMapHeight(const float& x, const float& y)
{
  if (!grid3d_ || !grid3d_->keyposes_3d || grid3d_->keyposes_3d->empty())
    throw std::runtime_error("getMapHeight: no keyposes loaded");
  const float q[3] = { x, y, 0.f };
  float best = std::numeric_limits<float>::infinity();
  size_t bi = 0;
  const auto& pts = grid3d_->keyposes_3d->points;
  for (size_t k = 0; k < pts.size(); ++k)
  {
    float result = 0.f, diff;
    diff = q[0] - pts[k].x;
    result += diff * diff;
    diff = q[1] - pts[k].y;
    result += diff * diff;
    diff = q[2] - pts[k].z;
    result += diff * diff;
    if (result < best)
    {
      best = result;
      bi = k;
    }
  }
  PointType out;
  out.x = pts[bi].x;
  out.y = pts[bi].y;
  out.z = pts[bi].z;
  return out;
}

void MonteCarloLocalization::RelocPose(const int num_particles, const float x_init, const float y_init, const float z_init,
                                       const float roll_init, const float pitch_init, const float yaw_init)
{
  (void)z_init;  // the reference replaces it by the height of the nearest keypose (amcl3d.cpp:169-174)
  const PointType nearest = getMapHeight(x_init, y_init);
  set_pose_ = Particle(x_init, y_init, nearest.z, roll_init, pitch_init, yaw_init);
  ParticleFilter::relocPose(num_particles, x_init, y_init, nearest.z, roll_init, pitch_init, yaw_init,
                            localization_params_.init_x_dev_, localization_params_.init_y_dev_,
                            localization_params_.init_z_dev_, localization_params_.init_roll_dev_,
                            localization_params_.init_pitch_dev_, localization_params_.init_yaw_dev_);
  setParticleNum(localization_params_.max_particle_num_global, localization_params_.min_particle_num_global);
  global_loc_done_ = false;
}

int MonteCarloLocalization::downsampleInput(const pcl::PointCloud<PointType>::Ptr& input_cloud, pcl::PointCloud<PointType>* cloud_ds)
{
  const uint32_t n = input_cloud ? static_cast<uint32_t>(input_cloud->size()) : 0u;
  constexpr uint32_t stride = sizeof(PointType) / sizeof(float);
  const int ioff = static_cast<int>(offsetof(PointType, intensity) / sizeof(float));
  uint32_t kept = 0;
  const amcl3d_b200_pf_t h = handle();
  check(amcl3d_b200_pf_set_cloud_downsampled(h, n ? &input_cloud->points[0].x : nullptr, stride, ioff, n, voxel_size_,
                                             (input_cloud && !input_cloud->is_dense) ? 1 : 0, &kept),
        "downsampleInput");
  if (cloud_ds)
  {
    std::vector<float> xyzi(static_cast<size_t>(kept) * 4);
    check(amcl3d_b200_pf_get_cloud(h, xyzi.data(), kept, &kept), "download filtered cloud");
    cloud_ds->points.resize(kept);
    for (uint32_t i = 0; i < kept; ++i)
    {
      PointType& q = cloud_ds->points[i];
      q.x = xyzi[4 * i];
      q.y = xyzi[4 * i + 1];
      q.z = xyzi[4 * i + 2];
      q.intensity = xyzi[4 * i + 3];
    }
    cloud_ds->width = kept;
    cloud_ds->height = 1;
    cloud_ds->is_dense = true;
  }
  return static_cast<int>(kept);
}

// the steps below run on the first device's handle; p_ stays on the device between them (lazy host mirror)
void MonteCarloLocalization::updateDownsampled()
{
  const amcl3d_b200_pf_t h = handle();
  check(amcl3d_b200_pf_update(h, nullptr, 0, 1.f, 1.f, nullptr), "update");
  deviceChanged();
}

void MonteCarloLocalization::PFMove(const Particle& d, const Particle& noise)
{
  ParticleFilter::predict(d.x, d.y, d.z, d.roll, d.pitch, d.yaw, noise.x, noise.y, noise.z, noise.roll, noise.pitch, noise.yaw);
}

int MonteCarloLocalization::computeSampleNum(const int cnt_per_grid)
{
  int s = 0;
  check(amcl3d_b200_pf_compute_sample_num(handle(), cnt_per_grid, min_resamp_num_, max_resamp_num_, &s), "computeSampleNum");
  return s;
}

void MonteCarloLocalization::updateSampleHeight()
{
  const std::vector<float> kp = keyposeArray();
  check(amcl3d_b200_pf_update_sample_height(handle(), kp.data(), kp.size() / 3, nullptr), "updateSampleHeight");
  deviceChanged();
}

void MonteCarloLocalization::PFResample(const bool weight_multiply)
{
  const std::vector<float> kp = keyposeArray();
  int new_size = 0;
  check(amcl3d_b200_pf_resample_step(handle(), weight_multiply ? 1 : 0, min_resamp_num_, max_resamp_num_,
                                     kp.empty() ? nullptr : kp.data(), kp.size() / 3, &new_size),
        "PFResample");
  deviceChanged();
}

}  // namespace amcl3d
