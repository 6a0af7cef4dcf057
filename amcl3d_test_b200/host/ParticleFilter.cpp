// ParticleFilter.cpp -- host shim of the B200 particle filter over the C-ABI.
#include "amcl3d/ParticleFilter.h"

#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace amcl3d
{
namespace
{
constexpr uint32_t kStride = sizeof(PointType) / sizeof(float);

void check(int rc, const char* what)
{
  if (rc != AMCL3D_B200_OK)
    throw std::runtime_error(std::string(what) + ": " + amcl3d_b200_last_error());
}
}  // namespace

ParticleFilter::ParticleFilter(VSCOMMON::LoggerPtr& log, int device)
  : grid3d_(new Grid3d(log, device)), generator_(rd_()), g_log(log)
{
  check(amcl3d_b200_pf_create(grid3d_->handle(), &handle_), "ParticleFilter");
}

ParticleFilter::ParticleFilter(int device) : grid3d_(new Grid3d(device)), generator_(rd_()), g_log(new VSCOMMON::Logger("PF"))
{
  check(amcl3d_b200_pf_create(grid3d_->handle(), &handle_), "ParticleFilter");
}

ParticleFilter::~ParticleFilter()
{
  amcl3d_b200_pf_destroy(handle_);
  delete grid3d_;  // the reference leaks it (ParticleFilter.cpp:23); owning it here is not observable by callers
}

void ParticleFilter::upload()
{
  check(amcl3d_b200_pf_set_particles(handle_, p_.empty() ? nullptr : &p_[0].x, p_.size()), "upload p_");
}

void ParticleFilter::download()
{
  uint64_t n = 0;
  check(amcl3d_b200_pf_size(handle_, &n), "pf_size");
  p_.resize(n);
  check(amcl3d_b200_pf_get_particles(handle_, p_.empty() ? nullptr : &p_[0].x, p_.size(), &n), "download p_");
}

float ParticleFilter::lastKernelMs() const
{
  float ms = 0.f;
  amcl3d_b200_pf_last_kernel_ms(handle_, &ms);
  return ms;
}

// The pose cloud and its Gaussian weights are host work in the reference as well (six host RNG draws per
// particle decide everything); the arithmetic below follows ParticleFilter.cpp:36-86 operation for operation.
void ParticleFilter::relocPose(const int num_particles, const float x_init, const float y_init, const float z_init,
                               const float roll_init, const float pitch_init, const float yaw_init, const float x_dev,
                               const float y_dev, const float z_dev, const float roll_dev, const float pitch_dev,
                               const float yaw_dev)
{
  p_.assign(static_cast<size_t>(std::abs(num_particles)), Particle());
  if (p_.empty())
    return;
  const float spread = std::max(std::max(x_dev, y_dev), z_dev);
  const float k1 = 1. / (spread * sqrt(2 * M_PI));
  const float k2 = 1. / (2 * spread * spread);

  Particle& seed = p_[0];
  seed.x = x_init;
  seed.y = y_init;
  seed.z = z_init;
  seed.roll = roll_init;
  seed.pitch = pitch_init;
  seed.yaw = yaw_init;
  seed.w = k1;
  float total = seed.w;
  for (size_t i = 1; i < p_.size(); ++i)
  {
    Particle& q = p_[i];
    q.x = seed.x + ranGaussian(0, x_dev);
    q.y = seed.y + ranGaussian(0, y_dev);
    q.z = seed.z + ranGaussian(0, z_dev);
    q.roll = seed.roll + ranGaussian(0, roll_dev);
    q.pitch = seed.pitch + ranGaussian(0, pitch_dev);
    q.yaw = seed.yaw + ranGaussian(0, yaw_dev);
    const float dx = q.x - seed.x, dy = q.y - seed.y, dz = q.z - seed.z;
    const float dist = sqrt(dx * dx + dy * dy + dz * dz);
    q.w = k1 * exp(-dist * dist * k2);
    total += q.w;
  }
  Particle mean_p;
  for (auto& q : p_)
  {
    q.w /= total;
    mean_p.x += q.w * q.x;
    mean_p.y += q.w * q.y;
    mean_p.z += q.w * q.z;
    mean_p.roll += q.w * q.roll;
    mean_p.pitch += q.w * q.pitch;
    mean_p.yaw += q.w * q.yaw;
  }
  mean_ = mean_p;
  upload();
  initialized_ = true;
  g_log->info("mcl initialized.");
}

void ParticleFilter::predict(const double delta_x, const double delta_y, const double delta_z, const double delta_roll,
                             const double delta_pitch, const double delta_yaw, const double odom_x_noise,
                             const double odom_y_noise, const double odom_z_noise, const double odom_roll_noise,
                             const double odom_pitch_noise, const double odom_yaw_noise)
{
  // draws in the reference's order: x, y, z, roll, pitch, yaw per particle (ParticleFilter.cpp:103-110)
  std::vector<float> noise(p_.size() * 6);
  const double dev[6] = { odom_x_noise, odom_y_noise, odom_z_noise, odom_roll_noise, odom_pitch_noise, odom_yaw_noise };
  for (size_t i = 0; i < p_.size(); ++i)
    for (int k = 0; k < 6; ++k)
      noise[6 * i + k] = ranGaussian(0, dev[k]);
  const double delta[6] = { delta_x, delta_y, delta_z, delta_roll, delta_pitch, delta_yaw };
  upload();
  check(amcl3d_b200_pf_predict(handle_, delta, noise.empty() ? nullptr : noise.data()), "predict");
  download();
}

void ParticleFilter::update(const pcl::PointCloud<PointType>::Ptr& cloud)
{
  upload();
  const uint32_t n = cloud ? static_cast<uint32_t>(cloud->size()) : 0u;
  check(amcl3d_b200_pf_update_cloud(handle_, n ? &cloud->points[0].x : nullptr, kStride, n), "update");
  download();
}

void ParticleFilter::update(const pcl::PointCloud<PointType>::Ptr& cloud, const std::vector<amcl3d_b200_beacon_t>& range_data,
                            const double alpha, const double sigma)
{
  upload();
  const uint32_t n = cloud ? static_cast<uint32_t>(cloud->size()) : 0u;
  check(amcl3d_b200_pf_set_cloud(handle_, n ? &cloud->points[0].x : nullptr, kStride, n), "set_cloud");
  check(amcl3d_b200_pf_update(handle_, range_data.empty() ? nullptr : range_data.data(), static_cast<uint32_t>(range_data.size()),
                              static_cast<float>(alpha), static_cast<float>(sigma), nullptr),
        "update (cloud + range)");
  download();
}

void ParticleFilter::particleNormalize()
{
  upload();
  check(amcl3d_b200_pf_normalize(handle_, nullptr), "particleNormalize");
  download();
}

void ParticleFilter::addParticleWeight()
{
  upload();
  check(amcl3d_b200_pf_scale_by_max(handle_), "addParticleWeight");
  download();
}

Particle ParticleFilter::getMean()
{
  upload();
  Particle m;
  check(amcl3d_b200_pf_mean(handle_, /*exact=*/1, &m.x), "getMean");
  mean_ = m;
  return m;
}

Particle ParticleFilter::getBestParticle()
{
  upload();
  Particle b;
  check(amcl3d_b200_pf_best(handle_, &b.x), "getBestParticle");
  return b;
}

void ParticleFilter::resample()
{
  if (p_.empty())
    return;
  last_u01_ = rngUniform(0, 1);  // r = factor * rngUniform(0, 1) (reference: ParticleFilter.cpp:234)
  upload();
  check(amcl3d_b200_pf_resample(handle_, last_u01_, 0, 0, nullptr, nullptr), "resample");
  download();
}

void ParticleFilter::uniformSample(int& sample_num)
{
  upload();
  check(amcl3d_b200_pf_uniform_sample(handle_, sample_num, 0, 0, nullptr, nullptr, nullptr), "uniformSample");
  download();
}

void ParticleFilter::setParticleNum(int max_sample, int min_sample)
{
  max_resamp_num_ = max_sample;
  min_resamp_num_ = min_sample;
}

// a fresh distribution object per draw, like the reference (ParticleFilter.cpp:282-292): no cached second variate
float ParticleFilter::ranGaussian(const double mean, const double sigma)
{
  std::normal_distribution<float> distribution(mean, sigma);
  return distribution(generator_);
}

float ParticleFilter::rngUniform(const float range_from, const float range_to)
{
  std::uniform_real_distribution<float> distribution(range_from, range_to);
  return distribution(generator_);
}

}  // namespace amcl3d
