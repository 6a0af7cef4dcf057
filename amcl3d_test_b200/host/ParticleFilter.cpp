// ParticleFilter.cpp -- host shim of the B200 particle filter over the C-ABI.
#include "amcl3d/ParticleFilter.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
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

std::vector<int> devices_from_env(int first)
{
  std::vector<int> d(1, first);
  const char* env = std::getenv("AMCL3D_B200_DEVICES");
  if (!env || !*env)
    return d;
  std::vector<int> parsed;
  const char* c = env;
  while (*c)
  {
    char* end = nullptr;
    const long v = std::strtol(c, &end, 10);
    if (end == c)
      break;
    parsed.push_back(static_cast<int>(v));
    c = (*end == ',') ? end + 1 : end;
  }
  if (parsed.size() > 1 && parsed[0] == first)
    return parsed;
  return d;
}
}  // namespace

// ---- ParticleMirror -------------------------------------------------------------------------------------------
const std::vector<Particle>& ParticleMirror::host() const
{
  if (!host_fresh_)
  {
    owner_->pull(host_);
    host_fresh_ = true;
    ++downloads_;
  }
  return host_;
}

// ---- construction ---------------------------------------------------------------------------------------------
ParticleFilter::ParticleFilter(VSCOMMON::LoggerPtr& log, int device)
  : grid3d_(new Grid3d(log, device)), generator_(rd_()), g_log(log)
{
  p_.owner_ = this;
  check(amcl3d_b200_pf_create(grid3d_->handle(), &handle_), "ParticleFilter");
  // stream order: a step that hands nothing back to the host (update, particleNormalize, resample ...) returns once it is
  // launched; the host waits where it receives something (getMean, getBestParticle, reading p_)
  check(amcl3d_b200_pf_set_async(handle_, 1), "set_async");
  devices_ = devices_from_env(device);
}

ParticleFilter::ParticleFilter(int device) : grid3d_(new Grid3d(device)), generator_(rd_()), g_log(new VSCOMMON::Logger("PF"))
{
  p_.owner_ = this;
  check(amcl3d_b200_pf_create(grid3d_->handle(), &handle_), "ParticleFilter");
  // stream order: a step that hands nothing back to the host (update, particleNormalize, resample ...) returns once it is
  // launched; the host waits where it receives something (getMean, getBestParticle, reading p_)
  check(amcl3d_b200_pf_set_async(handle_, 1), "set_async");
  devices_ = devices_from_env(device);
}

ParticleFilter::ParticleFilter(amcl3d_b200_grid_t borrowed_grid)
  : grid3d_(new Grid3d(borrowed_grid)), generator_(rd_()), g_log(new VSCOMMON::Logger("PF"))
{
  p_.owner_ = this;
  check(amcl3d_b200_pf_create(grid3d_->handle(), &handle_), "ParticleFilter");
  // stream order: a step that hands nothing back to the host (update, particleNormalize, resample ...) returns once it is
  // launched; the host waits where it receives something (getMean, getBestParticle, reading p_)
  check(amcl3d_b200_pf_set_async(handle_, 1), "set_async");
  amcl3d_b200_grid_info_t info;
  check(amcl3d_b200_grid_get_info(borrowed_grid, &info), "grid_get_info");
  devices_ = devices_from_env(info.device);
}

ParticleFilter::~ParticleFilter()
{
  dropShards();
  for (amcl3d_b200_grid_t g : replicas_)
    amcl3d_b200_grid_destroy(g);
  amcl3d_b200_pf_destroy(handle_);
  delete grid3d_;  // the reference leaks it (ParticleFilter.cpp:23); owning it here is not observable by callers
}

// ---- where the set lives ----------------------------------------------------------------------------------------
// p_.host_fresh_: the host vector is current.  p_.dev_fresh_: a device-side home is current -- where_ says which one:
// the first device's plain handle (kSingle) or the shard handles (kSharded).
void ParticleFilter::pull(std::vector<Particle>& out) const
{
  if (!p_.dev_fresh_)
    throw std::logic_error("ParticleFilter: neither the host nor a device holds the current set");
  p_.will_hold(p_.dev_n_);
  out.resize(p_.dev_n_);
  p_.pin();
  if (where_ == kSingle)
  {
    uint64_t n = 0;
    check(amcl3d_b200_pf_get_particles(handle_, out.empty() ? nullptr : &out[0].x, out.size(), &n), "download p_");
    return;
  }
  for (size_t r = 0; r < shard_.size(); ++r)
  {
    uint64_t lo = 0, hi = 0, n = 0;
    check(amcl3d_b200_pf_shard_info(shard_[r], &lo, &hi, nullptr), "shard_info");
    if (hi > lo)
      check(amcl3d_b200_pf_get_particles(shard_[r], &out[lo].x, hi - lo, &n), "download p_ (shard)");
  }
}

void ParticleFilter::syncDevice()
{
  if (p_.dev_fresh_ && where_ == kSingle)
    return;
  if (p_.dev_fresh_ && where_ == kSharded && !shard_.empty())
  {
    // the blocks of the shards into the first device's handle, device to device over NVLink: no host copy
    check(amcl3d_b200_pf_shard_gather_local(shard_.data(), static_cast<int>(shard_.size()), handle_), "gather the shards");
    ++gathers_;
    where_ = kSingle;  // the host copy, if current, stays current: the set did not change
    return;
  }
  const std::vector<Particle>& h = p_.host();
  p_.pin();
  check(amcl3d_b200_pf_set_particles(handle_, h.empty() ? nullptr : &h[0].x, h.size()), "upload p_");
  ++p_.uploads_;
  p_.dev_fresh_ = true;
  p_.dev_n_ = h.size();
  where_ = kSingle;
}

void ParticleFilter::deviceChanged()
{
  uint64_t n = 0;
  check(amcl3d_b200_pf_size(handle_, &n), "pf_size");
  p_.dev_n_ = n;
  p_.dev_fresh_ = true;
  p_.host_fresh_ = false;
  where_ = kSingle;
}

amcl3d_b200_pf_t ParticleFilter::handle()
{
  syncDevice();
  return handle_;
}

void ParticleFilter::dropShards()
{
  for (amcl3d_b200_pf_t h : shard_)
    amcl3d_b200_pf_destroy(h);
  shard_.clear();
  shard_n_ = 0;
  shard_cap_ = 0;
}

void ParticleFilter::setDevices(const std::vector<int>& devices)
{
  if (devices.empty())
    throw std::invalid_argument("setDevices: empty device list");
  amcl3d_b200_grid_info_t info;
  check(amcl3d_b200_grid_get_info(grid3d_->handle(), &info), "grid_get_info");
  if (devices[0] != info.device)
    throw std::invalid_argument("setDevices: the first device must be the one grid3d_ lives on");
  if (where_ == kSharded && p_.dev_fresh_)
    p_.host();  // bring the set home before the shards go away
  if (where_ == kSharded)
    p_.dev_fresh_ = false;
  dropShards();
  for (amcl3d_b200_grid_t g : replicas_)
    amcl3d_b200_grid_destroy(g);
  replicas_.clear();
  devices_ = devices;
}

// the current set spread over the shard handles (replicas of the grid and the handles are (re)made when the grid or
// the particle count changed)
void ParticleFilter::syncShards()
{
  const size_t world = devices_.size();
  if (replicas_.size() != world || replica_generation_ != grid3d_->generation())
  {
    if (where_ == kSharded && p_.dev_fresh_)
      p_.host();
    if (where_ == kSharded)
      p_.dev_fresh_ = false;
    dropShards();
    for (amcl3d_b200_grid_t g : replicas_)
      amcl3d_b200_grid_destroy(g);
    replicas_.assign(world, nullptr);
    for (size_t r = 1; r < world; ++r)
      check(amcl3d_b200_grid_clone(grid3d_->handle(), devices_[r], 0, &replicas_[r]), "grid_clone");
    replica_generation_ = grid3d_->generation();
  }
  if (p_.dev_fresh_ && where_ == kSharded && shard_n_ == p_.dev_n_)
    return;
  const uint64_t n_now = p_.size();
  if (n_now < world)
    throw std::runtime_error("ParticleFilter: fewer particles than devices");
  if (shard_.size() != world || n_now > shard_cap_)
  {
    // exchange regions with room for the largest set the resampling policy can ask for (setParticleNum)
    dropShards();
    const uint64_t cap = std::max<uint64_t>(n_now, static_cast<uint64_t>(std::max(max_resamp_num_, 0)));
    shard_.assign(world, nullptr);
    for (size_t r = 0; r < world; ++r)
    {
      check(amcl3d_b200_pf_create(r == 0 ? grid3d_->handle() : replicas_[r], &shard_[r]), "pf_create (shard)");
      check(amcl3d_b200_pf_shard_init_capacity(shard_[r], n_now, cap, static_cast<int>(r), static_cast<int>(world)), "shard_init");
    }
    check(amcl3d_b200_pf_shard_attach_local(shard_.data(), static_cast<int>(world)), "shard_attach_local");
    shard_n_ = n_now;
    shard_cap_ = cap;
  }
  if (p_.dev_fresh_ && where_ == kSingle)
  {
    // the set is on the first device's handle (a step without a sharded form ran there): device to device over NVLink
    check(amcl3d_b200_pf_shard_scatter_local(handle_, shard_.data(), static_cast<int>(world)), "scatter to the shards");
    ++scatters_;
    shard_n_ = n_now;
    where_ = kSharded;
    return;
  }
  const std::vector<Particle>& h = p_.host();
  p_.pin();
  if (shard_n_ != n_now)
  {
    for (size_t r = 0; r < world; ++r)
      check(amcl3d_b200_pf_shard_resize(shard_[r], n_now), "shard_resize");
    shard_n_ = n_now;
  }
  for (size_t r = 0; r < world; ++r)
  {
    uint64_t lo = 0, hi = 0;
    check(amcl3d_b200_pf_shard_info(shard_[r], &lo, &hi, nullptr), "shard_info");
    check(amcl3d_b200_pf_set_particles(shard_[r], hi > lo ? &h[lo].x : nullptr, hi - lo), "upload p_ (shard)");
  }
  ++p_.uploads_;
  p_.dev_fresh_ = true;
  p_.dev_n_ = h.size();
  where_ = kSharded;
}

void ParticleFilter::shardsChanged()
{
  p_.dev_n_ = shard_n_;
  p_.dev_fresh_ = true;
  p_.host_fresh_ = false;
  where_ = kSharded;
}

float ParticleFilter::lastKernelMs() const
{
  float ms = 0.f;
  amcl3d_b200_pf_last_kernel_ms(where_ == kSharded && !shard_.empty() ? shard_[0] : handle_, &ms);
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
  initialized_ = true;
  g_log->info("mcl initialized.");
}

void ParticleFilter::predict(const double delta_x, const double delta_y, const double delta_z, const double delta_roll,
                             const double delta_pitch, const double delta_yaw, const double odom_x_noise,
                             const double odom_y_noise, const double odom_z_noise, const double odom_roll_noise,
                             const double odom_pitch_noise, const double odom_yaw_noise)
{
  // draws in the reference's order: x, y, z, roll, pitch, yaw per particle (ParticleFilter.cpp:103-110)
  const size_t n = p_.size();
  std::vector<float> noise(n * 6);
  const double dev[6] = { odom_x_noise, odom_y_noise, odom_z_noise, odom_roll_noise, odom_pitch_noise, odom_yaw_noise };
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 6; ++k)
      noise[6 * i + k] = ranGaussian(0, dev[k]);
  const double delta[6] = { delta_x, delta_y, delta_z, delta_roll, delta_pitch, delta_yaw };
  if (sharded())
  {
    syncShards();
    for (size_t r = 0; r < shard_.size(); ++r)
    {
      uint64_t lo = 0, hi = 0;
      check(amcl3d_b200_pf_shard_info(shard_[r], &lo, &hi, nullptr), "shard_info");
      check(amcl3d_b200_pf_predict(shard_[r], delta, hi > lo ? &noise[6 * lo] : nullptr), "predict (shard)");
    }
    shardsChanged();
    return;
  }
  syncDevice();
  check(amcl3d_b200_pf_predict(handle_, delta, noise.empty() ? nullptr : noise.data()), "predict");
  deviceChanged();
}

void ParticleFilter::setCloudAll(const pcl::PointCloud<PointType>::Ptr& cloud)
{
  const uint32_t n = cloud ? static_cast<uint32_t>(cloud->size()) : 0u;
  for (amcl3d_b200_pf_t h : shard_)
    check(amcl3d_b200_pf_set_cloud(h, n ? &cloud->points[0].x : nullptr, kStride, n), "set_cloud (shard)");
}

void ParticleFilter::update(const pcl::PointCloud<PointType>::Ptr& cloud)
{
  if (sharded())
  {
    syncShards();
    setCloudAll(cloud);
    // every device works on its own block at the same time: launch in stream order, then wait for all of them
    for (amcl3d_b200_pf_t h : shard_)
    {
      check(amcl3d_b200_pf_set_async(h, 1), "set_async");
      check(amcl3d_b200_pf_update(h, nullptr, 0, 1.f, 1.f, nullptr), "update (shard)");
      check(amcl3d_b200_pf_set_async(h, 0), "set_async");
    }
    for (amcl3d_b200_pf_t h : shard_)
      check(amcl3d_b200_pf_sync(h), "sync (shard)");
    shardsChanged();
    return;
  }
  syncDevice();
  const uint32_t n = cloud ? static_cast<uint32_t>(cloud->size()) : 0u;
  check(amcl3d_b200_pf_update_cloud(handle_, n ? &cloud->points[0].x : nullptr, kStride, n), "update");
  deviceChanged();
}

void ParticleFilter::update(const pcl::PointCloud<PointType>::Ptr& cloud, const std::vector<amcl3d_b200_beacon_t>& range_data,
                            const double alpha, const double sigma)
{
  syncDevice();  // the fused range weight needs two global sums: runs on the first device
  const uint32_t n = cloud ? static_cast<uint32_t>(cloud->size()) : 0u;
  check(amcl3d_b200_pf_set_cloud(handle_, n ? &cloud->points[0].x : nullptr, kStride, n), "set_cloud");
  check(amcl3d_b200_pf_update(handle_, range_data.empty() ? nullptr : range_data.data(), static_cast<uint32_t>(range_data.size()),
                              static_cast<float>(alpha), static_cast<float>(sigma), nullptr),
        "update (cloud + range)");
  deviceChanged();
}

void ParticleFilter::particleNormalize()
{
  if (sharded())
  {
    syncShards();
    check(amcl3d_b200_pf_shard_normalize_local(shard_.data(), static_cast<int>(shard_.size())), "particleNormalize (sharded)");
    shardsChanged();
    return;
  }
  syncDevice();
  check(amcl3d_b200_pf_normalize(handle_, nullptr), "particleNormalize");
  deviceChanged();
}

void ParticleFilter::addParticleWeight()
{
  syncDevice();
  check(amcl3d_b200_pf_scale_by_max(handle_), "addParticleWeight");
  deviceChanged();
}

Particle ParticleFilter::getMean()
{
  Particle m;
  if (sharded())
  {
    // exact (default): ONE sequential f32 chain carried from rank to rank, the bits one device produces (serial over the
    // ranks); fast (setExactMean(false)): partial sums per device combined in rank order in f64, last-ulp differences
    syncShards();
    if (exact_mean_)
      check(amcl3d_b200_pf_shard_mean_exact_local(shard_.data(), static_cast<int>(shard_.size()), &m.x), "getMean (sharded, exact)");
    else
      check(amcl3d_b200_pf_shard_mean_local(shard_.data(), static_cast<int>(shard_.size()), &m.x), "getMean (sharded)");
    mean_ = m;
    return m;
  }
  syncDevice();
  check(amcl3d_b200_pf_mean(handle_, /*exact=*/1, &m.x), "getMean");
  mean_ = m;
  return m;
}

Particle ParticleFilter::getBestParticle()
{
  syncDevice();
  Particle b;
  check(amcl3d_b200_pf_best(handle_, &b.x), "getBestParticle");
  return b;
}

void ParticleFilter::resample()
{
  if (p_.empty())
    return;
  last_u01_ = rngUniform(0, 1);  // r = factor * rngUniform(0, 1) (reference: ParticleFilter.cpp:234)
  if (sharded())
  {
    syncShards();
    check(amcl3d_b200_pf_shard_resample_local(shard_.data(), static_cast<int>(shard_.size()), last_u01_, /*normalize=*/0, nullptr,
                                              1.f),
          "resample (sharded)");
    shardsChanged();
    return;
  }
  syncDevice();
  check(amcl3d_b200_pf_resample(handle_, last_u01_, 0, 0, nullptr, nullptr), "resample");
  deviceChanged();
}

void ParticleFilter::uniformSample(int& sample_num)
{
  syncDevice();
  check(amcl3d_b200_pf_uniform_sample(handle_, sample_num, 0, 0, nullptr, nullptr, nullptr), "uniformSample");
  deviceChanged();
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

// Benchmark / test hook (bench.py "e2e_cpp_class", tests): `steps` frames of the reference's call sequence through the
// class on an existing grid handle -- p_ written by the host, update(cloud), particleNormalize(), resample(), getMean(),
// p_ read back by the host -- wall-clock ms per frame.  devices: NULL or n_devices CUDA ordinals to shard over.
extern "C" __attribute__((visibility("default"))) int amcl3d_host_bench_frames(amcl3d_b200_grid_t grid, const float* particles7,
                                                                                uint64_t n, const float* cloud_xyz, uint32_t npts,
                                                                                int steps, int warmup, double* ms_per_frame,
                                                                                float mean7[7], const int* devices, int n_devices,
                                                                                uint64_t* transfers2, double* phase_ms6)
{
  try
  {
    using namespace amcl3d;
    ParticleFilter pf(grid);
    if (devices && n_devices > 1)
      pf.setDevices(std::vector<int>(devices, devices + n_devices));
    pf.setSeed(4004);
    pcl::PointCloud<PointType>::Ptr cloud(new pcl::PointCloud<PointType>());
    cloud->points.resize(npts);
    for (uint32_t i = 0; i < npts; ++i)
    {
      cloud->points[i].x = cloud_xyz[3 * i];
      cloud->points[i].y = cloud_xyz[3 * i + 1];
      cloud->points[i].z = cloud_xyz[3 * i + 2];
    }
    cloud->width = npts;
    cloud->height = 1;
    const Particle* src = reinterpret_cast<const Particle*>(particles7);
    std::vector<Particle> fresh(src, src + n);
    Particle m;
    double checksum = 0;
    double phase[6] = { 0, 0, 0, 0, 0, 0 };
    uint64_t up0 = 0, down0 = 0;
    std::chrono::steady_clock::time_point t0;
    for (int s = 0; s < warmup + steps; ++s)
    {
      if (s == warmup)
      {
        t0 = std::chrono::steady_clock::now();
        up0 = pf.p_.uploads();
        down0 = pf.p_.downloads();
      }
      typedef std::chrono::steady_clock clk;
      const clk::time_point a0 = clk::now();
      pf.p_ = fresh;  // the host hands over this frame's particles
      const clk::time_point a1 = clk::now();
      pf.update(cloud);
      const clk::time_point a2 = clk::now();
      pf.particleNormalize();
      const clk::time_point a3 = clk::now();
      pf.resample();
      const clk::time_point a4 = clk::now();
      m = pf.getMean();
      const clk::time_point a5 = clk::now();
      const std::vector<Particle>& out = pf.p_;  // the host looks at the resampled set
      checksum += out[out.size() / 2].x;
      const clk::time_point a6 = clk::now();
      if (s >= warmup)
      {
        const clk::time_point t[7] = { a0, a1, a2, a3, a4, a5, a6 };
        for (int k = 0; k < 6; ++k)
          phase[k] += std::chrono::duration<double, std::milli>(t[k + 1] - t[k]).count();
      }
    }
    if (phase_ms6)
      for (int k = 0; k < 6; ++k)
        phase_ms6[k] = phase[k] / (steps > 0 ? steps : 1);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms_per_frame)
      *ms_per_frame = ms / (steps > 0 ? steps : 1);
    if (mean7)
    {
      const float v[7] = { m.x, m.y, m.z, m.roll, m.pitch, m.yaw, m.w };
      for (int k = 0; k < 7; ++k)
        mean7[k] = v[k];
    }
    if (transfers2)
    {
      transfers2[0] = pf.p_.uploads() - up0;
      transfers2[1] = pf.p_.downloads() - down0;
    }
    return checksum == checksum ? 0 : 1;
  }
  catch (const std::exception& e)
  {
    std::fprintf(stderr, "amcl3d_host_bench_frames: %s\n", e.what());
    return -1;
  }
}
