// ParticleFilter.h -- particle filter core of amcl3d, B200 edition.
//
// Source-compatible with amcl3d/src/ParticleFilter.h of the reference for the weighting path:
//   relocPose (and the upstream name init) / predict / update / particleNormalize / resample /
//   uniformSample / getMean / getBestParticle / setParticleNum / isInitialized,
//   public p_, mean_, grid3d_, localization_params_, max_resamp_num_, min_resamp_num_.
// update / normalise / resample / mean run as CUDA kernels behind include/amcl3d_b200.h.  p_ is a LAZILY SYNCED host
// mirror (ParticleMirror below: the std::vector surface callers use -- size, [], iteration, assignment, conversion to
// const std::vector<Particle>&): the particle set stays in HBM between device steps and crosses PCIe only when the
// host really looks at it (one D2H) or has edited it (one H2D before the next device step), so
// predict -> update -> particleNormalize -> resample -> getMean moves the set at most once each way.
// Callers may keep reading and editing p_ between calls exactly as they do with the reference.
// setDevices({0,1,...}) (or AMCL3D_B200_DEVICES=0,1,...) shards the particles over several GPUs of one NVLink box:
// the grid is replicated, predict / update / particleNormalize / resample / getMean run sharded with the exchange in
// the library's own peer-memory kernels (csrc/shard.cu), bit-identical to one GPU; every other step (uniformSample,
// addParticleWeight, getBestParticle, the MonteCarloLocalization steps) runs on the first device, and the set moves there
// and back DEVICE TO DEVICE over NVLink (amcl3d_b200_pf_shard_gather_local / _scatter_local), never through the host; the
// shards are sized for max(max_resamp_num_, the current count), so a changing particle count re-draws the partition
// without remaking them.  getMean of a sharded set is one f32 chain carried from device to device (setExactMean).  Host RNG (std::mt19937) as in the reference; setSeed() makes runs reproducible (the reference seeds
// from std::random_device).
#pragma once

#include <random>
#include <vector>

#include "amcl3d/Grid3d.h"

namespace amcl3d
{
/*! Reference: ParticleFilter.h:32-48 -- seven packed floats. */
struct Particle
{
  float x;
  float y;
  float z;
  float roll;
  float pitch;
  float yaw;
  float w;

  Particle() : x(0), y(0), z(0), roll(0), pitch(0), yaw(0), w(0) {}
  Particle(float _x, float _y, float _z, float _roll, float _pitch, float _yaw)
    : x(_x), y(_y), z(_z), roll(_roll), pitch(_pitch), yaw(_yaw), w(-1)
  {
  }
};
static_assert(sizeof(Particle) == 28, "Particle must stay 7 packed floats (C-ABI layout)");

class ParticleFilter;

/*! p_: the std::vector<Particle> surface of the reference over a particle set that lives on the device(s).
 *  Reading (const access, conversion to const std::vector<Particle>&, size()) pulls the set to the host once if a device
 *  step changed it; any non-const access additionally marks the device copy stale, so the next device step uploads. */
class ParticleMirror
{
public:
  ParticleMirror() {}
  ~ParticleMirror() { unpin(); }
  ParticleMirror(const ParticleMirror&) = delete;
  ParticleMirror& operator=(const ParticleMirror& o) { return *this = o.host(); }  // pf.p_ = other.p_ copies the set
  typedef Particle value_type;
  typedef std::vector<Particle>::iterator iterator;
  typedef std::vector<Particle>::const_iterator const_iterator;

  size_t size() const { return host_fresh_ ? host_.size() : dev_n_; }
  bool empty() const { return size() == 0; }
  const Particle& operator[](size_t i) const { return host()[i]; }
  Particle& operator[](size_t i) { return edit()[i]; }
  const Particle& at(size_t i) const { return host().at(i); }
  Particle& at(size_t i) { return edit().at(i); }
  const Particle& front() const { return host().front(); }
  const Particle& back() const { return host().back(); }
  Particle& front() { return edit().front(); }
  Particle& back() { return edit().back(); }
  const Particle* data() const { return host().data(); }
  Particle* data() { return edit().data(); }
  const_iterator begin() const { return host().begin(); }
  const_iterator end() const { return host().end(); }
  const_iterator cbegin() const { return host().begin(); }
  const_iterator cend() const { return host().end(); }
  iterator begin() { return edit().begin(); }
  iterator end() { return edit().end(); }
  operator const std::vector<Particle>&() const { return host(); }

  // whole-set replacement: nothing is pulled from the device first
  ParticleMirror& operator=(const std::vector<Particle>& v)
  {
    will_hold(v.size());
    host_ = v;
    replaced();
    return *this;
  }
  ParticleMirror& operator=(std::vector<Particle>&& v)
  {
    unpin();  // the buffer itself is replaced
    host_ = std::move(v);
    replaced();
    return *this;
  }
  void assign(size_t n, const Particle& v)
  {
    will_hold(n);
    host_.assign(n, v);
    replaced();
  }
  template <class It>
  void assign(It first, It last)
  {
    unpin();
    host_.assign(first, last);
    replaced();
  }
  void clear()
  {
    host_.clear();
    replaced();
  }
  void swap(std::vector<Particle>& v)
  {
    unpin();
    edit().swap(v);
  }
  void resize(size_t n)
  {
    will_hold(n);
    edit().resize(n);
  }
  void resize(size_t n, const Particle& v)
  {
    will_hold(n);
    edit().resize(n, v);
  }
  void reserve(size_t n)
  {
    will_hold(n);
    host_.reserve(n);
  }
  void push_back(const Particle& v)
  {
    will_hold(size() + 1);
    edit().push_back(v);
  }

  /*! Read-only view, pulled from the device if a device step changed the set since the last look. */
  const std::vector<Particle>& host() const;
  /*! Editable view: as host(), and the device copy becomes stale. */
  std::vector<Particle>& edit()
  {
    host();
    dev_fresh_ = false;
    return host_;
  }
  /*! Number of host<->device transfers of the whole set so far (tests / benchmarks). */
  uint64_t uploads() const { return uploads_; }
  uint64_t downloads() const { return downloads_; }

private:
  friend class ParticleFilter;
  void replaced()
  {
    host_fresh_ = true;
    dev_fresh_ = false;
  }
  // the vector's buffer is page-locked once it is large enough to matter (>= 256 KB) and re-registered when the vector
  // reallocated: the set then crosses PCIe as one DMA transfer each way
  void pin() const
  {
    const void* ptr = host_.data();
    const size_t bytes = host_.capacity() * sizeof(Particle);
    if (ptr == pinned_ptr_ && bytes == pinned_bytes_)
      return;
    unpin();
    if (bytes >= (256u << 10) && amcl3d_b200_host_pin(const_cast<void*>(ptr), bytes) == AMCL3D_B200_OK)
    {
      pinned_ptr_ = ptr;
      pinned_bytes_ = bytes;
    }
  }
  // a page-locked buffer is released BEFORE the vector may reallocate it away
  void will_hold(size_t n) const
  {
    if (n > host_.capacity())
      unpin();
  }
  void unpin() const
  {
    if (pinned_ptr_)
      amcl3d_b200_host_unpin(const_cast<void*>(pinned_ptr_));
    pinned_ptr_ = nullptr;
    pinned_bytes_ = 0;
  }
  mutable const void* pinned_ptr_{ nullptr };
  mutable size_t pinned_bytes_{ 0 };
  mutable std::vector<Particle> host_;
  mutable bool host_fresh_{ true };  // host_ holds the current set
  bool dev_fresh_{ false };          // the device(s) hold the current set
  size_t dev_n_{ 0 };
  const ParticleFilter* owner_{ nullptr };
  mutable uint64_t uploads_{ 0 }, downloads_{ 0 };
};

/*! Reference: ParticleFilter.h:50-99 (defaults as compiled there; init_pitch_dev_(0,2) evaluates to 2). */
struct LocalizationParam
{
  int min_particle_num_global{ 150 };
  int min_particle_num_local{ 50 };
  int max_particle_num_global{ 600 };
  int max_particle_num_local{ 200 };
  double init_x_dev_{ 0.5 }, init_y_dev_{ 0.5 }, init_z_dev_{ 0.4 };
  double init_roll_dev_{ 0.2 }, init_pitch_dev_{ 2 }, init_yaw_dev_{ 0.4 };
  double odom_x_mod_{ 0.4 }, odom_y_mod_{ 0.4 }, odom_z_mod_{ 0.05 };
  double odom_roll_mod_{ 0.05 }, odom_pitch_mod_{ 0.05 }, odom_yaw_mod_{ 0.2 };
  double min_xy_noise_{ 0.1 }, min_z_noise_{ 0.05 }, min_rp_noise_{ 0.05 }, min_yaw_noise_{ 0.05 };
  double weight_global_{ 4.0 };
};

class ParticleFilter
{
public:
  explicit ParticleFilter(VSCOMMON::LoggerPtr& log, int device = 0);
  explicit ParticleFilter(int device = 0);
  /*! On a grid that lives behind an existing C-ABI handle (borrowed, not destroyed): resident / benchmark callers. */
  explicit ParticleFilter(amcl3d_b200_grid_t borrowed_grid);
  virtual ~ParticleFilter();
  ParticleFilter(const ParticleFilter&) = delete;
  ParticleFilter& operator=(const ParticleFilter&) = delete;

  bool isInitialized() const { return initialized_; }

  Particle getMean();
  Particle getBestParticle();

  /*! PF init around a pose (reference: ParticleFilter.cpp:31-89); host RNG, six draws per particle. */
  void relocPose(const int num_particles, const float x_init, const float y_init, const float z_init,
                 const float roll_init, const float pitch_init, const float yaw_init, const float x_dev,
                 const float y_dev, const float z_dev, const float roll_dev, const float pitch_dev, const float yaw_dev);
  /*! Upstream name of relocPose (north_star / reference tests call it init). */
  void init(const int num_particles, const float x_init, const float y_init, const float z_init, const float roll_init,
            const float pitch_init, const float yaw_init, const float x_dev, const float y_dev, const float z_dev,
            const float roll_dev, const float pitch_dev, const float yaw_dev)
  {
    relocPose(num_particles, x_init, y_init, z_init, roll_init, pitch_init, yaw_init, x_dev, y_dev, z_dev, roll_dev,
              pitch_dev, yaw_dev);
  }

  /*! Odometry prediction with Gaussian noise (reference: ParticleFilter.cpp:91-112): noise drawn on the host
   *  in the reference's order, the particle update itself runs on the device. */
  void predict(const double delta_x, const double delta_y, const double delta_z, const double delta_roll,
               const double delta_pitch, const double delta_yaw, const double odom_x_noise, const double odom_y_noise,
               const double odom_z_noise, const double odom_roll_noise, const double odom_pitch_noise,
               const double odom_yaw_noise);

  /*! Measurement update (reference: ParticleFilter.cpp:114-139): w = 0 outside the map, else computeCloudWeight. */
  void update(const pcl::PointCloud<PointType>::Ptr& cloud);
  /*! Same with the range-only beacon weight fused in (row R of the survey; not in the reference tree):
   *  w = alpha*wp/sum(wp) + (1-alpha)*wr/sum(wr). */
  void update(const pcl::PointCloud<PointType>::Ptr& cloud, const std::vector<amcl3d_b200_beacon_t>& range_data,
              const double alpha, const double sigma);

  void particleNormalize();
  void resample();
  void uniformSample(int& sample_num);
  void setParticleNum(int max_sample = 800, int min_sample = 150);

  /*! MonteCarloLocalization::addParticleWeight (reference: amcl3d.cpp:205-217), the step between update and
   *  uniformSample in global mode. */
  void addParticleWeight();

  /*! Reproducibility hooks (the reference seeds from std::random_device and keeps the draw private). */
  void setSeed(uint32_t seed) { generator_.seed(seed); }
  float lastResampleDraw() const { return last_u01_; }
  /*! Device time of the kernels of the last device step, ms. */
  float lastKernelMs() const;
  /*! The C-ABI handle of the first device, holding the current set (uploads p_ first if the host edited it). */
  amcl3d_b200_pf_t handle();

  /*! Shards the particles over these CUDA devices of one box (grid replicated, see the header comment); one device =
   *  the plain path.  Also read from AMCL3D_B200_DEVICES at construction.  The first device must be grid3d_'s. */
  void setDevices(const std::vector<int>& devices);
  const std::vector<int>& devices() const { return devices_; }
  /*! getMean() of a sharded set: true (default) = one sequential f32 chain carried from device to device, bit-identical to
   *  one device (serial over the devices: about 3 us per 1000 particles in total); false = per-device partial sums combined
   *  in f64 (tens of microseconds, last-ulp differences).  One device always runs the exact chain. */
  void setExactMean(bool exact) { exact_mean_ = exact; }
  /*! Device-to-device moves of the whole set between the shards and the first device's handle so far (tests). */
  uint64_t deviceGathers() const { return gathers_; }
  uint64_t deviceScatters() const { return scatters_; }

public:
  ParticleMirror p_;
  Particle mean_;
  Grid3d* grid3d_;
  LocalizationParam localization_params_;
  int max_resamp_num_{ 800 }, min_resamp_num_{ 150 };

private:
protected:
  /*! Before a device step on the first device's handle: the current set is there (uploaded / gathered if need be). */
  void syncDevice();
  /*! After a device step on the first device's handle that changed the set (n = its new size). */
  void deviceChanged();

private:
  friend class ParticleMirror;
  float ranGaussian(const double mean, const double sigma);
  float rngUniform(const float range_from, const float range_to);
  void pull(std::vector<Particle>& out) const;  // device(s) -> host
  bool sharded() const { return devices_.size() > 1 && p_.size() >= devices_.size(); }
  void syncShards();     // the current set is spread over the shard handles
  void shardsChanged();  // a sharded device step changed the set
  void dropShards();
  void setCloudAll(const pcl::PointCloud<PointType>::Ptr& cloud);

  bool initialized_{ false };
  std::random_device rd_;
  std::mt19937 generator_;
  VSCOMMON::LoggerPtr g_log;
  amcl3d_b200_pf_t handle_{ nullptr };
  float last_u01_{ 0.f };
  // multi-GPU: rank r = shard_[r] on devices_[r]; replicas_[r] (r >= 1) are clones of grid3d_'s cells
  std::vector<int> devices_;
  std::vector<amcl3d_b200_pf_t> shard_;
  std::vector<amcl3d_b200_grid_t> replicas_;
  uint64_t shard_n_{ 0 }, shard_cap_{ 0 }, replica_generation_{ 0 };
  uint64_t gathers_{ 0 }, scatters_{ 0 };
  bool exact_mean_{ true };
  enum Where { kNowhere, kSingle, kSharded };
  Where where_{ kNowhere };  // which device-side home holds the set when p_.dev_fresh_
};

}  // namespace amcl3d
