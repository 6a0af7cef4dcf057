// ParticleFilter.h -- particle filter core of amcl3d, B200 edition.
//
// Source-compatible with amcl3d/src/ParticleFilter.h of the reference for the weighting path:
//   relocPose (and the upstream name init) / predict / update / particleNormalize / resample /
//   uniformSample / getMean / getBestParticle / setParticleNum / isInitialized,
//   public p_, mean_, grid3d_, localization_params_, max_resamp_num_, min_resamp_num_.
// update / normalise / resample / mean run as CUDA kernels behind include/amcl3d_b200.h; p_ is the host
// mirror: it is uploaded before and downloaded after every device step, so callers may keep editing it
// between calls exactly as they do with the reference.  Host RNG (std::mt19937) as in the reference;
// setSeed() makes runs reproducible (the reference seeds from std::random_device).
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
  amcl3d_b200_pf_t handle() const { return handle_; }

public:
  std::vector<Particle> p_;
  Particle mean_;
  Grid3d* grid3d_;
  LocalizationParam localization_params_;
  int max_resamp_num_{ 800 }, min_resamp_num_{ 150 };

private:
  float ranGaussian(const double mean, const double sigma);
  float rngUniform(const float range_from, const float range_to);
  void upload();
  void download();

  bool initialized_{ false };
  std::random_device rd_;
  std::mt19937 generator_;
  VSCOMMON::LoggerPtr g_log;
  amcl3d_b200_pf_t handle_{ nullptr };
  float last_u01_{ 0.f };
};

}  // namespace amcl3d
