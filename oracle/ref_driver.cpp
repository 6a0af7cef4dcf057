// ref_driver.cpp -- extern "C" harness around the UNMODIFIED reference classes.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; it is compiled together with the reference's own
// Grid3d.cpp / ParticleFilter.cpp / PointCloudTools.cpp (read in place from /root/reference, never
// copied) into oracle/_ref/libamcl3d_ref.so by oracle/Makefile.  It only builds inputs, calls the
// reference's public API and copies results out, so that tests can pin oracle/amcl3d_oracle.c (and,
// through it, the CUDA path) against the reference's real arithmetic.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

// the reference keeps its RNG and the .grid reader/writer private; the harness needs to seed the
// former and call the latter (all standard headers are already included above)
#define private public
#include "amcl3d.h"
#undef private

using amcl3d::Grid3d;
using amcl3d::Particle;
using amcl3d::ParticleFilter;
using amcl3d::PointType;

namespace
{
VSCOMMON::LoggerPtr g_logger(new VSCOMMON::Logger());

pcl::PointCloud<PointType>::Ptr make_cloud(const float* xyz, uint32_t stride, uint64_t n)
{
  pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>());
  c->points.resize(n);
  for (uint64_t i = 0; i < n; ++i)
  {
    c->points[i].x = xyz[i * stride + 0];
    c->points[i].y = xyz[i * stride + 1];
    c->points[i].z = xyz[i * stride + 2];
  }
  c->width = (uint32_t)n;
  return c;
}

void install_map(Grid3d* g, const double mn[3], const double mx[3], double resol, const float* prob, const uint32_t size[3],
                 double sensor_dev)
{
  g->pc_info_.reset(new amcl3d::PointCloudInfo());
  g->pc_info_->octo_min_x = mn[0];
  g->pc_info_->octo_min_y = mn[1];
  g->pc_info_->octo_min_z = mn[2];
  g->pc_info_->octo_max_x = mx[0];
  g->pc_info_->octo_max_y = mx[1];
  g->pc_info_->octo_max_z = mx[2];
  g->pc_info_->octo_resol = resol;
  if (prob)
  {
    g->grid_info_.reset(new amcl3d::Grid3dInfo());
    g->grid_info_->sensor_dev = sensor_dev;
    g->grid_info_->size_x = size[0];
    g->grid_info_->size_y = size[1];
    g->grid_info_->size_z = size[2];
    g->grid_info_->step_y = size[0];
    g->grid_info_->step_z = size[0] * size[1];
    const uint32_t cells = size[0] * size[1] * size[2];
    g->grid_info_->grid.resize(cells);
    for (uint32_t i = 0; i < cells; ++i)
      g->grid_info_->grid[i].prob = prob[i];
  }
}
}  // namespace

extern "C" {

// ---- Grid3d ------------------------------------------------------------------------------------
void* ref_grid_create()
{
  return new Grid3d(g_logger);
}
void ref_grid_destroy(void* g)
{
  delete static_cast<Grid3d*>(g);
}
void ref_grid_install(void* g, const double mn[3], const double mx[3], double resol, const float* prob,
                      const uint32_t size[3], double sensor_dev)
{
  install_map(static_cast<Grid3d*>(g), mn, mx, resol, prob, size, sensor_dev);
}
float ref_grid_cloud_weight(void* g, const float* xyz, uint32_t stride, uint64_t n, float tx, float ty, float tz, float roll,
                            float pitch, float yaw)
{
  auto cloud = make_cloud(xyz, stride, n);
  return static_cast<Grid3d*>(g)->computeCloudWeight(cloud, tx, ty, tz, roll, pitch, yaw);
}
int ref_grid_is_into_map(void* g, float x, float y, float z)
{
  return static_cast<Grid3d*>(g)->isIntoMap(x, y, z) ? 1 : 0;
}
int ref_grid_open(void* g, const char* map_dir, double sensor_dev)
{
  return static_cast<Grid3d*>(g)->open(map_dir, sensor_dev) ? 1 : 0;
}
int ref_grid_save(void* g, const char* path)
{
  return static_cast<Grid3d*>(g)->saveGrid(path) ? 1 : 0;
}
int ref_grid_load(void* g, const char* path, double sensor_dev)
{
  return static_cast<Grid3d*>(g)->loadGrid(path, sensor_dev) ? 1 : 0;
}
// copies geometry (+ cells if prob != NULL) out of the Grid3d
int ref_grid_get(void* gv, double mn[3], double mx[3], double* resol, uint32_t size[3], double* sensor_dev, float* prob,
                 uint64_t cap)
{
  Grid3d* g = static_cast<Grid3d*>(gv);
  if (g->pc_info_)
  {
    mn[0] = g->pc_info_->octo_min_x; mn[1] = g->pc_info_->octo_min_y; mn[2] = g->pc_info_->octo_min_z;
    mx[0] = g->pc_info_->octo_max_x; mx[1] = g->pc_info_->octo_max_y; mx[2] = g->pc_info_->octo_max_z;
    *resol = g->pc_info_->octo_resol;
  }
  if (!g->grid_info_)
    return 0;
  size[0] = g->grid_info_->size_x; size[1] = g->grid_info_->size_y; size[2] = g->grid_info_->size_z;
  *sensor_dev = g->grid_info_->sensor_dev;
  if (prob)
  {
    const uint64_t cells = g->grid_info_->grid.size();
    if (cells > cap)
      return -1;
    for (uint64_t i = 0; i < cells; ++i)
      prob[i] = g->grid_info_->grid[i].prob;
  }
  return 1;
}

// ---- grid builders (PointCloudTools.cpp) -------------------------------------------------------
// which: 0 = computeGrid (kd-tree NN, quirk), 2 = computeGrid2 (splat).  Bounds come from
// computePointCloud on the same cloud unless use_bounds != 0.  Returns cell count or -1 on throw.
int64_t ref_compute_grid(const float* xyz, uint32_t stride, uint64_t n, double resol, double sensor_dev, int which,
                         int use_bounds, double mn[3], double mx[3], uint32_t size[3], float* prob, uint64_t cap)
{
  try
  {
    auto cloud = make_cloud(xyz, stride, n);
    amcl3d::PointCloudInfo::Ptr pc = amcl3d::computePointCloud(cloud, resol);
    if (use_bounds)
    {
      pc->octo_min_x = mn[0]; pc->octo_min_y = mn[1]; pc->octo_min_z = mn[2];
      pc->octo_max_x = mx[0]; pc->octo_max_y = mx[1]; pc->octo_max_z = mx[2];
    }
    else
    {
      mn[0] = pc->octo_min_x; mn[1] = pc->octo_min_y; mn[2] = pc->octo_min_z;
      mx[0] = pc->octo_max_x; mx[1] = pc->octo_max_y; mx[2] = pc->octo_max_z;
    }
    amcl3d::Grid3dInfo::Ptr gi = (which == 2) ? amcl3d::computeGrid2(pc, sensor_dev) : amcl3d::computeGrid(pc, sensor_dev);
    size[0] = gi->size_x; size[1] = gi->size_y; size[2] = gi->size_z;
    const uint64_t cells = gi->grid.size();
    if (prob)
    {
      if (cells > cap)
        return -2;
      for (uint64_t i = 0; i < cells; ++i)
        prob[i] = gi->grid[i].prob;
    }
    return (int64_t)cells;
  }
  catch (std::exception&)
  {
    return -1;
  }
}

// ---- ParticleFilter ----------------------------------------------------------------------------
void* ref_pf_create(void* grid)
{
  ParticleFilter* pf = new ParticleFilter(g_logger);
  if (grid)
  {
    delete pf->grid3d_;  // the reference leaks this one (ParticleFilter.cpp:23); the harness swaps it
    pf->grid3d_ = static_cast<Grid3d*>(grid);
  }
  return pf;
}
void ref_pf_destroy(void* pf)
{
  delete static_cast<ParticleFilter*>(pf);
}
void ref_pf_seed(void* pf, uint32_t seed)
{
  static_cast<ParticleFilter*>(pf)->generator_.seed(seed);
}
void ref_pf_set(void* pfv, const float* aos7, uint64_t n)
{
  ParticleFilter* pf = static_cast<ParticleFilter*>(pfv);
  pf->p_.resize(n);
  static_assert(sizeof(Particle) == 28, "Particle is 7 packed floats");
  if (n)
    std::memcpy(pf->p_.data(), aos7, n * sizeof(Particle));
}
uint64_t ref_pf_size(void* pfv)
{
  return static_cast<ParticleFilter*>(pfv)->p_.size();
}
void ref_pf_get(void* pfv, float* aos7)
{
  ParticleFilter* pf = static_cast<ParticleFilter*>(pfv);
  if (!pf->p_.empty())
    std::memcpy(aos7, pf->p_.data(), pf->p_.size() * sizeof(Particle));
}
void ref_pf_update(void* pfv, const float* xyz, uint32_t stride, uint64_t n)
{
  auto cloud = make_cloud(xyz, stride, n);
  static_cast<ParticleFilter*>(pfv)->update(cloud);
}
void ref_pf_normalize(void* pfv)
{
  static_cast<ParticleFilter*>(pfv)->particleNormalize();
}
// resample(): draws r = factor * U(0,1) from generator_.  The harness peeks the draw on a copy of the
// generator so the caller can hand the same u01 to the oracle / CUDA path.
float ref_pf_resample(void* pfv)
{
  ParticleFilter* pf = static_cast<ParticleFilter*>(pfv);
  std::mt19937 peek = pf->generator_;
  std::uniform_real_distribution<float> distribution(0, 1);
  const float u01 = distribution(peek);
  pf->resample();
  return u01;
}
int ref_pf_uniform_sample(void* pfv, int sample_num)
{
  ParticleFilter* pf = static_cast<ParticleFilter*>(pfv);
  pf->uniformSample(sample_num);
  return (int)pf->p_.size();
}
void ref_pf_mean(void* pfv, float out7[7])
{
  Particle m = static_cast<ParticleFilter*>(pfv)->getMean();
  std::memcpy(out7, &m, sizeof(Particle));
}
void ref_pf_best(void* pfv, float out7[7])
{
  Particle m = static_cast<ParticleFilter*>(pfv)->getBestParticle();
  std::memcpy(out7, &m, sizeof(Particle));
}
// relocPose with a known seed; noise_out (6*(n-1) floats, optional) receives the ranGaussian results in
// draw order, replayed from a copy of the generator exactly as ParticleFilter.cpp:282-286 draws them.
void ref_pf_reloc_pose(void* pfv, uint32_t seed, int n, const float init[6], const float dev[6], float* noise_out)
{
  ParticleFilter* pf = static_cast<ParticleFilter*>(pfv);
  pf->generator_.seed(seed);
  if (noise_out)
  {
    std::mt19937 peek = pf->generator_;
    for (int i = 1; i < std::abs(n); ++i)
      for (int k = 0; k < 6; ++k)
      {
        std::normal_distribution<float> distribution(0, (double)dev[k]);
        noise_out[6 * (i - 1) + k] = distribution(peek);
      }
  }
  pf->relocPose(n, init[0], init[1], init[2], init[3], init[4], init[5], dev[0], dev[1], dev[2], dev[3], dev[4], dev[5]);
}
// predict with a known seed; noise_out (6*n floats, optional) as above, order x,y,z,roll,pitch,yaw
void ref_pf_predict(void* pfv, uint32_t seed, const double delta[6], const double noise_dev[6], float* noise_out)
{
  ParticleFilter* pf = static_cast<ParticleFilter*>(pfv);
  pf->generator_.seed(seed);
  if (noise_out)
  {
    std::mt19937 peek = pf->generator_;
    for (size_t i = 0; i < pf->p_.size(); ++i)
      for (int k = 0; k < 6; ++k)
      {
        std::normal_distribution<float> distribution(0, noise_dev[k]);
        noise_out[6 * i + k] = distribution(peek);
      }
  }
  pf->predict(delta[0], delta[1], delta[2], delta[3], delta[4], delta[5], noise_dev[0], noise_dev[1], noise_dev[2],
              noise_dev[3], noise_dev[4], noise_dev[5]);
}
// ---- MonteCarloLocalization (amcl3d.cpp): the direct callers of the path ---------------------------------------
void* ref_mcl_create(void* grid, const float* keyposes_xyz, uint64_t nkey)
{
  amcl3d::MonteCarloLocalization* m = new amcl3d::MonteCarloLocalization(g_logger);
  if (grid)
  {
    delete m->grid3d_;
    m->grid3d_ = static_cast<Grid3d*>(grid);
  }
  m->grid3d_->keyposes_3d = make_cloud(keyposes_xyz, 3, nkey);
  m->kdtree_keypose_3d_.reset(new pcl::KdTreeFLANN<PointType>());
  m->kdtree_keypose_3d_->setInputCloud(m->grid3d_->keyposes_3d);
  return m;
}
void ref_mcl_destroy(void* m)
{
  delete static_cast<amcl3d::MonteCarloLocalization*>(m);
}
void ref_mcl_set(void* mv, const float* aos7, uint64_t n)
{
  auto* m = static_cast<amcl3d::MonteCarloLocalization*>(mv);
  m->p_.resize(n);
  if (n)
    std::memcpy(m->p_.data(), aos7, n * sizeof(Particle));
}
uint64_t ref_mcl_size(void* mv)
{
  return static_cast<amcl3d::MonteCarloLocalization*>(mv)->p_.size();
}
void ref_mcl_get(void* mv, float* aos7)
{
  auto* m = static_cast<amcl3d::MonteCarloLocalization*>(mv);
  if (!m->p_.empty())
    std::memcpy(aos7, m->p_.data(), m->p_.size() * sizeof(Particle));
}
void ref_mcl_set_particle_num(void* mv, int max_sample, int min_sample)
{
  static_cast<amcl3d::MonteCarloLocalization*>(mv)->setParticleNum(max_sample, min_sample);
}
void ref_mcl_add_particle_weight(void* mv)
{
  static_cast<amcl3d::MonteCarloLocalization*>(mv)->addParticleWeight();
}
int ref_mcl_compute_sample_num(void* mv, int cnt_per_grid)
{
  return static_cast<amcl3d::MonteCarloLocalization*>(mv)->computeSampleNum(cnt_per_grid);
}
void ref_mcl_update_sample_height(void* mv)
{
  static_cast<amcl3d::MonteCarloLocalization*>(mv)->updateSampleHeight();
}
void ref_mcl_pf_resample(void* mv, int weight_multiply)
{
  static_cast<amcl3d::MonteCarloLocalization*>(mv)->PFResample(weight_multiply != 0);
}

// the input downsample exactly as amcl3d.cpp:37,51-52 drives it: the MonteCarloLocalization's own ds_ member
// (private, reached like the other members), setLeafSize(voxel_size x3), setInputCloud, filter.  The filter
// class itself is the stand-in of oracle/stubs (PCL is absent), so this pins the call pattern, not PCL.
uint64_t ref_mcl_downsample(void* mv, const float* xyzi, uint64_t n, double voxel_size, int is_dense, float* out4)
{
  auto* m = static_cast<amcl3d::MonteCarloLocalization*>(mv);
  pcl::PointCloud<PointType>::Ptr in(new pcl::PointCloud<PointType>());
  in->points.resize(n);
  for (uint64_t i = 0; i < n; ++i)
  {
    in->points[i].x = xyzi[4 * i];
    in->points[i].y = xyzi[4 * i + 1];
    in->points[i].z = xyzi[4 * i + 2];
    in->points[i].intensity = xyzi[4 * i + 3];
  }
  in->width = (uint32_t)n;
  in->is_dense = is_dense != 0;
  m->ds_.setLeafSize(voxel_size, voxel_size, voxel_size);
  m->ds_.setInputCloud(in);
  pcl::PointCloud<PointType> out;
  m->ds_.filter(out);
  for (size_t i = 0; i < out.points.size(); ++i)
  {
    out4[4 * i] = out.points[i].x;
    out4[4 * i + 1] = out.points[i].y;
    out4[4 * i + 2] = out.points[i].z;
    out4[4 * i + 3] = out.points[i].intensity;
  }
  return out.points.size();
}

// which trig overload did `const auto sr = sin(roll)` (Grid3d.cpp:240) bind to in THIS build?
// Returns 8 for double, 4 for float.  Mirrors the same unqualified call with the same headers visible.
int ref_trig_width()
{
  const float roll = 0.5f;
  const auto sr = sin(roll);
  return (int)sizeof(sr);
}
}  // extern "C"
