/*
 * amcl3d_oracle.c -- CPU restatement of the amcl3d particle-weighting hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see amcl3d_oracle.h).  Plain C11; build with
 *   gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 * (no -march: the reference is built "-std=c++11 -O3" for baseline x86-64, amcl3d/CMakeLists.txt:9,
 * so every f32 operation is rounded individually and nothing is fused).
 *
 * All citations are to files under /root/reference.
 */
#define _GNU_SOURCE
#include "amcl3d_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static int clamp_threads(int threads)
{
  if (threads < 1)
    threads = 1;
  return threads;
}

/* ------------------------------------------------------------------------------------------ */
/* bounds + geometry                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* amcl3d/src/PointCloudTools.cpp:84-101: octo_size<=1 throws; pcl::getMinMax3D is a per-axis
 * f32 min/max over the finite points, stored into the f64 members. */
int orc_compute_bounds(const float* xyz, uint32_t stride, uint64_t n, double mn[3], double mx[3])
{
  if (!xyz || n <= 1)
    return -1;
  float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
  float hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  for (uint64_t i = 0; i < n; ++i)
  {
    const float* p = xyz + i * stride;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2]))
      continue;
    for (int a = 0; a < 3; ++a)
    {
      if (p[a] < lo[a])
        lo[a] = p[a];
      if (p[a] > hi[a])
        hi[a] = p[a];
    }
  }
  for (int a = 0; a < 3; ++a)
  {
    mn[a] = lo[a];
    mx[a] = hi[a];
  }
  return 0;
}

/* amcl3d/src/PointCloudTools.cpp:120-130: sizes by f64 ceil; steps and grid_size in u32 arithmetic */
void orc_grid_geometry(orc_map* m)
{
  for (int a = 0; a < 3; ++a)
  {
    const double octo_size = m->max[a] - m->min[a];
    m->size[a] = (uint32_t)ceil(octo_size / m->resol);
  }
  m->step_y = m->size[0];
  m->step_z = m->size[0] * m->size[1];
  const uint32_t grid_size = m->size[0] * m->size[1] * m->size[2]; /* u32 product, wraps like :130 */
  m->n_cells = grid_size;
}

/* ------------------------------------------------------------------------------------------ */
/* likelihood-field query                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* amcl3d/src/Grid3d.cpp:365-372: f32 arguments promoted against the f64 bounds */
int orc_is_into_map(const orc_map* m, float x, float y, float z)
{
  if (!m || !m->has_pc)
    return 0;
  return (x >= m->min[0] && x < m->max[0] && y >= m->min[1] && y < m->max[1] && z >= m->min[2] && z < m->max[2]);
}

/* amcl3d/src/Grid3d.cpp:240-250 */
static void rotation(float roll, float pitch, float yaw, int trig_mode, float r[9])
{
  if (trig_mode == ORC_TRIG_F64)
  {
    /* `const auto sr = sin(roll)` binding ::sin(double): f64 products, one rounding to f32 each */
    const double sr = sin((double)roll), cr = cos((double)roll);
    const double sp = sin((double)pitch), cp = cos((double)pitch);
    const double sy = sin((double)yaw), cy = cos((double)yaw);
    r[0] = (float)(cy * cp);
    r[1] = (float)(cy * sp * sr - sy * cr);
    r[2] = (float)(cy * sp * cr + sy * sr);
    r[3] = (float)(sy * cp);
    r[4] = (float)(sy * sp * sr + cy * cr);
    r[5] = (float)(sy * sp * cr - cy * sr);
    r[6] = (float)(-sp);
    r[7] = (float)(cp * sr);
    r[8] = (float)(cp * cr);
  }
  else
  {
    const float sr = sinf(roll), cr = cosf(roll);
    const float sp = sinf(pitch), cp = cosf(pitch);
    const float sy = sinf(yaw), cy = cosf(yaw);
    r[0] = cy * cp;
    r[1] = cy * sp * sr - sy * cr;
    r[2] = cy * sp * cr + sy * sr;
    r[3] = sy * cp;
    r[4] = sy * sp * sr + cy * cr;
    r[5] = sy * sp * cr - cy * sr;
    r[6] = -sp;
    r[7] = cp * sr;
    r[8] = cp * cr;
  }
}

/* amcl3d/src/Grid3d.cpp:234-301 */
float orc_cloud_weight(const orc_map* m, const float* cloud, uint32_t stride, uint64_t npts, float tx, float ty,
                       float tz, float roll, float pitch, float yaw, int trig_mode, int* n_in)
{
  if (n_in)
    *n_in = 0;
  if (!m || !m->prob || !m->has_pc) /* :237-238 */
    return 0;

  float r[9];
  rotation(roll, pitch, yaw, trig_mode, r);
  const float r00 = r[0], r01 = r[1], r02 = r[2], r10 = r[3], r11 = r[4], r12 = r[5], r20 = r[6], r21 = r[7],
              r22 = r[8];

  const double octo_size_x = m->max[0] - m->min[0]; /* :252-254 */
  const double octo_size_y = m->max[1] - m->min[1];
  const double octo_size_z = m->max[2] - m->min[2];

  const double offset_x = tx - m->min[0]; /* :256-258, f32 - f64 -> f64 */
  const double offset_y = ty - m->min[1];
  const double offset_z = tz - m->min[2];

  const float* grid_ptr = m->prob;
  const size_t grid_size = (size_t)m->n_cells;
  float weight = 0.f;
  int n = 0;
  const float error_z = 0;

  for (uint64_t k = 0; k < npts; ++k)
  {
    const float* point = cloud + k * stride;
    const float px = point[0], py = point[1], pz = point[2];
    /* :275-277 -- ((x*r00 + y*r01) + (z+0)*r02) in f32, then + f64 offset, rounded to the f32 member */
    const float nx = (float)(px * r00 + py * r01 + (pz + error_z) * r02 + offset_x);
    const float ny = (float)(px * r10 + py * r11 + (pz + error_z) * r12 + offset_y);
    const float nz = (float)(px * r20 + py * r21 + (pz + error_z) * r22 + offset_z);

    if (nx >= 0 && nx < octo_size_x && ny >= 0 && ny < octo_size_y && nz >= 0 && nz < octo_size_z) /* :279-280 */
    {
      const uint32_t ix = (uint32_t)floor(nx / m->resol); /* :282-284, f64 divide */
      const uint32_t iy = (uint32_t)floor(ny / m->resol);
      const uint32_t iz = (uint32_t)floor(nz / m->resol);
      if (ix < m->size[0] && iy < m->size[1] && iz < m->size[2]) /* :286 */
      {
        const uint32_t grid_index = ix + iy * m->step_y + iz * m->step_z; /* :288, u32 */
        if (grid_index < grid_size) /* :290 */
        {
          weight += grid_ptr[grid_index]; /* :292, sequential f32 */
          n += 1;
        }
      }
    }
  }
  if (n_in)
    *n_in = n;
  return (n < 10) ? 0 : weight / npts; /* :299, f32 / (float)size_t */
}

/* amcl3d/src/ParticleFilter.cpp:114-139 */
void orc_update(const orc_map* m, orc_particle* p, uint64_t n, const float* cloud, uint32_t stride, uint64_t npts,
                int trig_mode, int threads)
{
  threads = clamp_threads(threads);
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) if (threads > 1)
  for (int64_t i = 0; i < (int64_t)n; ++i)
  {
    const float tx = p[i].x, ty = p[i].y, tz = p[i].z;
    if (!orc_is_into_map(m, tx, ty, tz)) /* :126-131 */
    {
      p[i].w = 0;
      continue;
    }
    p[i].w = orc_cloud_weight(m, cloud, stride, npts, tx, ty, tz, p[i].roll, p[i].pitch, p[i].yaw, trig_mode, NULL);
  }
}

uint64_t orc_count_in_bounds(const orc_map* m, const orc_particle* p, uint64_t n, const float* cloud,
                             uint32_t stride, uint64_t npts, int trig_mode, int threads)
{
  threads = clamp_threads(threads);
  uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) reduction(+ : total) if (threads > 1)
  for (int64_t i = 0; i < (int64_t)n; ++i)
  {
    if (!orc_is_into_map(m, p[i].x, p[i].y, p[i].z))
      continue;
    int cnt = 0;
    (void)orc_cloud_weight(m, cloud, stride, npts, p[i].x, p[i].y, p[i].z, p[i].roll, p[i].pitch, p[i].yaw,
                           trig_mode, &cnt);
    total += (uint64_t)cnt;
  }
  return total;
}

/* ------------------------------------------------------------------------------------------ */
/* normalise / resample / mean                                                                */
/* ------------------------------------------------------------------------------------------ */

/* amcl3d/src/ParticleFilter.cpp:168-196 */
float orc_normalize(const orc_map* m, orc_particle* p, uint64_t n)
{
  float wtp = 0, w_max = 0;
  for (uint64_t i = 0; i < n; i++)
  {
    if (p[i].w > 0)
      wtp += p[i].w;
    if (p[i].w > w_max)
      w_max = p[i].w;
  }
  for (uint64_t i = 0; i < n; ++i)
  {
    if (wtp > 0)
      p[i].w /= wtp;
    else
      p[i].w = 0;
    if (!orc_is_into_map(m, p[i].x, p[i].y, p[i].z))
      p[i].w = 0;
  }
  return wtp;
}

/* amcl3d/src/amcl3d.cpp:205-217: k = 1.0/max_w is an f64 divide rounded to the f32 k */
void orc_scale_by_max(orc_particle* p, uint64_t n)
{
  float max_w = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (p[i].w > max_w)
      max_w = p[i].w;
  float k = 1;
  if (max_w > 0)
    k = 1.0 / max_w;
  for (uint64_t i = 0; i < n; ++i)
    p[i].w = k * p[i].w;
}

/* amcl3d/src/ParticleFilter.cpp:230-254.  `factor * m`: m is uint32_t -> converted to f32. */
void orc_resample(const orc_particle* p, uint64_t n, float u01, orc_particle* out, int32_t* idx)
{
  if (n == 0)
    return;
  const float factor = 1.f / n;
  const float r = factor * u01; /* r = factor * rngUniform(0, 1) */
  float c = p[0].w;
  float u;
  uint64_t i = 0;
  for (uint32_t mth = 0; mth < n; ++mth)
  {
    u = r + factor * mth;
    while (u > c)
    {
      if (++i >= n)
        break;
      c += p[i].w;
    }
    /* :248 reads p_[i] with i possibly == n (UB in the reference): clamp, documented in DESIGN.md */
    const uint64_t src = (i >= n) ? n - 1 : i;
    if (i >= n)
      i = n; /* keep the walk saturated instead of running further past the end */
    if (out)
    {
      out[mth] = p[src];
      out[mth].w = factor;
    }
    if (idx)
      idx[mth] = (int32_t)src;
  }
}

/* amcl3d/src/ParticleFilter.cpp:259-264 */
void orc_prefix(const orc_particle* p, uint64_t n, float* out)
{
  if (n == 0)
    return;
  out[0] = p[0].w;
  for (uint64_t i = 1; i < n; i++)
    out[i] = p[i].w + out[i - 1];
}

/* amcl3d/src/ParticleFilter.cpp:256-280 */
int orc_uniform_sample(const orc_map* m, orc_particle* p, uint64_t n, int S, orc_particle* out, int32_t* idx)
{
  orc_normalize(m, p, n); /* :258 */
  float* left_sum = (float*)malloc(sizeof(float) * (n ? n : 1));
  orc_prefix(p, n, left_sum);

  if (out)
    memset(out, 0, sizeof(orc_particle) * (size_t)(S > 0 ? S : 0)); /* std::vector<Particle> new_p(S): zeros */
  if (idx)
    for (int k = 0; k < S; ++k)
      idx[k] = -1;
  const float inteval = 1.0f / S;
  float samp_value = 0.5f / S;
  int k = 0;
  for (uint64_t i = 0; i < n; i++)
  {
    while (k < S && left_sum[i] > samp_value)
    {
      if (out)
      {
        out[k] = p[i];
        out[k].w = inteval;
      }
      if (idx)
        idx[k] = (int32_t)i;
      k++;
      samp_value += inteval;
    }
  }
  free(left_sum);
  return k;
}

/* amcl3d/src/amcl3d.cpp:222-241 */
static void compute_boundary(const orc_particle* p, uint64_t n, float* xmin, float* ymin, float* zmin, float* xmax, float* ymax,
                             float* zmax)
{
  *xmin = *xmax = *ymin = *ymax = *zmin = *zmax = 0.0;
  if (n == 0)
    return;
  *xmin = *xmax = p[0].x;
  *ymin = *ymax = p[0].y;
  *zmin = *zmax = p[0].z;
  for (uint64_t i = 0; i < n; ++i)
  {
    const orc_particle _p = p[i];
    *xmax = *xmax > _p.x ? *xmax : _p.x;
    *xmin = *xmin > _p.x ? _p.x : *xmin;
    *ymax = *ymax > _p.y ? *ymax : _p.y;
    *ymin = *ymin > _p.y ? _p.y : *ymin;
    *zmax = *zmax > _p.z ? *zmax : _p.z;
    *zmin = *zmin > _p.z ? _p.z : *zmin;
  }
}

/* amcl3d/src/amcl3d.cpp:242-288 */
int orc_compute_sample_num(const orc_particle* p, uint64_t n, int cnt_per_grid, int min_resamp, int max_resamp)
{
  float xyz_step = 0.2;
  float theta_step = 8.0;
  float xmin, xmax, ymin, ymax, zmin, zmax;
  compute_boundary(p, n, &xmin, &ymin, &zmin, &xmax, &ymax, &zmax);
  int xwidth = ceilf((xmax - xmin) / xyz_step) + 1; /* std::ceil(float) -> float, +1, truncated to int */
  int ywidth = ceilf((ymax - ymin) / xyz_step) + 1;
  int zwidth = ceilf((zmax - zmin) / xyz_step) + 1;
  int theta = ceil((M_PI * 2 / theta_step * 180)) + 1;
  if (xwidth * ywidth * zwidth == 0)
    return min_resamp;
  else if (xwidth * ywidth * zwidth > 1e3)
    return max_resamp;
  const size_t cells = (size_t)xwidth * ywidth * zwidth * theta;
  unsigned char* grid = (unsigned char*)calloc(cells, 1);
  for (uint64_t i = 0; i < n; ++i)
  {
    const orc_particle _p = p[i];
    int x = (_p.x - xmin) / xyz_step;
    int y = (_p.y - ymin) / xyz_step;
    int z = (_p.z - zmin) / xyz_step;
    int t = (_p.yaw - (-M_PI)) / theta_step;
    if (x >= 0 && x < xwidth && y >= 0 && y < ywidth && z >= 0 && z < zwidth && t >= 0 && y < theta)
    {
      if (t < theta) /* the reference writes out of bounds here (UB); not counted */
        grid[(((size_t)x * ywidth + y) * zwidth + z) * theta + t] = 1;
    }
  }
  int cnt = 0;
  for (size_t c = 0; c < cells; ++c)
    cnt += grid[c] ? 1 : 0;
  free(grid);
  int sample_num = cnt * cnt_per_grid;
  if (sample_num < min_resamp)
    return min_resamp;
  if (sample_num > max_resamp)
    return max_resamp;
  return sample_num;
}

/* amcl3d/src/amcl3d.cpp:295-321; the kd-tree query is an exact 1-NN with FLANN L2_Simple<float> distances */
float orc_update_sample_height(orc_particle* p, uint64_t n, const float* keyposes, uint64_t nk)
{
  orc_particle pt;
  orc_mean(p, n, &pt);
  const float q[3] = { pt.x, pt.y, 0.f };
  float best = FLT_MAX;
  uint64_t bi = 0;
  for (uint64_t k = 0; k < nk; ++k)
  {
    float result = 0.f, diff;
    diff = q[0] - keyposes[3 * k];
    result += diff * diff;
    diff = q[1] - keyposes[3 * k + 1];
    result += diff * diff;
    diff = q[2] - keyposes[3 * k + 2];
    result += diff * diff;
    if (result < best)
    {
      best = result;
      bi = k;
    }
  }
  const float delta_h = keyposes[3 * bi + 2] - pt.z;
  for (uint64_t i = 0; i < n; ++i)
    p[i].z += delta_h;
  return delta_h;
}

/* amcl3d/src/amcl3d.cpp:192-203; computeSampleNum() is called with its default cnt_per_grid = 3 (amcl3d.h) */
int orc_pf_resample_step(const orc_map* m, orc_particle* p, uint64_t n, int weight_multiply, int min_resamp, int max_resamp,
                         const float* keyposes, uint64_t nk, orc_particle* out)
{
  if (weight_multiply)
    orc_scale_by_max(p, n);
  int sample_num = orc_compute_sample_num(p, n, 3, min_resamp, max_resamp);
  orc_uniform_sample(m, p, n, sample_num, out, NULL);
  orc_update_sample_height(out, (uint64_t)sample_num, keyposes, nk);
  return sample_num;
}

/* amcl3d/src/amcl3d.cpp:37,51-52 -> pcl::VoxelGrid<PointXYZI>::applyFilter (PCL filters/impl/voxel_grid.hpp) */
static int cmp_u64(const void* a, const void* b)
{
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x > y) - (x < y);
}
int64_t orc_voxel_filter(const float* pts, uint32_t stride, int ioff, uint64_t n, float leaf, int skip_nonfinite, float* out)
{
  const float inv = 1.0f / leaf; /* inverse_leaf_size_ = Array4f::Ones() / leaf_size_ */
  float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  uint64_t valid = 0;
  for (uint64_t i = 0; i < n; ++i) /* getMinMax3D */
  {
    const float* p = pts + i * stride;
    if (skip_nonfinite && (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])))
      continue;
    ++valid;
    for (int a = 0; a < 3; ++a)
    {
      if (p[a] < lo[a])
        lo[a] = p[a];
      if (p[a] > hi[a])
        hi[a] = p[a];
    }
  }
  if (valid == 0)
    return 0;
  const int64_t dx = (int64_t)((hi[0] - lo[0]) * inv) + 1;
  const int64_t dy = (int64_t)((hi[1] - lo[1]) * inv) + 1;
  const int64_t dz = (int64_t)((hi[2] - lo[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT_MAX)
    return -1;
  int min_b[3], div_b[3];
  for (int a = 0; a < 3; ++a)
  {
    min_b[a] = (int)floorf(lo[a] * inv);
    div_b[a] = (int)floorf(hi[a] * inv) - min_b[a] + 1;
  }
  const int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
  uint64_t* keyed = (uint64_t*)malloc(sizeof(uint64_t) * (valid ? valid : 1));
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    const float* p = pts + i * stride;
    if (skip_nonfinite && (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])))
      continue;
    const int ijk0 = (int)(floorf(p[0] * inv) - (float)min_b[0]);
    const int ijk1 = (int)(floorf(p[1] * inv) - (float)min_b[1]);
    const int ijk2 = (int)(floorf(p[2] * inv) - (float)min_b[2]);
    const int idx = ijk0 + ijk1 * mul1 + ijk2 * mul2;
    keyed[m++] = ((uint64_t)(unsigned int)idx << 32) | (uint64_t)(uint32_t)i; /* input order inside a cell */
  }
  qsort(keyed, m, sizeof(uint64_t), cmp_u64);
  int64_t cells = 0;
  uint64_t index = 0;
  while (index < m)
  {
    uint64_t i = index + 1;
    while (i < m && (keyed[i] >> 32) == (keyed[index] >> 32))
      ++i;
    float s[4] = { 0.f, 0.f, 0.f, 0.f };
    for (uint64_t li = index; li < i; ++li)
    {
      const float* p = pts + (keyed[li] & 0xFFFFFFFFu) * stride;
      s[0] += p[0];
      s[1] += p[1];
      s[2] += p[2];
      if (ioff >= 0)
        s[3] += p[ioff];
    }
    const float cnt = (float)(i - index);
    float* o = out + 4 * cells;
    o[0] = s[0] / cnt;
    o[1] = s[1] / cnt;
    o[2] = s[2] / cnt;
    o[3] = s[3] / cnt;
    ++cells;
    index = i;
  }
  free(keyed);
  return cells;
}

/* amcl3d/src/ParticleFilter.cpp:198-216 */
void orc_mean(const orc_particle* p, uint64_t n, orc_particle* mean)
{
  orc_particle mean_p = { 0, 0, 0, 0, 0, 0, 0 };
  float w_max = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    mean_p.x += p[i].w * p[i].x;
    mean_p.y += p[i].w * p[i].y;
    mean_p.z += p[i].w * p[i].z;
    mean_p.roll += p[i].w * p[i].roll;
    mean_p.pitch += p[i].w * p[i].pitch;
    mean_p.yaw += p[i].w * p[i].yaw;
    if (p[i].w > w_max)
      w_max = p[i].w;
  }
  mean_p.w = w_max;
  *mean = mean_p;
}

/* amcl3d/src/ParticleFilter.cpp:218-228: first strict arg-max with w > 0, else the zero particle */
void orc_best(const orc_particle* p, uint64_t n, orc_particle* best)
{
  orc_particle best_p = { 0, 0, 0, 0, 0, 0, 0 };
  for (uint64_t i = 0; i < n; ++i)
    if (p[i].w > best_p.w)
      best_p = p[i];
  *best = best_p;
}

/* ------------------------------------------------------------------------------------------ */
/* host steps with injected noise                                                             */
/* ------------------------------------------------------------------------------------------ */

/* amcl3d/src/ParticleFilter.cpp:31-89 */
void orc_reloc_pose(orc_particle* p, uint64_t n, const float init[6], const float devs[6], const float* noise,
                    orc_particle* mean)
{
  if (n == 0)
    return;
  const float x_dev = devs[0], y_dev = devs[1], z_dev = devs[2];
  const float dev = fmaxf(fmaxf(x_dev, y_dev), z_dev);            /* :40 */
  const float gauss_const_1 = 1. / (dev * sqrt(2 * M_PI));        /* :41, f64 then -> f32 */
  const float gauss_const_2 = 1. / (2 * dev * dev);               /* :42: 2*dev*dev in f32, 1./ in f64 */

  p[0].x = init[0];
  p[0].y = init[1];
  p[0].z = init[2];
  p[0].roll = init[3];
  p[0].pitch = init[4];
  p[0].yaw = init[5];
  p[0].w = gauss_const_1;

  float wt = p[0].w;
  float dist;
  for (uint64_t i = 1; i < n; ++i)
  {
    const float* g = noise + 6 * (i - 1);
    p[i].x = p[0].x + g[0];
    p[i].y = p[0].y + g[1];
    p[i].z = p[0].z + g[2];
    p[i].roll = p[0].roll + g[3];
    p[i].pitch = p[0].pitch + g[4];
    p[i].yaw = p[0].yaw + g[5];
    /* :64-65 sqrt(float) -> the f32 sqrt under either overload set once rounded to `float dist` */
    dist = sqrt((p[i].x - p[0].x) * (p[i].x - p[0].x) + (p[i].y - p[0].y) * (p[i].y - p[0].y) +
                (p[i].z - p[0].z) * (p[i].z - p[0].z));
    /* :67 exp(float) binding ::exp(double) on the reference platform, result rounded into f32 w */
    p[i].w = gauss_const_1 * exp(-dist * dist * gauss_const_2);
    wt += p[i].w;
  }
  orc_particle mean_p = { 0, 0, 0, 0, 0, 0, 0 };
  for (uint64_t i = 0; i < n; ++i)
  {
    p[i].w /= wt;
    mean_p.x += p[i].w * p[i].x;
    mean_p.y += p[i].w * p[i].y;
    mean_p.z += p[i].w * p[i].z;
    mean_p.roll += p[i].w * p[i].roll;
    mean_p.pitch += p[i].w * p[i].pitch;
    mean_p.yaw += p[i].w * p[i].yaw;
  }
  if (mean)
    *mean = mean_p;
}

/* amcl3d/src/ParticleFilter.cpp:91-112.  sa/ca are f32 variables holding sin/cos(f32 yaw) computed
 * through ::sin(double); rand_x = delta_x(f64) + noise(f32) rounded to the f32 rand_x. */
void orc_predict(orc_particle* p, uint64_t n, const double delta[6], const float* noise)
{
  float sa, ca, rand_x, rand_y;
  for (uint64_t i = 0; i < n; ++i)
  {
    const float* g = noise + 6 * i;
    sa = sin(p[i].yaw);
    ca = cos(p[i].yaw);
    rand_x = delta[0] + g[0];
    rand_y = delta[1] + g[1];
    p[i].x += ca * rand_x - sa * rand_y;
    p[i].y += sa * rand_x + ca * rand_y;
    p[i].z += delta[2] + g[2];
    p[i].roll += delta[3] + g[3];
    p[i].pitch += delta[4] + g[4];
    p[i].yaw += delta[5] + g[5];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* grid builders                                                                              */
/* ------------------------------------------------------------------------------------------ */

static void gauss_consts(double sensor_dev, float* c1, float* c2)
{
  *c1 = (float)(1. / (sensor_dev * sqrt(2 * M_PI)));    /* PointCloudTools.cpp:139 */
  *c2 = (float)(1. / (2. * sensor_dev * sensor_dev));   /* PointCloudTools.cpp:140 */
}

/* Exact 1-NN by bucket search, standing in for pcl::KdTreeFLANN::nearestKSearch(k=1).
 * FLANN 1.8.4 L2_Simple<float>: result += diff*diff over x,y,z in f32.  Returns the minimum of
 * those f32 squared distances over all points (the only thing computeGrid consumes). */
typedef struct
{
  const float* xyz;
  uint32_t stride;
  uint64_t n;
  float lo[3];
  float cell;
  int dim[3];
  uint32_t* start; /* dim product + 1 */
  uint32_t* items;
} nn_index;

static inline int nn_cell(const nn_index* ix, float v, int a)
{
  int c = (int)floorf((v - ix->lo[a]) / ix->cell);
  if (c < 0)
    c = 0;
  if (c >= ix->dim[a])
    c = ix->dim[a] - 1;
  return c;
}

static nn_index* nn_build(const float* xyz, uint32_t stride, uint64_t n, float cell)
{
  nn_index* ix = (nn_index*)calloc(1, sizeof(nn_index));
  ix->xyz = xyz;
  ix->stride = stride;
  ix->n = n;
  ix->cell = cell;
  float hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  ix->lo[0] = ix->lo[1] = ix->lo[2] = FLT_MAX;
  for (uint64_t i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a)
    {
      const float v = xyz[i * stride + a];
      if (v < ix->lo[a])
        ix->lo[a] = v;
      if (v > hi[a])
        hi[a] = v;
    }
  size_t cells = 1;
  for (int a = 0; a < 3; ++a)
  {
    ix->dim[a] = (int)floorf((hi[a] - ix->lo[a]) / cell) + 1;
    cells *= (size_t)ix->dim[a];
  }
  ix->start = (uint32_t*)calloc(cells + 1, sizeof(uint32_t));
  ix->items = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
  for (uint64_t i = 0; i < n; ++i)
  {
    const float* p = xyz + i * stride;
    const size_t c = ((size_t)nn_cell(ix, p[2], 2) * ix->dim[1] + nn_cell(ix, p[1], 1)) * ix->dim[0] + nn_cell(ix, p[0], 0);
    ix->start[c + 1]++;
  }
  for (size_t c = 0; c < cells; ++c)
    ix->start[c + 1] += ix->start[c];
  uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * (cells ? cells : 1));
  memcpy(fill, ix->start, sizeof(uint32_t) * cells);
  for (uint64_t i = 0; i < n; ++i)
  {
    const float* p = xyz + i * stride;
    const size_t c = ((size_t)nn_cell(ix, p[2], 2) * ix->dim[1] + nn_cell(ix, p[1], 1)) * ix->dim[0] + nn_cell(ix, p[0], 0);
    ix->items[fill[c]++] = (uint32_t)i;
  }
  free(fill);
  return ix;
}

static void nn_free(nn_index* ix)
{
  if (!ix)
    return;
  free(ix->start);
  free(ix->items);
  free(ix);
}

/* minimum f32 squared distance from q to any indexed point */
static float nn_query(const nn_index* ix, const float q[3])
{
  int c0[3];
  for (int a = 0; a < 3; ++a)
    c0[a] = nn_cell(ix, q[a], a);
  float best = FLT_MAX;
  const int max_ring = ix->dim[0] + ix->dim[1] + ix->dim[2];
  for (int ring = 0; ring <= max_ring; ++ring)
  {
    /* every point in a cell at Chebyshev ring distance >= ring is farther than (ring-1)*cell from q
     * (q may sit outside the indexed box, then the clamped cell only makes the bound conservative
     * when q is inside; for outside queries we simply never stop early on the first rings) */
    if (ring >= 2 && best < FLT_MAX)
    {
      const double lb = (double)(ring - 1) * ix->cell;
      int inside = 1;
      for (int a = 0; a < 3; ++a)
      {
        const float rel = (q[a] - ix->lo[a]) / ix->cell;
        if (rel < 0 || rel >= ix->dim[a])
          inside = 0;
      }
      if (inside && lb * lb > (double)best * (1.0 + 1e-5))
        break;
    }
    const int z0 = c0[2] - ring, z1 = c0[2] + ring;
    for (int cz = z0; cz <= z1; ++cz)
    {
      if (cz < 0 || cz >= ix->dim[2])
        continue;
      const int y0 = c0[1] - ring, y1 = c0[1] + ring;
      for (int cy = y0; cy <= y1; ++cy)
      {
        if (cy < 0 || cy >= ix->dim[1])
          continue;
        const int on_shell_zy = (cz == z0 || cz == z1 || cy == y0 || cy == y1);
        const int x0 = c0[0] - ring, x1 = c0[0] + ring;
        const int xstep = on_shell_zy ? 1 : (2 * ring > 0 ? 2 * ring : 1);
        for (int cx = x0; cx <= x1; cx += xstep)
        {
          if (cx < 0 || cx >= ix->dim[0])
            continue;
          const size_t c = ((size_t)cz * ix->dim[1] + cy) * ix->dim[0] + cx;
          for (uint32_t t = ix->start[c]; t < ix->start[c + 1]; ++t)
          {
            const float* p = ix->xyz + (size_t)ix->items[t] * ix->stride;
            float result = 0.f;
            float diff;
            diff = q[0] - p[0];
            result += diff * diff;
            diff = q[1] - p[1];
            result += diff * diff;
            diff = q[2] - p[2];
            result += diff * diff;
            if (result < best)
              best = result;
          }
        }
      }
    }
  }
  return best;
}

/* amcl3d/src/PointCloudTools.cpp:111-174 */
int orc_compute_grid_nn(const orc_map* m, const float* xyz, uint32_t stride, uint64_t n, int mode, float* prob,
                        int threads)
{
  if (!m || !xyz || n == 0 || !prob)
    return -1;
  threads = clamp_threads(threads);
  float gauss_const1, gauss_const2;
  gauss_consts(m->sensor_dev, &gauss_const1, &gauss_const2);
  nn_index* ix = nn_build(xyz, stride, n, (float)(m->resol * 4.0));
  const uint32_t sx = m->size[0], sy = m->size[1], sz = m->size[2];
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
  for (int64_t iz = 0; iz < (int64_t)sz; ++iz)
    for (uint32_t iy = 0; iy < sy; ++iy)
      for (uint32_t ixx = 0; ixx < sx; ++ixx)
      {
        float q[3];
        q[0] = m->min[0] + (ixx * m->resol); /* :152-154, f64 then stored into the f32 PointType */
        q[1] = m->min[1] + (iy * m->resol);
        q[2] = m->min[2] + ((uint32_t)iz * m->resol);
        const uint32_t index = ixx + iy * m->step_y + (uint32_t)iz * m->step_z;
        if (index >= m->n_cells)
          continue;
        const float dist = nn_query(ix, q); /* :160 -- the SQUARED distance */
        if (mode == ORC_GRID_GAUSSIAN)
          prob[index] = gauss_const1 * expf(-dist * gauss_const2);
        else
          prob[index] = gauss_const1 * expf(-dist * dist * gauss_const2); /* :162 */
      }
  nn_free(ix);
  return 0;
}

/* amcl3d/src/PointCloudTools.cpp:176-254 (the #else branch, TTTTT is never defined) */
int orc_compute_grid2(const orc_map* m, const float* xyz, uint32_t stride, uint64_t n, float* prob)
{
  if (!m || !xyz || !prob)
    return -1;
  const uint32_t grid_size = (uint32_t)m->n_cells;
  memset(prob, 0, sizeof(float) * (size_t)grid_size);
  float gauss_const1, gauss_const2;
  gauss_consts(m->sensor_dev, &gauss_const1, &gauss_const2);
  uint32_t index;
  float dist;
  float sp[3];
  for (uint64_t t = 0; t < n; ++t)
  {
    const float* pt = xyz + t * stride;
    const uint32_t ix = (uint32_t)((pt[0] - m->min[0]) / m->resol); /* :208-210 */
    const uint32_t iy = (uint32_t)((pt[1] - m->min[1]) / m->resol);
    const uint32_t iz = (uint32_t)((pt[2] - m->min[2]) / m->resol);
    int xx, yy, zz;
    for (int i = -1; i <= 1; i++)
    {
      xx = ix + i;
      /* :228 `xx >= grid_info->size_x` compares int with uint32_t: xx converts to unsigned, so a
       * negative xx is caught by either clause */
      if (xx < 0 || (uint32_t)xx >= m->size[0])
        continue;
      for (int j = -1; j <= 1; j++)
      {
        yy = iy + j;
        if (yy < 0 || (uint32_t)yy >= m->size[1])
          continue;
        for (int k = -1; k <= 1; k++)
        {
          zz = iz + k;
          if (zz < 0 || (uint32_t)zz > m->size[2]) /* :236 `>` as written */
            continue;
          index = xx + yy * m->step_y + zz * m->step_z;
          if (index >= grid_size) /* :238 */
            continue;
          sp[0] = m->min[0] + xx * m->resol; /* :239-241 f64 -> f32 member */
          sp[1] = m->min[1] + yy * m->resol;
          sp[2] = m->min[2] + zz * m->resol;
          dist = sqrtf((sp[0] - pt[0]) * (sp[0] - pt[0]) + (sp[1] - pt[1]) * (sp[1] - pt[1]) +
                       (sp[2] - pt[2]) * (sp[2] - pt[2])); /* :242-244 std::sqrt(float) */
          const float v = gauss_const1 * expf(-dist * dist * gauss_const2);
          prob[index] = (v > prob[index]) ? v : prob[index]; /* :246 */
        }
      }
    }
  }
  return 0;
}

/* ---- exact EDT ------------------------------------------------------------------------------ */

#define EDT_INF INT32_MAX

/* lattice coordinate of a map point: nearest node of the pitch resol/sub lattice */
static inline long snap(double v, double mn, double pitch)
{
  return lrint((v - mn) / pitch);
}

/* 1-D lower envelope.  Sites at lattice coordinate s_j = 2*j + par (in half-voxel units when sub == 2;
 * when sub == 1 coordinates are scaled by 1 and par == 0) carrying cost g[j]; queries at x_i = step*i.
 * out[i] = min_j g[j] + (x_i - s_j)^2.  All in int64.  v/z are caller-provided scratch. */
static void envelope_1d(const int64_t* g, int nj, int step, int par, int64_t* out, int ni, int* v, double* z)
{
  int k = -1;
  for (int j = 0; j < nj; ++j)
  {
    if (g[j] >= (int64_t)EDT_INF)
      continue;
    const double sj = (double)(step * j + par);
    double s = 0;
    while (k >= 0)
    {
      const double sk = (double)(step * v[k] + par);
      /* intersection of parabolas rooted at sk (g[v[k]]) and sj (g[j]); exact in double: all terms < 2^53 */
      s = ((double)g[j] + sj * sj - (double)g[v[k]] - sk * sk) / (2.0 * (sj - sk));
      if (s <= z[k])
        --k;
      else
        break;
    }
    if (k < 0)
    {
      k = 0;
      v[0] = j;
      z[0] = -1e300;
    }
    else
    {
      ++k;
      v[k] = j;
      z[k] = s;
    }
  }
  if (k < 0)
  {
    for (int i = 0; i < ni; ++i)
      out[i] = EDT_INF;
    return;
  }
  int t = 0;
  for (int i = 0; i < ni; ++i)
  {
    const double x = (double)(step * i);
    while (t < k && z[t + 1] < x)
      ++t;
    const int64_t d = (int64_t)(step * i) - (int64_t)(step * v[t] + par);
    /* the envelope picks a minimiser up to ties at the breakpoints; ties have equal value */
    int64_t best = g[v[t]] + d * d;
    /* guard against double rounding at breakpoints: also test the neighbours on the stack */
    if (t > 0)
    {
      const int64_t d2 = (int64_t)(step * i) - (int64_t)(step * v[t - 1] + par);
      const int64_t c = g[v[t - 1]] + d2 * d2;
      if (c < best)
        best = c;
    }
    if (t < k)
    {
      const int64_t d2 = (int64_t)(step * i) - (int64_t)(step * v[t + 1] + par);
      const int64_t c = g[v[t + 1]] + d2 * d2;
      if (c < best)
        best = c;
    }
    out[i] = best;
  }
}

int orc_edt_exact(const orc_map* m, const float* xyz, uint32_t stride, uint64_t n, int sub, int32_t* d2out,
                  int threads)
{
  if (!m || !xyz || !d2out || (sub != 1 && sub != 2))
    return -1;
  threads = clamp_threads(threads);
  const int sx = (int)m->size[0], sy = (int)m->size[1], sz = (int)m->size[2];
  /* site arrays have one extra node per axis: a point on the max face snaps to node `size` */
  const int wx = sx + 1, wy = sy + 1, wz = sz + 1;
  const size_t wcells = (size_t)wx * wy * wz;
  const size_t qcells = (size_t)sx * sy * sz;
  const double pitch = m->resol / sub;
  for (size_t i = 0; i < qcells; ++i)
    d2out[i] = EDT_INF;

  uint8_t* occ = (uint8_t*)malloc(wcells);
  int32_t* a = (int32_t*)malloc(sizeof(int32_t) * (size_t)sx * wy * wz); /* after x pass: [wz][wy][sx] */
  int32_t* b = (int32_t*)malloc(sizeof(int32_t) * (size_t)sx * sy * wz); /* after y pass: [wz][sy][sx] */
  if (!occ || !a || !b)
    return -2;

  const int nclass = (sub == 2) ? 8 : 1;
  for (int cls = 0; cls < nclass; ++cls)
  {
    const int px = cls & 1, py = (cls >> 1) & 1, pz = (cls >> 2) & 1;
    memset(occ, 0, wcells);
    uint64_t nsites = 0;
    for (uint64_t t = 0; t < n; ++t)
    {
      const float* p = xyz + t * stride;
      const long hx = snap(p[0], m->min[0], pitch), hy = snap(p[1], m->min[1], pitch), hz = snap(p[2], m->min[2], pitch);
      long jx, jy, jz;
      if (sub == 2)
      {
        if ((hx & 1) != px || (hy & 1) != py || (hz & 1) != pz)
          continue;
        jx = (hx - px) / 2;
        jy = (hy - py) / 2;
        jz = (hz - pz) / 2;
      }
      else
      {
        jx = hx;
        jy = hy;
        jz = hz;
      }
      if (jx < 0 || jx >= wx || jy < 0 || jy >= wy || jz < 0 || jz >= wz)
        continue;
      occ[((size_t)jz * wy + jy) * wx + jx] = 1;
      ++nsites;
    }
    if (nsites == 0)
      continue;
    const int step = sub; /* query lattice coordinate = sub * i; site = sub * j + parity */

    /* x pass: rows (jy, jz) of sites -> queries ix */
#pragma omp parallel num_threads(threads) if (threads > 1)
    {
      const int L = (wx > wy ? wx : wy) > wz ? (wx > wy ? wx : wy) : wz;
      int64_t* g = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);
      int64_t* o = (int64_t*)malloc(sizeof(int64_t) * (size_t)L);
      int* v = (int*)malloc(sizeof(int) * (size_t)L);
      double* z = (double*)malloc(sizeof(double) * ((size_t)L + 1));
#pragma omp for schedule(static)
      for (int64_t row = 0; row < (int64_t)wy * wz; ++row)
      {
        const uint8_t* r = occ + (size_t)row * wx;
        for (int j = 0; j < wx; ++j)
          g[j] = r[j] ? 0 : (int64_t)EDT_INF;
        envelope_1d(g, wx, step, px, o, sx, v, z);
        for (int i = 0; i < sx; ++i)
          a[(size_t)row * sx + i] = (int32_t)o[i];
      }
      /* y pass: columns (ix, jz) over jy -> queries iy */
#pragma omp for schedule(static)
      for (int64_t col = 0; col < (int64_t)sx * wz; ++col)
      {
        const int ixx = (int)(col % sx);
        const int jz = (int)(col / sx);
        for (int j = 0; j < wy; ++j)
          g[j] = a[((size_t)jz * wy + j) * sx + ixx];
        envelope_1d(g, wy, step, py, o, sy, v, z);
        for (int i = 0; i < sy; ++i)
          b[((size_t)jz * sy + i) * sx + ixx] = (int32_t)o[i];
      }
      /* z pass: columns (ix, iy) over jz -> queries iz, min-combined across parity classes */
#pragma omp for schedule(static)
      for (int64_t col = 0; col < (int64_t)sx * sy; ++col)
      {
        for (int j = 0; j < wz; ++j)
          g[j] = b[(size_t)j * sx * sy + (size_t)col];
        envelope_1d(g, wz, step, pz, o, sz, v, z);
        for (int i = 0; i < sz; ++i)
        {
          const size_t q = (size_t)i * sx * sy + (size_t)col;
          const int64_t val = o[i] >= (int64_t)EDT_INF ? (int64_t)EDT_INF : o[i];
          if (val < d2out[q])
            d2out[q] = (int32_t)val;
        }
      }
      free(g);
      free(o);
      free(v);
      free(z);
    }
  }
  free(occ);
  free(a);
  free(b);
  return 0;
}

int orc_edt_brute(const orc_map* m, const float* xyz, uint32_t stride, uint64_t n, int sub, int32_t* d2out)
{
  if (!m || !xyz || !d2out || (sub != 1 && sub != 2))
    return -1;
  const int sx = (int)m->size[0], sy = (int)m->size[1], sz = (int)m->size[2];
  const double pitch = m->resol / sub;
  long* h = (long*)malloc(sizeof(long) * 3 * (n ? n : 1));
  uint64_t ns = 0;
  for (uint64_t t = 0; t < n; ++t)
  {
    const float* p = xyz + t * stride;
    const long hx = snap(p[0], m->min[0], pitch), hy = snap(p[1], m->min[1], pitch), hz = snap(p[2], m->min[2], pitch);
    if (hx < 0 || hx > (long)sub * sx || hy < 0 || hy > (long)sub * sy || hz < 0 || hz > (long)sub * sz)
      continue;
    h[3 * ns] = hx;
    h[3 * ns + 1] = hy;
    h[3 * ns + 2] = hz;
    ++ns;
  }
  for (int iz = 0; iz < sz; ++iz)
    for (int iy = 0; iy < sy; ++iy)
      for (int ixx = 0; ixx < sx; ++ixx)
      {
        int64_t best = EDT_INF;
        for (uint64_t t = 0; t < ns; ++t)
        {
          const int64_t dx = (int64_t)sub * ixx - h[3 * t], dy = (int64_t)sub * iy - h[3 * t + 1],
                        dz = (int64_t)sub * iz - h[3 * t + 2];
          const int64_t d = dx * dx + dy * dy + dz * dz;
          if (d < best)
            best = d;
        }
        d2out[((size_t)iz * sy + iy) * sx + ixx] = (int32_t)best;
      }
  free(h);
  return 0;
}

void orc_grid_from_edt(const orc_map* m, const int32_t* d2, int sub, int mode, float* prob)
{
  float c1, c2;
  gauss_consts(m->sensor_dev, &c1, &c2);
  const double pitch = m->resol / sub;
  const double unit = pitch * pitch;
  const size_t cells = (size_t)m->size[0] * m->size[1] * m->size[2];
  for (size_t i = 0; i < cells; ++i)
  {
    if (d2[i] == EDT_INF)
    {
      prob[i] = 0.0f; /* PointCloudTools.cpp:164-168 */
      continue;
    }
    const float dist = (float)((double)d2[i] * unit); /* squared NN distance in m^2 */
    if (mode == ORC_GRID_GAUSSIAN)
      prob[i] = c1 * expf(-dist * c2);
    else
      prob[i] = c1 * expf(-dist * dist * c2); /* PointCloudTools.cpp:160-162 */
  }
}

/* ------------------------------------------------------------------------------------------ */
/* range-beacon weight (row R) -- NOT in /root/reference; upstream-lineage formula            */
/* ------------------------------------------------------------------------------------------ */

float orc_range_weight(const orc_particle* p, const orc_beacon* b, uint32_t nb, float sigma)
{
  if (nb == 0)
    return 0.f;
  float w = 1;
  const float k1 = 1.f / (sigma * sqrt(2 * M_PI));
  const float k2 = 0.5f / (sigma * sigma);
  for (uint32_t i = 0; i < nb; ++i)
  {
    const float ax = (float)b[i].position[0], ay = (float)b[i].position[1], az = (float)b[i].position[2];
    const float rr = (float)b[i].range;
    const float r = sqrtf((p->x - ax) * (p->x - ax) + (p->y - ay) * (p->y - ay) + (p->z - az) * (p->z - az));
    w *= k1 * expf(-k2 * (r - rr) * (r - rr));
  }
  return w;
}

void orc_fuse_range(orc_particle* p, uint64_t n, const float* wr, float alpha)
{
  float wtp = 0, wtr = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    wtp += p[i].w;
    wtr += wr[i];
  }
  for (uint64_t i = 0; i < n; ++i)
  {
    const float wp = wtp > 0 ? p[i].w / wtp : 0.f;
    const float wq = wtr > 0 ? wr[i] / wtr : 0.f;
    p[i].w = wp * alpha + wq * (1 - alpha);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* .grid cache file                                                                           */
/* ------------------------------------------------------------------------------------------ */

/* amcl3d/src/Grid3d.cpp:374-402: u32 size_x,size_y,size_z, f64 sensor_dev, then the cells */
int orc_save_grid(const char* path, const orc_map* m)
{
  if (!m || !m->prob)
    return -1;
  FILE* pf = fopen(path, "wb");
  if (!pf)
    return -1;
  fwrite(&m->size[0], sizeof(uint32_t), 1, pf);
  fwrite(&m->size[1], sizeof(uint32_t), 1, pf);
  fwrite(&m->size[2], sizeof(uint32_t), 1, pf);
  fwrite(&m->sensor_dev, sizeof(double), 1, pf);
  const uint32_t grid_size = m->size[0] * m->size[1] * m->size[2];
  fwrite(m->prob, sizeof(float), grid_size, pf);
  fclose(pf);
  return 0;
}

/* amcl3d/src/Grid3d.cpp:404-442 */
int orc_load_grid(const char* path, double sensor_dev, uint32_t size[3], double* file_dev, float* prob, uint64_t cap)
{
  FILE* pf = fopen(path, "rb");
  if (!pf)
    return -1;
  double dev = 0;
  if (fread(&size[0], sizeof(uint32_t), 1, pf) != 1 || fread(&size[1], sizeof(uint32_t), 1, pf) != 1 ||
      fread(&size[2], sizeof(uint32_t), 1, pf) != 1 || fread(&dev, sizeof(double), 1, pf) != 1)
  {
    fclose(pf);
    return -3;
  }
  if (file_dev)
    *file_dev = dev;
  if (fabs(dev - sensor_dev) >= DBL_EPSILON) /* :422 */
  {
    fclose(pf);
    return -2;
  }
  if (prob)
  {
    const uint32_t grid_size = size[0] * size[1] * size[2];
    if (grid_size > cap || fread(prob, sizeof(float), grid_size, pf) != grid_size)
    {
      fclose(pf);
      return -3;
    }
  }
  fclose(pf);
  return 0;
}
