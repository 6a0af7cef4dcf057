/*
 * amcl3d_oracle.h -- CPU restatement of the amcl3d particle-weighting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and there only as the checker / CPU baseline, never as the thing shipped.
 *
 * Every function restates one reference function operation by operation (mixed f32/f64
 * arithmetic, sequential f32 sums, u32 index wrap); the citation is to /root/reference.
 * Parity pin: the restatement is validated against oracle/_ref (the reference's own
 * Grid3d.cpp / ParticleFilter.cpp / PointCloudTools.cpp compiled where they lie against
 * stub third-party headers, see oracle/Makefile) in tests/test_oracle_vs_ref.py and against
 * the golden vectors in tests/golden/ that oracle/_ref generated.
 */
#ifndef AMCL3D_ORACLE_H
#define AMCL3D_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* amcl3d/src/ParticleFilter.h:32-48 -- 7 packed floats */
typedef struct
{
  float x, y, z, roll, pitch, yaw, w;
} orc_particle;

/* PointCloudInfo (amcl3d/src/PointCloudTools.h:56-70) + Grid3dInfo (:39-52) flattened */
typedef struct
{
  double min[3];      /* octo_min_{x,y,z} */
  double max[3];      /* octo_max_{x,y,z} */
  double resol;       /* octo_resol */
  double sensor_dev;  /* Grid3dInfo::sensor_dev */
  uint32_t size[3];   /* size_x, size_y, size_z */
  uint32_t step_y;    /* = size_x (u32) */
  uint32_t step_z;    /* = size_x * size_y (u32, wraps) */
  uint64_t n_cells;   /* grid.size() */
  const float* prob;  /* Grid3dCell::prob, x fastest; NULL == "grid_info_ not set" */
  int has_pc;         /* 0 == "pc_info_ not set" */
} orc_map;

/* rosinrange_msg/msg/RangePose.msg:1-6 -- the fields the range weight needs */
typedef struct
{
  double position[3];
  double range;
} orc_beacon;

enum
{
  ORC_TRIG_F64 = 0, /* ::sin(double) binds (GCC 5 / Ubuntu 16.04, the reference platform) */
  ORC_TRIG_F32 = 1  /* std::sin(float) binds (libstdc++ >= 6 with the <math.h> wrapper visible) */
};

enum
{
  ORC_GRID_QUIRK = 0,    /* computeGrid as written: dist is the SQUARED NN distance -> exp(-d^4 c2) */
  ORC_GRID_GAUSSIAN = 1, /* proper Gaussian exp(-d^2 c2) (labelled extension, not the reference) */
  ORC_GRID_SPLAT = 2     /* computeGrid2: 27-neighbour max-splat (what the fork's open() runs) */
};

int orc_num_threads(void);

/* ---- grid geometry / bounds -------------------------------------------------------------- */
/* computePointCloud, PointCloudTools.cpp:84-109 (pcl::getMinMax3D part); returns -1 for n<=1 */
int orc_compute_bounds(const float* xyz, uint32_t stride, uint64_t n, double mn[3], double mx[3]);
/* PointCloudTools.cpp:120-130 */
void orc_grid_geometry(orc_map* m);

/* ---- likelihood-field query -------------------------------------------------------------- */
/* Grid3d::isIntoMap, Grid3d.cpp:365-372 */
int orc_is_into_map(const orc_map* m, float x, float y, float z);
/* Grid3d::computeCloudWeight (7-arg), Grid3d.cpp:234-301.  n_in (optional) receives `n`. */
float orc_cloud_weight(const orc_map* m, const float* cloud, uint32_t stride, uint64_t npts, float tx, float ty,
                       float tz, float roll, float pitch, float yaw, int trig_mode, int* n_in);
/* ParticleFilter::update(cloud), ParticleFilter.cpp:114-139; threads<=1: the reference's serial loop */
void orc_update(const orc_map* m, orc_particle* p, uint64_t n, const float* cloud, uint32_t stride, uint64_t npts,
                int trig_mode, int threads);
/* counts evaluations that pass every bounds test (the ones that load a voxel) */
uint64_t orc_count_in_bounds(const orc_map* m, const orc_particle* p, uint64_t n, const float* cloud,
                             uint32_t stride, uint64_t npts, int trig_mode, int threads);

/* ---- normalise / resample / mean --------------------------------------------------------- */
/* ParticleFilter::particleNormalize, ParticleFilter.cpp:168-196; returns wtp */
float orc_normalize(const orc_map* m, orc_particle* p, uint64_t n);
/* MonteCarloLocalization::addParticleWeight, amcl3d.cpp:205-217 */
void orc_scale_by_max(orc_particle* p, uint64_t n);
/* ParticleFilter::resample, ParticleFilter.cpp:230-254 with r = factor*u01; idx (optional) gets
 * the source index of every output; the walk-off-the-end UB read (:244-248) is clamped to n-1. */
void orc_resample(const orc_particle* p, uint64_t n, float u01, orc_particle* out, int32_t* idx);
/* ParticleFilter::uniformSample, ParticleFilter.cpp:256-280 (calls particleNormalize on p first);
 * out has S entries, the tail stays all-zero; idx[k] = -1 for the tail; returns emitted count */
int orc_uniform_sample(const orc_map* m, orc_particle* p, uint64_t n, int S, orc_particle* out, int32_t* idx);
/* ParticleFilter::getMean, ParticleFilter.cpp:198-216 */
void orc_mean(const orc_particle* p, uint64_t n, orc_particle* mean);
/* ParticleFilter::getBestParticle, ParticleFilter.cpp:218-228 */
void orc_best(const orc_particle* p, uint64_t n, orc_particle* best);
/* MonteCarloLocalization::computeSampleNum, amcl3d.cpp:242-288 (with computeBoundary :222-241): occupied
 * 0.2 m x yaw bins times cnt_per_grid, clamped to [min_resamp, max_resamp].  The reference indexes its yaw
 * axis unchecked (`y < theta` typo, :268); a yaw bin >= theta (|yaw| > 1100 rad) is not counted here. */
int orc_compute_sample_num(const orc_particle* p, uint64_t n, int cnt_per_grid, int min_resamp, int max_resamp);
/* MonteCarloLocalization::updateSampleHeight + getMapHeight, amcl3d.cpp:295-321: z += (z of the keypose nearest
 * to (mean.x, mean.y, 0)) - mean.z.  keyposes: nk x 3 floats.  Returns delta_h. */
float orc_update_sample_height(orc_particle* p, uint64_t n, const float* keyposes, uint64_t nk);
/* MonteCarloLocalization::PFResample, amcl3d.cpp:192-203; out must hold max_resamp particles; returns new size */
int orc_pf_resample_step(const orc_map* m, orc_particle* p, uint64_t n, int weight_multiply, int min_resamp, int max_resamp,
                         const float* keyposes, uint64_t nk, orc_particle* out);
/* The input downsample in front of update(): ds_.setLeafSize(voxel_size x3) / ds_.filter(), amcl3d.cpp:37,51-52.
 * pcl::VoxelGrid is third-party and absent from /root/reference (PCL as shipped with ROS Kinetic, 1.7.2; same
 * arithmetic in 1.8): this restates the published applyFilter of filters/impl/voxel_grid.hpp for PointXYZI with
 * default settings -- f32 bounds, inverse leaf 1.0f/leaf, ijk = (int)(floorf(v*inv) - (float)min_b),
 * idx = ijk.(1, div_x, div_x*div_y), cells in increasing idx, centroid = f32 running sums / (float)count.
 * PCL orders the points of one cell by an unstable std::sort on idx alone; here they are taken in input order
 * (the one deterministic choice), so centroids can differ from a PCL run in the last ulp.  PARITY UNPINNED for
 * this row: no PCL here and no reference test covers it.
 * pts: n points, `stride` floats apart, x y z first, intensity at float `ioff` (ioff < 0: none, output 0).
 * skip_nonfinite = !is_dense.  out: 4 floats per cell (x y z intensity), capacity n.  Returns the cell count, or
 * -1 when PCL's "leaf size too small" guard trips (PCL then passes the input through unchanged). */
int64_t orc_voxel_filter(const float* pts, uint32_t stride, int ioff, uint64_t n, float leaf, int skip_nonfinite, float* out);
/* sequential f32 inclusive prefix (ParticleFilter.cpp:259-264) -- exposed for the scan tests */
void orc_prefix(const orc_particle* p, uint64_t n, float* out);

/* ---- host steps either side of the path (noise injected) --------------------- */
/* ParticleFilter::relocPose, ParticleFilter.cpp:31-89; noise = 6*(n-1) N(0,1) draws already scaled
 * by dev in draw order x,y,z,roll,pitch,yaw (i.e. what ranGaussian(0,dev) returned) */
void orc_reloc_pose(orc_particle* p, uint64_t n, const float init[6], const float dev[6], const float* noise,
                    orc_particle* mean);
/* ParticleFilter::predict, ParticleFilter.cpp:91-112; noise = 6*n ranGaussian results in draw order
 * x,y,z,roll,pitch,yaw per particle */
void orc_predict(orc_particle* p, uint64_t n, const double delta[6], const float* noise);

/* ---- grid builders ----------------------------------------------------------------------- */
/* computeGrid, PointCloudTools.cpp:111-174, with an exact 1-NN (bucket search) standing in for
 * pcl::KdTreeFLANN (FLANN 1.8.4 L2_Simple<float>: squared distance accumulated in f32 x,y,z order).
 * mode = ORC_GRID_QUIRK (reference) or ORC_GRID_GAUSSIAN.  prob must hold m->n_cells floats. */
int orc_compute_grid_nn(const orc_map* m, const float* xyz, uint32_t stride, uint64_t n, int mode, float* prob,
                        int threads);
/* computeGrid2, PointCloudTools.cpp:176-254 */
int orc_compute_grid2(const orc_map* m, const float* xyz, uint32_t stride, uint64_t n, float* prob);

/* Exact Euclidean distance transform on the lattice of pitch resol/sub (sub = 1 or 2).
 * Sites: map points snapped to that lattice, h = lrint((p-min)/(resol/sub)) per axis.
 * Queries: voxel corners i*resol (lattice coordinate sub*i).  d2[i] = min over sites of the squared
 * distance in units of (resol/sub)^2, INT32_MAX if there is no site.  Separable lower-envelope
 * (Felzenszwalb-Huttenlocher) in integer arithmetic. */
int orc_edt_exact(const orc_map* m, const float* xyz, uint32_t stride, uint64_t n, int sub, int32_t* d2,
                  int threads);
/* O(V*P) brute force of the same definition, for small grids (validates orc_edt_exact) */
int orc_edt_brute(const orc_map* m, const float* xyz, uint32_t stride, uint64_t n, int sub, int32_t* d2);
/* Gaussian of the EDT: dist = (float)(d2 * (resol/sub)^2); prob = c1*expf(-dist*dist*c2) (quirk) or
 * c1*expf(-dist*c2) (gaussian); 0 where d2 == INT32_MAX.  PointCloudTools.cpp:139-140,160-167 */
void orc_grid_from_edt(const orc_map* m, const int32_t* d2, int sub, int mode, float* prob);

/* ---- range-beacon weight (row R; NOT in the reference tree -- parity unpinned) ----------- */
/* wr = prod_b k1*exp(-k2*(range_b - |p - a_b|)^2), f32; 0 if nb == 0 */
float orc_range_weight(const orc_particle* p, const orc_beacon* b, uint32_t nb, float sigma);
/* w = alpha*wp/sum(wp) + (1-alpha)*wr/sum(wr), then the caller normalises */
void orc_fuse_range(orc_particle* p, uint64_t n, const float* wr, float alpha);

/* ---- .grid cache file, Grid3d.cpp:374-442 ------------------------------------------------ */
int orc_save_grid(const char* path, const orc_map* m);
/* reads header into size/sensor_dev; if prob != NULL reads the cells too. returns 0 ok, -1 open
 * failure, -2 sensor_dev mismatch (Grid3d.cpp:422), -3 short file */
int orc_load_grid(const char* path, double sensor_dev, uint32_t size[3], double* file_dev, float* prob,
                  uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
