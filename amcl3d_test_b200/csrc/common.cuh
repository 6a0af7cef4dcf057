// common.cuh -- internal types shared by the sm_100a kernels and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/amcl3d_b200.h"

namespace a3d
{

// Device view of PointCloudInfo + Grid3dInfo (amcl3d/src/PointCloudTools.h:39-70) with the derived
// constants the kernels need so that f32-vs-f64 comparisons of the reference can be done exactly in f32.
struct MapDev
{
  double min[3];
  double max[3];
  double resol;
  double rinv;        // 1.0 / resol (f64, rounded)
  float size_up[3];   // smallest f32 >= (max - min): for f32 x,  x < (double)size  <=>  x < size_up
  float eps[3];       // half-width of the "too close to a voxel face for the f32 quotient" band
  float rinv_f;       // (float)(1.0 / resol)
  uint32_t size[3];
  uint64_t step_z;    // size_x * size_y in 64 bits (equals the reference's u32 value below 2^32 cells)
  uint64_t n_cells;
  const float* prob;
  int has_map;
  int has_grid;
};

constexpr int kMaxBeacons = 16;

struct BeaconSet
{
  float ax[kMaxBeacons], ay[kMaxBeacons], az[kMaxBeacons], r[kMaxBeacons];
  uint32_t n;
  float k1, k2;
};

// ---- update (update.cu) ---------------------------------------------------------------------
// particles: AoS 7 floats; cloud4: (x,y,z,0) per point padded with NaN points to a multiple of 32.
// wr_out (nullable) receives the raw range-beacon weight per particle; counter (nullable) the number
// of evaluations that loaded a voxel.
cudaError_t launch_update(const MapDev& map, float* particles, uint32_t n, const float4* cloud4, uint32_t npts,
                          uint32_t npts_padded, const BeaconSet* beacons, float* wr_out, int trig_mode,
                          unsigned long long* counter, cudaStream_t stream);
// packs a strided cloud (device memory) into cloud4 and pads with NaN points up to npts_padded
cudaError_t launch_pack_cloud(const float* src, uint32_t stride, uint32_t npts, float4* dst, uint32_t npts_padded,
                              cudaStream_t stream);
cudaError_t launch_gather_bench(const float* cells, uint64_t window, uint32_t blocks, uint32_t iters, int coherent,
                                float* sink, cudaStream_t stream);

// ---- particle-set ops (pfops.cu) --------------------------------------------------------------
// serial f32 chain over w (stride 7 floats): prefix[i] (nullable) = sequential inclusive prefix starting
// from 0; positive_only: add only w > 0 (particleNormalize).  total[0] receives the final value.
cudaError_t launch_chain_prefix(const float* particles, uint64_t n, int positive_only, float* prefix, float* total,
                                cudaStream_t stream);
// thresholds samp_0 = 0.5f/S, samp_{k+1} = samp_k + 1.0f/S (sequential f32 chain)
cudaError_t launch_chain_thresholds(int S, float* thr, cudaStream_t stream);
// w = total>0 ? w/total : 0 ; w = 0 outside the map
cudaError_t launch_normalize(const MapDev& map, float* particles, uint64_t n, const float* total, cudaStream_t stream);
// range fusion: w = alpha*wp/sum_p + (1-alpha)*wr/sum_r (sums: sequential f32 chains)
cudaError_t launch_fuse_range(float* particles, const float* wr, uint64_t n, const float* sum_p, const float* sum_r,
                              float alpha, cudaStream_t stream);
cudaError_t launch_chain_sum_array(const float* v, uint64_t n, float* total, cudaStream_t stream);
// max weight (and first arg-max with w > 0) -> out_max[0], out_arg[0] (UINT64_MAX if none)
cudaError_t launch_max_weight(const float* particles, uint64_t n, float* out_max, unsigned long long* out_arg,
                              void* scratch, cudaStream_t stream);
cudaError_t launch_scale_by_max(float* particles, uint64_t n, const float* max_w, cudaStream_t stream);
// resample(): out[m-m0] = p[lower_bound(prefix, r + factor*m)], w = factor, for m in [m0, m1)
cudaError_t launch_resample_search(const float* particles, const float* prefix, uint64_t n, float u01, uint64_t m0,
                                   uint64_t m1, float* out, int32_t* idx, cudaStream_t stream);
// uniformSample(): out[k-k0] = p[upper_bound(prefix, thr[k])] or zeros, for k in [k0, k1); counts emitted
cudaError_t launch_uniform_search(const float* particles, const float* prefix, uint64_t n, const float* thr, int S,
                                  uint64_t k0, uint64_t k1, float* out, int32_t* idx, cudaStream_t stream);
cudaError_t launch_count_emitted(const float* thr, int S, const float* prefix, uint64_t n, int* out, cudaStream_t stream);
// mean: exact = sequential f32 chains (bit-exact), else f64 tree reduction. out7 device (7 floats)
cudaError_t launch_mean(const float* particles, uint64_t n, int exact, float* out7, void* scratch,
                        cudaStream_t stream);
size_t reduce_scratch_bytes();
cudaError_t launch_predict(float* particles, uint64_t n, const double* delta6_dev, const float* noise6,
                           cudaStream_t stream);

// ---- grid build (gridbuild.cu) ------------------------------------------------------------------
cudaError_t launch_bounds(const float* xyz, uint32_t stride, uint64_t n, float* out6 /*min3,max3*/,
                          cudaStream_t stream);
struct EdtResult
{
  int32_t* d2;  // device, size_x*size_y*size_z, caller frees with cudaFree (nullptr if !keep)
};
// full build: EDT (+ optional kept distances) and fused Gaussian into prob (device, n_cells floats)
cudaError_t build_grid_edt(const MapDev& map, const float* d_xyz, uint32_t stride, uint64_t n, double sensor_dev,
                           int mode, int sub, float* prob, int32_t** keep_d2, cudaStream_t stream,
                           char* err, size_t errlen);
cudaError_t build_grid_splat(const MapDev& map, const float* d_xyz, uint32_t stride, uint64_t n, double sensor_dev,
                             float* prob, cudaStream_t stream);

}  // namespace a3d
