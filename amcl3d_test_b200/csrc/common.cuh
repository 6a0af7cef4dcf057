// common.cuh -- internal types shared by the sm_100a kernels and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/amcl3d_b200.h"

namespace a3d
{

#ifdef __CUDACC__
// ---- TMA bulk copy + mbarrier helpers (SASS: UBLKCP / SYNCS) shared by the streaming kernels -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t done;
  do
  {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif  // __CUDACC__

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
  float band;         // 0.5 - max(eps): |frac(v*rinv_f) - 0.5| < band  <=>  the f32 voxel index is exact on that axis
  uint32_t size[3];
  uint64_t step_z;    // size_x * size_y in 64 bits (equals the reference's u32 value below 2^32 cells)
  uint64_t n_cells;
  const float* prob;
  int has_map;
  int has_grid;
  // Lossless dictionary-coded copy of the grid for the weighting kernel (device-private; the public
  // format stays the linear float grid).  An EDT-built field takes only a few hundred distinct float
  // values (one per attainable integer squared distance below the expf underflow), so every voxel is
  // stored as a 1- or 2-byte code into `lut` and the kernel adds lut[code] -- the very same f32 the
  // float grid holds, hence bit-identical weights -- while a 128-byte line now covers an 8x4x4 (u8) or
  // 4x4x4 (u16) voxel brick instead of 32 voxels of one x-row.
  int rep;                 // kRepFloat / kRepCode8 / kRepCode16
  const void* codes;       // bricked codes, 2 MB aligned
  const float* lut;        // lut_n floats
  uint32_t lut_n;
  uint32_t nbx, nby;       // 128^3-voxel blocks along x and y
};

enum
{
  kRepFloat = 0,
  kRepCode8 = 1,   // brick 8 x 4 x 4 voxels = 128 B
  kRepCode16 = 2   // brick 4 x 4 x 4 voxels = 128 B
};

// Two-level layout of the coded copy.  Level 1: 128-byte bricks (8x4x4 one-byte or 4x4x4 two-byte codes)
// so that one cache line is a compact 3-D neighbourhood.  Level 2: 128^3-voxel blocks stored contiguously
// (2 MB of u8 codes = exactly one large page, 4 MB of u16 codes): a scattered gather over a compact 3-D
// region then touches a handful of pages instead of one page per (y,z) row group -- measured on B200,
// random 4-byte gathers collapse from ~2.6e11/s to ~4e10/s once the touched address range outgrows the
// TLB reach, even when the touched lines themselves fit in L2.
template <int kRep>
__host__ __device__ __forceinline__ uint64_t coded_index(uint32_t ix, uint32_t iy, uint32_t iz, uint32_t nbx, uint32_t nby)
{
  const uint64_t block = ((uint64_t)(iz >> 7) * nby + (iy >> 7)) * nbx + (ix >> 7);
  if (kRep == kRepCode8)
  {
    const uint32_t brick = (((iz >> 2) & 31u) << 9) | (((iy >> 2) & 31u) << 4) | ((ix >> 3) & 15u);
    const uint32_t vox = ((iz & 3u) << 5) | ((iy & 3u) << 3) | (ix & 7u);
    return (block << 21) | (brick << 7) | vox;
  }
  const uint32_t brick = (((iz >> 2) & 31u) << 10) | (((iy >> 2) & 31u) << 5) | ((ix >> 2) & 31u);
  const uint32_t vox = ((iz & 3u) << 4) | ((iy & 3u) << 2) | (ix & 3u);
  return (block << 21) | (brick << 6) | vox;
}
// inverse of coded_index
template <int kRep>
__host__ __device__ __forceinline__ void coded_voxel(uint64_t t, uint32_t nbx, uint32_t nby, uint32_t& ix, uint32_t& iy, uint32_t& iz)
{
  const uint64_t block = t >> 21;
  const uint32_t within = (uint32_t)(t & ((1u << 21) - 1));
  const uint32_t Bx = (uint32_t)(block % nbx), By = (uint32_t)((block / nbx) % nby), Bz = (uint32_t)(block / ((uint64_t)nbx * nby));
  if (kRep == kRepCode8)
  {
    const uint32_t brick = within >> 7, vox = within & 127u;
    ix = (Bx << 7) | ((brick & 15u) << 3) | (vox & 7u);
    iy = (By << 7) | (((brick >> 4) & 31u) << 2) | ((vox >> 3) & 3u);
    iz = (Bz << 7) | ((brick >> 9) << 2) | (vox >> 5);
  }
  else
  {
    const uint32_t brick = within >> 6, vox = within & 63u;
    ix = (Bx << 7) | ((brick & 31u) << 2) | (vox & 3u);
    iy = (By << 7) | (((brick >> 5) & 31u) << 2) | ((vox >> 2) & 3u);
    iz = (Bz << 7) | ((brick >> 10) << 2) | (vox >> 4);
  }
}
constexpr uint32_t kMaxLutSmem = 8192;  // LUT entries staged in shared memory (32 KB)

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
// Large-set path state (nullptr / stage == nullptr: single-kernel path): the Morton-sorted cloud (w = index of the
// point in the caller's order), the per-particle poses of THIS call's particles (pose_floats(n) floats, SoA with
// pitch pose_pitch = n rounded up to the tile width, filled by launch_pose) and the staging buffer
// stage[(tile * npad + row) * stage_tile_width() + lane] of 1-byte codes, rows in the caller's order.
struct TwoPhase
{
  const float4* sorted;
  uint8_t* stage;
  const float* pose;
  uint32_t pose_pitch;
  cudaEvent_t ev_mid, ev_end;  // nullable: recorded between the sweep and the ordered sum / after the ordered sum
};
size_t pose_floats(uint32_t n);
bool two_phase_supported(const MapDev& map);  // u8 codes and a map the branch-free sweep arithmetic is valid for
cudaError_t launch_pose(const MapDev& map, const float* particles, uint32_t n, int trig_mode, float* pose, cudaStream_t stream);
cudaError_t launch_update(const MapDev& map, float* particles, uint32_t n, const float4* cloud4, uint32_t npts,
                          uint32_t npts_padded, const BeaconSet* beacons, float* wr_out, int trig_mode,
                          unsigned long long* counter, const TwoPhase* tp, cudaStream_t stream);
// Morton order of a strided device cloud: sorted (npts_padded float4, NaN padded)
cudaError_t launch_sort_cloud(const float* src, uint32_t stride, uint32_t npts, uint32_t npts_padded, float4* sorted,
                              void* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t sort_cloud_scratch_bytes(uint32_t npts);
uint32_t stage_tile_width();  // particles per block of the staging layout [particle tile][row][lane]
// packs a strided cloud (device memory) into cloud4 and pads with NaN points up to npts_padded
cudaError_t launch_pack_cloud(const float* src, uint32_t stride, uint32_t npts, float4* dst, uint32_t npts_padded,
                              cudaStream_t stream);
cudaError_t launch_gather_bench(const float* cells, uint64_t window, uint32_t blocks, uint32_t iters, int coherent,
                                float* sink, cudaStream_t stream);

// ---- input downsample (voxelfilter.cu) ----------------------------------------------------------
size_t voxel_filter_scratch_bytes(uint32_t n);
cudaError_t run_voxel_filter(const float* d_pts, uint32_t stride, int ioff, uint32_t n, float leaf, int skip_nonfinite,
                             float4* d_out, void* scratch, size_t scratch_bytes, int64_t* cells, uint32_t* launches,
                             cudaStream_t stream);

// ---- particle-set ops (pfops.cu) --------------------------------------------------------------
// serial f32 chain over w (stride 7 floats): prefix[i] (nullable) = sequential inclusive prefix starting
// from 0; positive_only: add only w > 0 (particleNormalize).  total[0] receives the final value.
// chain_mode: 0 exact scan over many CTAs, 1 serial chain, 2 exact scan on one CTA (same bits; per filter handle)
cudaError_t launch_chain_prefix(const float* particles, uint64_t n, int positive_only, float* prefix, float* total,
                                float* dense_scratch /* chain_scratch_floats(n), nullable */, int chain_mode,
                                cudaStream_t stream);
// thresholds samp_0 = 0.5f/S, samp_{k+1} = samp_k + 1.0f/S (sequential f32 chain)
cudaError_t launch_chain_thresholds(int S, float* thr, float* scratch /* chain_scratch_floats(S), nullable */, int chain_mode,
                                    cudaStream_t stream);
// w = total>0 ? w/total : 0 ; w = 0 outside the map
cudaError_t launch_normalize(const MapDev& map, float* particles, uint64_t n, const float* total, cudaStream_t stream);
// range fusion: w = alpha*wp/sum_p + (1-alpha)*wr/sum_r (sums: sequential f32 chains)
cudaError_t launch_fuse_range(float* particles, const float* wr, uint64_t n, const float* sum_p, const float* sum_r,
                              float alpha, cudaStream_t stream);
uint64_t chain_scratch_floats(uint64_t n);  // dense operands + per-tile bookkeeping of the many-CTA scan
uint32_t chain_extra_launches_take();       // kernels launched by the chain calls beyond one each, since the last take
cudaError_t launch_chain_sum_array(const float* v, uint64_t n, float* total, float* scratch, int chain_mode,
                                   cudaStream_t stream);
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
// the exact (sequential f32) mean of a block that continues the chains of the blocks before it: carry7 = their state
cudaError_t launch_mean_exact_carry(const float* particles, uint64_t n, const float* carry7, float* out7, cudaStream_t stream);
size_t reduce_scratch_bytes();
cudaError_t launch_predict(float* particles, uint64_t n, const double* delta6_dev, const float* noise6,
                           cudaStream_t stream);

// the chains above on a DENSE weight array (stride 1): used by the sharded path on the all-rank copy
cudaError_t launch_chain_dense(const float* v, uint64_t n, int positive_only, float* prefix, float* total,
                               float* scratch /* chain_scratch_floats(n) */, int chain_mode, cudaStream_t stream);

// ---- sharded particle set over peer memory (shard.cu) -------------------------------------------------
constexpr uint32_t kMaxShardWorld = 16;
// every rank's exchange region (same layout on all ranks), mapped into this process
struct ShardPeers
{
  char* base[kMaxShardWorld];
  uint32_t world;
  uint64_t off_p, p_bytes;  // particle blocks P[2], p_bytes each
  uint64_t off_w, off_wr;   // dense f32 [2][n_capacity]
  uint64_t off_inmap;       // u8 [2][n_capacity]
  uint64_t off_mean;        // f32 [2][kMaxShardWorld][8]
  uint64_t off_flags;       // u32 [kShardFlagKinds][kMaxShardWorld]
  uint64_t off_sums;        // f32 [2 steps][2 parity][kMaxShardWorld]: per-rank partial sums of the fast mode
  uint64_t off_idx;         // i32 [n_max]: global source index of every output of this rank (fast mode, pushed by the source)
};
constexpr uint32_t kShardFlagKinds = 8;  // 0 weights, 1 mean, 2 / 3 fast-mode sums, 4 fast-mode push done
cudaError_t launch_shard_publish(const MapDev& map, const ShardPeers& peers, const float* particles, const float* wr_local,
                                 uint32_t n, uint32_t lo, uint32_t dense_stride, uint32_t parity, uint32_t epoch, uint32_t rank,
                                 unsigned int* done_counter, cudaStream_t stream);
cudaError_t launch_shard_wait(const uint32_t* flags, uint32_t kind, uint32_t world, uint32_t epoch, uint32_t* error,
                              cudaStream_t stream);
cudaError_t launch_dense_normalize(float* w, const uint8_t* inmap, uint64_t n, const float* total, cudaStream_t stream);
cudaError_t launch_dense_fuse(float* w, const float* wr, uint64_t n, const float* sum_p, const float* sum_r, float alpha,
                              cudaStream_t stream);
cudaError_t launch_shard_search(const ShardPeers& peers, const float* prefix, uint64_t n, float u01, uint64_t m0, uint64_t m1,
                                uint32_t src_buf, float* out, int32_t* idx, cudaStream_t stream);
// fast (non-exact) mode: a per-rank scalar into every rank's sums[step][parity][rank], then flags[2 + step][rank] = epoch
cudaError_t launch_shard_publish_scalar(const ShardPeers& peers, const float* value, uint32_t step, uint32_t parity, uint32_t epoch,
                                        uint32_t rank, cudaStream_t stream);
// out2[0] = sum over ranks in rank order (sequential f32), out2[1] = sum over the ranks before `rank`
cudaError_t launch_shard_combine_scalar(const float* sums_all, uint32_t world, uint32_t rank, float* out2, cudaStream_t stream);
// source-side resampling: every output m whose threshold falls into this rank's stretch (base, base + local total] of the
// global prefix is found in the LOCAL prefix and the particle is STORED into the output block of the rank that owns m
cudaError_t launch_shard_push(const ShardPeers& peers, const float* particles, const float* prefix_local, uint32_t n_local,
                              uint64_t lo, uint64_t n_total, const float* base2 /* [1] = base */, float u01, uint32_t rank,
                              uint32_t dst_buf, uint32_t epoch, unsigned int* done_counter, cudaStream_t stream);
cudaError_t launch_shard_publish_mean(const ShardPeers& peers, const float* part7, uint32_t parity, uint32_t epoch, uint32_t rank,
                                      cudaStream_t stream);
cudaError_t launch_shard_combine_mean(const float* mean_all, uint32_t world, float* out7, cudaStream_t stream);

// ---- callers of the path (mcl.cu) ----------------------------------------------------------------
// ordered-uint keys of min (0..2) and max (3..5) of x,y,z over the particle set
cudaError_t launch_particle_bounds(const float* particles, uint64_t n, uint32_t* keys6, cudaStream_t stream);
// occupied (x,y,z,yaw) bins of computeSampleNum -> count[0]
cudaError_t launch_sample_bins(const float* particles, uint64_t n, float xmin, float ymin, float zmin, int xwidth, int ywidth,
                               int zwidth, int theta, uint32_t* bits, size_t bits_words, int* count, cudaStream_t stream);
cudaError_t launch_shift_z(float* particles, uint64_t n, float delta_h, cudaStream_t stream);

// ---- grid build (gridbuild.cu) ------------------------------------------------------------------
cudaError_t launch_bounds(const float* xyz, uint32_t stride, uint64_t n, float* out6 /*min3,max3*/,
                          cudaStream_t stream);
struct CodedGrid
{
  int rep;  // kRepFloat: nothing allocated
  void* codes;        // 2 MB aligned view into codes_alloc
  void* codes_alloc;  // what cudaFree takes
  float* lut;
  uint32_t lut_n, nbx, nby;
  uint64_t bytes;
};
// full build: EDT (+ optional kept distances) and fused Gaussian into prob (device, n_cells floats);
// coded (nullable) additionally receives the dictionary-coded bricked copy (caller frees codes / lut)
// fast_passes != 0: capped window passes when the exact EDT is not kept (same float grid bit for bit)
cudaError_t build_grid_edt(const MapDev& map, const float* d_xyz, uint32_t stride, uint64_t n, double sensor_dev,
                           int mode, int sub, float* prob, int32_t** keep_d2, CodedGrid* coded, int fast_passes,
                           cudaStream_t stream,
                           cudaEvent_t* ev /* nullable, 4 events: the kernels of the build run inside ev[0]..ev[1] and
                                              ev[2]..ev[3]; the fused build keeps the cudaMalloc / cudaFree of its GBs of
                                              workspace (host-side driver calls) between and after the two regions */,
                           char* err, size_t errlen);
cudaError_t build_grid_splat(const MapDev& map, const float* d_xyz, uint32_t stride, uint64_t n, double sensor_dev,
                             float* prob, cudaStream_t stream);

}  // namespace a3d
