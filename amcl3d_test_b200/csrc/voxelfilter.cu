// voxelfilter.cu -- the input downsample in front of update(): pcl::VoxelGrid<PointXYZI> centroid filter (sm_100a)
//
// Stands behind ds_.setLeafSize / ds_.filter      amcl3d/src/amcl3d.cpp:37,51-52
// (pcl::VoxelGrid::applyFilter, PCL filters/impl/voxel_grid.hpp: third-party, absent from the reference tree)
//
// Integer part (cell index of every point, the set of occupied cells, their order, the count per cell) is exact.
// The centroid of a cell is the f32 running sum of its points divided by (float)count; PCL takes the points of a
// cell in whatever order its unstable std::sort left them, here they are taken in input order (stable radix sort
// on the cell index) -- the one deterministic choice, and the one the CPU checker makes too.
//
//   minmax   f32 bounds of the finite points                (block reduce + ordered-int atomics)
//   keys     idx = ijk . (1, div_x, div_x*div_y), 0xFFFFFFFF for skipped points
//   sort     cub::DeviceRadixSort::SortPairs (idx, point)   stable
//   cells    head flags -> inclusive scan -> cell starts
//   centroid one thread per cell, sequential f32 sums in sorted order
#include "common.cuh"

#include <cub/cub.cuh>
#include <float.h>

namespace a3d
{
namespace
{
constexpr uint32_t kSkip = 0xFFFFFFFFu;

// order-preserving map f32 -> u32 so that min / max can use integer atomics
__device__ __forceinline__ uint32_t f2o(float f)
{
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ bool point_ok(const float* p, int skip_nonfinite)
{
  return !skip_nonfinite || (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]));
}

__global__ void __launch_bounds__(256)
    cloud_minmax_kernel(const float* __restrict__ pts, uint32_t stride, uint32_t n, int skip_nonfinite,
                        uint32_t* __restrict__ mm /* [0..2] min, [3..5] max, [6] valid count */)
{
  uint32_t lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0u, 0u, 0u }, cnt = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const float* p = pts + (uint64_t)i * stride;
    if (!point_ok(p, skip_nonfinite))
      continue;
    ++cnt;
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      const uint32_t o = f2o(p[a]);
      lo[a] = min(lo[a], o);
      hi[a] = max(hi[a], o);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt)
  {
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      atomicMin(mm + a, lo[a]);
      atomicMax(mm + 3 + a, hi[a]);
    }
    atomicAdd(mm + 6, cnt);
  }
}

struct VoxelGeom
{
  float inv;     // inverse_leaf_size_
  int min_b[3];  // floor(min * inv)
  int mul1, mul2;
};

__global__ void voxel_key_kernel(const float* __restrict__ pts, uint32_t stride, uint32_t n, int skip_nonfinite,
                                 VoxelGeom g, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float* p = pts + (uint64_t)i * stride;
  uint32_t key = kSkip;
  if (point_ok(p, skip_nonfinite))
  {
    const int ijk0 = (int)__fsub_rn(floorf(__fmul_rn(p[0], g.inv)), (float)g.min_b[0]);
    const int ijk1 = (int)__fsub_rn(floorf(__fmul_rn(p[1], g.inv)), (float)g.min_b[1]);
    const int ijk2 = (int)__fsub_rn(floorf(__fmul_rn(p[2], g.inv)), (float)g.min_b[2]);
    key = (uint32_t)(ijk0 + ijk1 * g.mul1 + ijk2 * g.mul2);
  }
  keys[i] = key;
  vals[i] = i;
}

__global__ void voxel_head_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ flags)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t k = keys[i];
  flags[i] = (k != kSkip && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
}

// ids = inclusive scan of the head flags: the cell of sorted element i is ids[i] - 1
__global__ void voxel_start_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ flags,
                                   const uint32_t* __restrict__ ids, uint32_t n, uint32_t* __restrict__ cell_start,
                                   uint32_t* __restrict__ counts /* [0] cells, [1] valid points */)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  if (flags[i])
    cell_start[ids[i] - 1] = i;
  if (keys[i] != kSkip && (i + 1 == n || keys[i + 1] == kSkip))
  {
    counts[0] = ids[i];
    counts[1] = i + 1;
    cell_start[ids[i]] = i + 1;
  }
}

__global__ void voxel_centroid_kernel(const float* __restrict__ pts, uint32_t stride, int ioff,
                                      const uint32_t* __restrict__ vals, const uint32_t* __restrict__ cell_start,
                                      const uint32_t* __restrict__ counts, float4* __restrict__ out)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= counts[0])
    return;
  const uint32_t i0 = cell_start[c], i1 = cell_start[c + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  for (uint32_t i = i0; i < i1; ++i)
  {
    const float* p = pts + (uint64_t)vals[i] * stride;
    sx = __fadd_rn(sx, p[0]);
    sy = __fadd_rn(sy, p[1]);
    sz = __fadd_rn(sz, p[2]);
    if (ioff >= 0)
      si = __fadd_rn(si, p[ioff]);
  }
  const float cnt = (float)(i1 - i0);
  out[c] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), __fdiv_rn(si, cnt));
}

inline float o2f(uint32_t o)
{
  const uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  float f;
  memcpy(&f, &b, 4);
  return f;
}
}  // namespace

size_t voxel_filter_scratch_bytes(uint32_t n)
{
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
  const size_t cub_bytes = ((sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 255) / 256 * 256;
  // keys x2, vals x2, flags, ids, cell_start (n+1) + 64 words of scalars, then the CUB workspace
  return (size_t)(7 * (size_t)n + 64 + 64) * 4 + 256 + cub_bytes;
}

// Returns cudaSuccess and *cells = number of output points (out: 4 floats each, capacity n), or *cells = -1 when
// PCL's "leaf size is too small" guard trips (the caller passes the input through unchanged, as PCL does).
// Synchronises the stream twice (bounds, count): both decide host-side control flow.
cudaError_t run_voxel_filter(const float* d_pts, uint32_t stride, int ioff, uint32_t n, float leaf, int skip_nonfinite,
                             float4* d_out, void* scratch, size_t scratch_bytes, int64_t* cells, uint32_t* launches,
                             cudaStream_t stream)
{
  *cells = 0;
  if (n == 0)
    return cudaSuccess;
  uint32_t* w = reinterpret_cast<uint32_t*>(scratch);
  uint32_t* scal = w;  // [0..6] minmax + count, [8..9] cells / valid
  uint32_t* keys = w + 64;
  uint32_t* keys2 = keys + n;
  uint32_t* vals = keys2 + n;
  uint32_t* vals2 = vals + n;
  uint32_t* flags = vals2 + n;
  uint32_t* ids = flags + n;
  uint32_t* cell_start = ids + n;
  char* cub_ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(cell_start + n + 64) + 255) / 256 * 256);
  size_t cub_bytes = scratch_bytes - (size_t)(cub_ws - reinterpret_cast<char*>(scratch));

  const uint32_t init[16] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u };
  cudaError_t e = cudaMemcpyAsync(scal, init, sizeof(init), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess)
    return e;
  const unsigned blocks = (n + 255) / 256;
  cloud_minmax_kernel<<<blocks < 592u ? blocks : 592u, 256, 0, stream>>>(d_pts, stride, n, skip_nonfinite, scal);
  uint32_t h[8];
  e = cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess)
    return e;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess)
    return e;
  *launches += 1;
  if (h[6] == 0)
    return cudaSuccess;  // no finite point
  // the same f32 host arithmetic as voxel_grid.hpp: bounds, guard, min_b / div_b / divb_mul
  const float inv = 1.0f / leaf;
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a)
  {
    lo[a] = o2f(h[a]);
    hi[a] = o2f(h[3 + a]);
  }
  const int64_t dx = (int64_t)((hi[0] - lo[0]) * inv) + 1;
  const int64_t dy = (int64_t)((hi[1] - lo[1]) * inv) + 1;
  const int64_t dz = (int64_t)((hi[2] - lo[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX)
  {
    *cells = -1;
    return cudaSuccess;
  }
  VoxelGeom g;
  g.inv = inv;
  int div_b[3];
  for (int a = 0; a < 3; ++a)
  {
    g.min_b[a] = (int)floorf(lo[a] * inv);
    div_b[a] = (int)floorf(hi[a] * inv) - g.min_b[a] + 1;
  }
  g.mul1 = div_b[0];
  g.mul2 = div_b[0] * div_b[1];
  voxel_key_kernel<<<blocks, 256, 0, stream>>>(d_pts, stride, n, skip_nonfinite, g, keys, vals);
  e = cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys, keys2, vals, vals2, (int)n, 0, 32, stream);
  if (e != cudaSuccess)
    return e;
  voxel_head_kernel<<<blocks, 256, 0, stream>>>(keys2, n, flags);
  e = cub::DeviceScan::InclusiveSum(cub_ws, cub_bytes, flags, ids, (int)n, stream);
  if (e != cudaSuccess)
    return e;
  voxel_start_kernel<<<blocks, 256, 0, stream>>>(keys2, flags, ids, n, cell_start, scal + 8);
  // at most one cell per valid point: launch for the bound, the kernel reads the real count
  voxel_centroid_kernel<<<(h[6] + 127) / 128, 128, 0, stream>>>(d_pts, stride, ioff, vals2, cell_start, scal + 8, d_out);
  e = cudaMemcpyAsync(h, scal + 8, 8, cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess)
    return e;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess)
    return e;
  *launches += 6;  // keys, sort (counted once), heads, scan, starts, centroids
  *cells = (int64_t)h[0];
  return cudaGetLastError();
}

}  // namespace a3d
