// gridbuild.cu -- K1/K2: probability-grid build on the device (sm_100a)
//
// Stands behind computePointCloud   amcl3d/src/PointCloudTools.cpp:84-109   (bounds reduction)
//               computeGrid         amcl3d/src/PointCloudTools.cpp:111-174  (kd-tree NN + Gaussian -> exact EDT)
//               computeGrid2        amcl3d/src/PointCloudTools.cpp:176-254  (27-neighbour max splat)
//
// computeGrid asks, for every voxel corner min + i*resol, the squared distance to the nearest map
// point.  For map points on the lattice of pitch resol/sub (octomap-derived maps sit on voxel centres,
// i.e. the half-voxel lattice, sub = 2) that is an exact Euclidean distance transform in integer
// arithmetic: sites h = sub*j + parity, queries x = sub*i.  Sites are split by the parity of their
// coordinates (8 classes when sub = 2); every populated class is one voxel-resolution separable EDT
//   x pass : nearest set bit left / right in an occupancy bit-row            -> uint16 |dx|
//   y pass : lower envelope of parabolas per column (Meijster, integer Sep)  -> int32 dx^2+dy^2
//   z pass : the same along z, min-combined across classes, and in the last class fused with the
//            Gaussian writer  prob = c1*expf(-d^2*d^2*c2)  straight into the linear float grid.
// Classes with only a handful of sites (e.g. the two extreme-corner points that pin the bounds) are
// folded into that last epilogue by brute force instead of paying three passes.
// The result is bit-exact in lattice units regardless of tie-breaking; the passes are HBM streaming
// kernels (no reuse, no tensor-core work).
#include "common.cuh"

#include <float.h>
#include <stdio.h>

#include <vector>

namespace a3d
{
namespace
{
constexpr int32_t kInf = INT32_MAX;
constexpr uint16_t kInf16 = 0xFFFF;
constexpr int kSparseMax = 256;  // classes with <= this many points are handled by brute force

// ---- bounds ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2key(float f)
{
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// the point passes keep four points per thread in flight: a batch of coalesced loads, then the work (the passes are
// latency-bound otherwise: one point, three f64 divides and one atomic per loop trip).  A slot beyond n reads as NaN, which
// every pass skips like any other non-finite point.
constexpr int kPointsInFlight = 4;
__device__ __forceinline__ void load_points(const float* __restrict__ xyz, uint32_t stride, uint64_t n, uint64_t i0, uint64_t T,
                                            float (&q)[kPointsInFlight][3])
{
#pragma unroll
  for (int u = 0; u < kPointsInFlight; ++u)
  {
    const uint64_t i = i0 + (uint64_t)u * T;
    if (i < n)
    {
      const float* p = xyz + i * stride;
      q[u][0] = p[0];
      q[u][1] = p[1];
      q[u][2] = p[2];
    }
    else
      q[u][0] = q[u][1] = q[u][2] = __int_as_float(0x7FC00000);
  }
}
__global__ void bounds_kernel(const float* __restrict__ xyz, uint32_t stride, uint64_t n, uint32_t* __restrict__ keys)
{
  uint32_t lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0, 0, 0 };
  const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += kPointsInFlight * T)
  {
    float q[kPointsInFlight][3];
    load_points(xyz, stride, n, i0, T, q);  // all loads of the batch before the first use
#pragma unroll
    for (int u = 0; u < kPointsInFlight; ++u)
    {
      if (!isfinite(q[u][0]) || !isfinite(q[u][1]) || !isfinite(q[u][2]))
        continue;
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        const uint32_t k = f2key(q[u][a]);
        lo[a] = min(lo[a], k);
        hi[a] = max(hi[a], k);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0)
    {
      atomicMin(&keys[a], lo[a]);
      atomicMax(&keys[3 + a], hi[a]);
    }
  }
}

// ---- voxelisation ---------------------------------------------------------------------------------
struct Lattice
{
  double min[3];
  double pitch;  // resol / sub
  double rinv;   // fl64(1 / pitch)
  int sub;
  int w[3];      // site nodes per axis = size + 1
  uint32_t words_per_row;
};

__device__ __forceinline__ bool snap_point(const Lattice& L, const float* p, long long h[3])
{
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    if (!isfinite(p[a]))
      return false;
    // lrint((p - min) / pitch): the product with the rounded reciprocal is within |q| * 2^-51 of the quotient, so both
    // round to the same integer unless the quotient sits that close to a tie -- only then is the real divide needed
    const double t = __dsub_rn((double)p[a], L.min[a]);
    const double q = __dmul_rn(t, L.rinv);
    const double r = rint(q);
    if (0.5 - fabs(__dsub_rn(q, r)) <= fabs(q) * 1e-15)
      h[a] = __double2ll_rn(__ddiv_rn(t, L.pitch));
    else
      h[a] = __double2ll_rn(r);
  }
  return true;
}
__device__ __forceinline__ int class_of(const Lattice& L, const long long h[3])
{
  if (L.sub == 1)
    return 0;
  return (int)((h[0] & 1) | ((h[1] & 1) << 1) | ((h[2] & 1) << 2));
}

__global__ void class_count_kernel(const __grid_constant__ Lattice L, const float* __restrict__ xyz, uint32_t stride,
                                   uint64_t n, unsigned long long* __restrict__ counts)
{
  __shared__ unsigned int sh[8];
  if (threadIdx.x < 8)
    sh[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += kPointsInFlight * T)
  {
    float q[kPointsInFlight][3];
    load_points(xyz, stride, n, i0, T, q);
#pragma unroll
    for (int u = 0; u < kPointsInFlight; ++u)
    {
      long long h[3];
      if (!snap_point(L, q[u], h))
        continue;
      atomicAdd(&sh[class_of(L, h)], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 8 && sh[threadIdx.x])
    atomicAdd(&counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// dense class `cls`: set occupancy bits; sparse classes (mask bit set): append lattice coordinates
__device__ __forceinline__ void voxelize_point(const Lattice& L, const float* p, int cls, uint32_t* __restrict__ bits,
                                               uint32_t sparse_mask, int4* __restrict__ sparse, unsigned int* __restrict__ sparse_count)
{
  long long h[3];
  if (!snap_point(L, p, h))
    return;
  const int c = class_of(L, h);
  const bool is_sparse = sparse != nullptr && ((sparse_mask >> c) & 1u);
  if (!is_sparse && (bits == nullptr || c != cls))
    return;
  // site node index along each axis; nodes outside [0, size] are dropped
  long long j[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    j[a] = (L.sub == 2) ? ((h[a] - ((c >> a) & 1)) >> 1) : h[a];
  if (j[0] < 0 || j[0] >= L.w[0] || j[1] < 0 || j[1] >= L.w[1] || j[2] < 0 || j[2] >= L.w[2])
    return;
  if (is_sparse)
  {
    const unsigned int slot = atomicAdd(sparse_count, 1u);
    if (slot < 8u * kSparseMax)
      sparse[slot] = make_int4((int)h[0], (int)h[1], (int)h[2], 0);
    return;
  }
  const uint64_t row = (uint64_t)j[2] * L.w[1] + (uint64_t)j[1];
  atomicOr(&bits[row * L.words_per_row + (uint32_t)(j[0] >> 5)], 1u << (j[0] & 31));
}
__global__ void voxelize_kernel(const __grid_constant__ Lattice L, const float* __restrict__ xyz, uint32_t stride,
                                uint64_t n, int cls, uint32_t* __restrict__ bits, uint32_t sparse_mask,
                                int4* __restrict__ sparse, unsigned int* __restrict__ sparse_count)
{
  const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += kPointsInFlight * T)
  {
    float q[kPointsInFlight][3];
    load_points(xyz, stride, n, i0, T, q);
#pragma unroll
    for (int u = 0; u < kPointsInFlight; ++u)
      voxelize_point(L, q[u], cls, bits, sparse_mask, sparse, sparse_count);
  }
}

// ---- x pass ---------------------------------------------------------------------------------------
// one thread per (row, query ix); a warp shares one occupancy word.  d1 = |sub*(ix - j) - par| of the
// nearest site j in the row, kInf16 if the row is empty.
__global__ void __launch_bounds__(256)
    xpass_kernel(const uint32_t* __restrict__ bits, uint32_t words_per_row, uint64_t rows, int sx, int wx, int sub, int par,
                 int max_scan_words /* how many further words to look at on each side (capped passes: the window) */,
                 uint16_t* __restrict__ d1)
{
  const uint32_t chunks = (uint32_t)((sx + 31) / 32);
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows * chunks)
    return;
  const uint64_t row = warp / chunks;
  const uint32_t wq = (uint32_t)(warp % chunks);
  const uint32_t* r = bits + row * words_per_row;
  const uint32_t W = r[wq];

  // nearest site at or left of this lane
  int left = -1;
  {
    const uint32_t m = W & (0xFFFFFFFFu >> (31 - lane));
    if (m)
      left = (int)(wq * 32u) + 31 - __clz(m);
    else
    {
      const int w_stop = max(0, (int)wq - max_scan_words);
      for (int w = (int)wq - 1; w >= w_stop; --w)
      {
        const uint32_t v = r[w];
        if (v)
        {
          left = w * 32 + 31 - __clz(v);
          break;
        }
      }
    }
  }
  int right = -1;
  {
    const uint32_t m = W & (0xFFFFFFFFu << lane);
    if (m)
      right = (int)(wq * 32u) + __ffs(m) - 1;
    else
    {
      const uint32_t w_stop = (uint32_t)min((long long)words_per_row, (long long)wq + 1 + max_scan_words);
      for (uint32_t w = wq + 1; w < w_stop; ++w)
      {
        const uint32_t v = r[w];
        if (v)
        {
          right = (int)(w * 32u) + __ffs(v) - 1;
          break;
        }
      }
    }
  }
  const int ix = (int)(wq * 32u) + lane;
  if (ix >= sx)
    return;
  int best = 0x7FFFFFFF;
  if (left >= 0)
    best = abs(sub * (ix - left) - par);
  if (right >= 0 && right < wx)
    best = min(best, abs(sub * (ix - right) - par));
  d1[row * (uint64_t)sx + ix] = (best >= 0xFFFF) ? kInf16 : (uint16_t)best;
}

// ---- lower-envelope passes --------------------------------------------------------------------------
struct Gauss
{
  float c1, c2;
  double unit;  // (resol/sub)^2 : lattice units^2 -> m^2
  int mode;
};

__device__ __forceinline__ float prob_of(const Gauss& g, int32_t d2)
{
  if (d2 == kInf)
    return 0.f;  // PointCloudTools.cpp:164-168
  const float dist = __double2float_rn(__dmul_rn((double)d2, g.unit));  // squared NN distance in m^2
  if (g.mode == AMCL3D_B200_GRID_GAUSSIAN)
    return __fmul_rn(g.c1, expf(__fmul_rn(-dist, g.c2)));
  return __fmul_rn(g.c1, expf(__fmul_rn(__fmul_rn(-dist, dist), g.c2)));  // PointCloudTools.cpp:160-162
}

__device__ __forceinline__ long long floor_div(long long num, long long den)  // den > 0
{
  long long q = (long long)floor((double)num / (double)den);
  while (q * den > num)
    --q;
  while ((q + 1) * den <= num)
    ++q;
  return q;
}

struct PassArgs
{
  uint64_t ncol;     // contiguous (fast) extent
  uint32_t nplane;   // outer extent
  int len_in;        // sites along the pass axis (size + 1)
  int len_out;       // queries along the pass axis (size)
  int sub;
  int par;
  uint64_t nthreads; // total threads of the launch (stack stride)
};

struct FinalArgs
{
  int enabled;            // this is the last dense class: write prob
  const int32_t* prev;    // d2 of earlier classes (nullable)
  int32_t* d2_out;        // where to keep the combined d2 (nullable)
  float* prob;
  const int4* sparse;     // brute-force sites in lattice coordinates
  int nsparse;
  int sx, sy;             // to recover (ix, iy) from the column index in the z pass
  Gauss g;
};

__device__ __forceinline__ int32_t load_g(const uint16_t* in, uint64_t idx)
{
  const uint16_t v = in[idx];
  return v == kInf16 ? kInf : (int32_t)v * (int32_t)v;
}
__device__ __forceinline__ int32_t load_g(const int32_t* in, uint64_t idx)
{
  return in[idx];
}

// in : [plane][j][col]  (j < len_in),  out : [plane][i][col]  (i < len_out)
// out[i] = min_j in[j] + (sub*i - (sub*j + par))^2
template <typename TIn, bool kFinal>
__global__ void __launch_bounds__(256) envelope_kernel(const TIn* __restrict__ in, int32_t* __restrict__ out,
                                                       const __grid_constant__ PassArgs A, int2* __restrict__ stack,
                                                       const __grid_constant__ FinalArgs F)
{
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = A.ncol * A.nplane;
  for (uint64_t c = gtid; c < total; c += A.nthreads)
  {
    const uint64_t plane = c / A.ncol;
    const uint64_t col = c % A.ncol;
    const TIn* src = in + plane * (uint64_t)A.len_in * A.ncol + col;
    int q = -1;
    // cached top of stack
    int top_s = 0, top_t = 0, top_g = 0;
    for (int j = 0; j < A.len_in; ++j)
    {
      const int32_t g = load_g(src, (uint64_t)j * A.ncol);
      if (g == kInf)
        continue;
      const long long sj = (long long)A.sub * j + A.par;
      while (q >= 0)
      {
        const long long x = (long long)A.sub * top_t;
        const long long st = (long long)A.sub * top_s + A.par;
        const long long ft = (x - st) * (x - st) + top_g;
        const long long fj = (x - sj) * (x - sj) + g;
        if (ft > fj)
        {
          --q;
          if (q >= 0)
          {
            const int2 e = stack[(uint64_t)q * A.nthreads + gtid];
            top_s = e.x & 0xFFFF;
            top_t = (e.x >> 16) & 0xFFFF;
            top_g = e.y;
          }
        }
        else
          break;
      }
      if (q < 0)
      {
        q = 0;
        top_s = j;
        top_t = 0;
        top_g = g;
        stack[gtid] = make_int2(j, g);
      }
      else
      {
        const long long st = (long long)A.sub * top_s + A.par;
        const long long num = sj * sj - st * st + (long long)g - (long long)top_g;
        const long long den = 2ll * A.sub * (sj - st);
        const long long w = floor_div(num, den) + 1;
        if (w < A.len_out)
        {
          ++q;
          top_s = j;
          top_t = (int)w;
          top_g = g;
          stack[(uint64_t)q * A.nthreads + gtid] = make_int2(j | ((int)w << 16), g);
        }
      }
    }
    // backward scan
    int32_t* dst = out ? out + plane * (uint64_t)A.len_out * A.ncol + col : nullptr;
    for (int i = A.len_out - 1; i >= 0; --i)
    {
      int32_t d2 = kInf;
      if (q >= 0)
      {
        const long long d = (long long)A.sub * i - ((long long)A.sub * top_s + A.par);
        d2 = (int32_t)(d * d + top_g);
        if (i == top_t)
        {
          --q;
          if (q >= 0)
          {
            const int2 e = stack[(uint64_t)q * A.nthreads + gtid];
            top_s = e.x & 0xFFFF;
            top_t = (e.x >> 16) & 0xFFFF;
            top_g = e.y;
          }
        }
      }
      const uint64_t o = plane * (uint64_t)A.len_out * A.ncol + (uint64_t)i * A.ncol + col;
      if (kFinal)
      {
        if (F.prev != nullptr)
          d2 = min(d2, F.prev[o]);
        if (F.enabled)
        {
          if (F.nsparse > 0)
          {
            const long long qx = (long long)A.sub * (long long)(col % (uint64_t)F.sx);
            const long long qy = (long long)A.sub * (long long)(col / (uint64_t)F.sx);
            const long long qz = (long long)A.sub * i;
            for (int s = 0; s < F.nsparse; ++s)
            {
              const int4 h = F.sparse[s];
              const long long dx = qx - h.x, dy = qy - h.y, dz = qz - h.z;
              const long long dd = dx * dx + dy * dy + dz * dz;
              if (dd < (long long)d2)
                d2 = (int32_t)dd;
            }
          }
          F.prob[o] = prob_of(F.g, d2);
          if (F.d2_out != nullptr)
            F.d2_out[o] = d2;
        }
        else
          F.d2_out[o] = d2;
      }
      else
        dst[(uint64_t)i * A.ncol] = d2;
    }
  }
}

// ---- capped lower envelope by direct search -------------------------------------------------------------
// The field only needs min(d2, cap): cap = the smallest squared distance whose probability is exactly 0.0f, a few
// dozen lattice steps (34 half-voxels at sigma = 0.05 m).  Below cap the minimising site is within R = floor(sqrt(cap-1))
// lattice steps along every axis, so  out[i] = min(cap, min_{|sub*(i-j)-par| <= R} in[j] + (sub*i-(sub*j+par))^2)  equals
// min(exact, cap) -- no stacks, no divisions, every access coalesced, inputs staged once per tile in shared memory.
// Layout as envelope_kernel: in [plane][j][col], out [plane][i][col].  Used when the caller does not keep the EDT.
constexpr int kWinTX = 64;    // columns per block
constexpr int kWinTI = 32;    // outputs per column per block
constexpr int kWinPer = 8;    // outputs per thread (256 threads = 64 columns x 4 groups of 8)
constexpr int kWinMaxW = 44;  // widest window (sites on each side) the tile is sized for (31 KB of shared memory)
constexpr int32_t kBig = 1 << 28;

__device__ __forceinline__ int32_t win_load(const uint16_t* in, uint64_t idx, int32_t cap, bool squared)
{
  const int32_t v = (int32_t)in[idx];
  if (squared)
    return v == (int32_t)kInf16 ? kBig : v * v;  // x pass output: |dx|
  return v >= cap ? kBig : v;                    // y pass output: already min(., cap)
}

template <bool kFinal>
__global__ void __launch_bounds__(256)
    window_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out16, const __grid_constant__ PassArgs A,
                  int32_t cap, int W, const __grid_constant__ FinalArgs F)
{
  __shared__ int32_t tile[kWinTI + 2 * kWinMaxW + 2][kWinTX];
  const uint64_t col_tiles = (A.ncol + kWinTX - 1) / kWinTX;
  const uint64_t i_tiles = ((uint64_t)A.len_out + kWinTI - 1) / kWinTI;
  uint64_t b = blockIdx.x;
  const uint64_t ct = b % col_tiles;
  b /= col_tiles;
  const uint64_t it = b % i_tiles;
  const uint64_t plane = b / i_tiles;
  const int tx = threadIdx.x & (kWinTX - 1), ty = threadIdx.x / kWinTX;
  const uint64_t col = ct * kWinTX + tx;
  const int i0 = (int)it * kWinTI;
  const int j_lo = max(0, i0 - W - 1);
  const int j_hi = min(A.len_in - 1, i0 + kWinTI - 1 + W + 1);
  const int rows = j_hi - j_lo + 1;
  const uint16_t* src = in + plane * (uint64_t)A.len_in * A.ncol;
  for (int r = ty; r < rows; r += 256 / kWinTX)
    tile[r][tx] = col < A.ncol ? win_load(src, (uint64_t)(j_lo + r) * A.ncol + col, cap, !kFinal) : kBig;
  __syncthreads();
  if (col >= A.ncol)
    return;
  const int ib = i0 + ty * kWinPer;  // first output of this thread
  // candidate of site j for output ib + k:  v + (d + sub*k)^2 = (v + d^2) + k*(2*sub*d) + sub^2*k^2  with d the offset
  // of output ib; the last term does not depend on the site, so it is added after the search (2 ops per pair)
  int32_t best[kWinPer];
#pragma unroll
  for (int k = 0; k < kWinPer; ++k)
    best[k] = cap - A.sub * A.sub * k * k;
  // rows that can matter for outputs ib .. ib+kWinPer-1
  const int r0 = max(ib - W - 1, j_lo) - j_lo, r1 = min(ib + kWinPer - 1 + W + 1, j_hi) - j_lo;
  for (int r = r0; r <= r1; ++r)
  {
    const int32_t v = tile[r][tx];
    if (v >= cap)
      continue;
    const int d = A.sub * (ib - (j_lo + r)) - A.par;
    const int32_t a = v + d * d, bb = 2 * A.sub * d;
#pragma unroll
    for (int k = 0; k < kWinPer; ++k)
      best[k] = min(best[k], a + k * bb);
  }
#pragma unroll
  for (int k = 0; k < kWinPer; ++k)
    best[k] += A.sub * A.sub * k * k;
#pragma unroll
  for (int k = 0; k < kWinPer; ++k)
  {
    const int i = ib + k;
    if (i >= A.len_out)
      break;
    int32_t d2 = best[k];  // = min(exact, cap)
    const uint64_t o = plane * (uint64_t)A.len_out * A.ncol + (uint64_t)i * A.ncol + col;
    if (!kFinal)
    {
      out16[o] = (uint16_t)d2;
      continue;
    }
    if (F.prev != nullptr)
      d2 = min(d2, F.prev[o]);
    if (F.enabled)
    {
      if (F.nsparse > 0)
      {
        // lattice coordinates stay below 26000 (checked by the host), so three squared differences fit an int32
        const int qx = A.sub * (int)(col % (uint64_t)F.sx);
        const int qy = A.sub * (int)(col / (uint64_t)F.sx);
        const int qz = A.sub * i;
        for (int sidx = 0; sidx < F.nsparse; ++sidx)
        {
          const int4 h = F.sparse[sidx];
          const int dx = qx - h.x, dy = qy - h.y, dz = qz - h.z;
          d2 = min(d2, dx * dx + dy * dy + dz * dz);
        }
      }
      F.prob[o] = prob_of(F.g, d2);  // prob_of(cap) is exactly 0 by the definition of cap
      if (F.d2_out != nullptr)
        F.d2_out[o] = d2;
    }
    else
      F.d2_out[o] = d2;
  }
}

// no dense class at all: brute force over the sparse sites only
__global__ void sparse_only_kernel(uint64_t cells, int sx, int sy, int sub, const int4* __restrict__ sparse, int nsparse,
                                   const __grid_constant__ Gauss g, float* __restrict__ prob, int32_t* __restrict__ d2_out)
{
  const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= cells)
    return;
  const long long qx = (long long)sub * (long long)(o % (uint64_t)sx);
  const long long qy = (long long)sub * (long long)((o / (uint64_t)sx) % (uint64_t)sy);
  const long long qz = (long long)sub * (long long)(o / ((uint64_t)sx * sy));
  long long best = kInf;
  for (int s = 0; s < nsparse; ++s)
  {
    const int4 h = sparse[s];
    const long long dx = qx - h.x, dy = qy - h.y, dz = qz - h.z;
    const long long dd = dx * dx + dy * dy + dz * dz;
    best = dd < best ? dd : best;
  }
  prob[o] = prob_of(g, (int32_t)best);
  if (d2_out)
    d2_out[o] = (int32_t)best;
}

// ---- computeGrid2 splat -----------------------------------------------------------------------------
struct SplatArgs
{
  double min[3];
  double resol;
  uint32_t size[3];
  uint32_t step_y, step_z, grid_size;  // u32 arithmetic like PointCloudTools.cpp:192-195
  float c1, c2;
};
__global__ void splat_kernel(const __grid_constant__ SplatArgs S, const float* __restrict__ xyz, uint32_t stride, uint64_t n,
                             int* __restrict__ prob_bits)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 27)
    return;
  const uint64_t pi = t / 27;
  const int nb = (int)(t % 27);
  const int i = nb / 9 - 1, j = (nb / 3) % 3 - 1, k = nb % 3 - 1;  // x, y, z offsets (PointCloudTools.cpp:225-233)
  const float* pt = xyz + pi * stride;
  // PointCloudTools.cpp:208-210: (f32 - f64) / f64 truncated to u32
  const uint32_t ix = __double2uint_rz(__ddiv_rn(__dsub_rn((double)pt[0], S.min[0]), S.resol));
  const uint32_t iy = __double2uint_rz(__ddiv_rn(__dsub_rn((double)pt[1], S.min[1]), S.resol));
  const uint32_t iz = __double2uint_rz(__ddiv_rn(__dsub_rn((double)pt[2], S.min[2]), S.resol));
  const int xx = (int)(ix + i), yy = (int)(iy + j), zz = (int)(iz + k);
  if (xx < 0 || (uint32_t)xx >= S.size[0])
    return;
  if (yy < 0 || (uint32_t)yy >= S.size[1])
    return;
  if (zz < 0 || (uint32_t)zz > S.size[2])  // `>` as written (PointCloudTools.cpp:236)
    return;
  const uint32_t index = (uint32_t)xx + (uint32_t)yy * S.step_y + (uint32_t)zz * S.step_z;
  if (index >= S.grid_size)  // PointCloudTools.cpp:238
    return;
  const float qx = __double2float_rn(__dadd_rn(S.min[0], __dmul_rn((double)xx, S.resol)));
  const float qy = __double2float_rn(__dadd_rn(S.min[1], __dmul_rn((double)yy, S.resol)));
  const float qz = __double2float_rn(__dadd_rn(S.min[2], __dmul_rn((double)zz, S.resol)));
  const float dx = __fsub_rn(qx, pt[0]), dy = __fsub_rn(qy, pt[1]), dz = __fsub_rn(qz, pt[2]);
  const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const float v = __fmul_rn(S.c1, expf(__fmul_rn(__fmul_rn(-dist, dist), S.c2)));
  // values are >= 0, so the IEEE bit patterns order like ints: max is associative -> deterministic
  atomicMax(&prob_bits[index], __float_as_int(v));
}

// ---- lossless dictionary coding of the field ----------------------------------------------------------
// smallest d2 in [0, limit) whose probability is exactly 0.0f (prob is non-increasing in d2)
__global__ void find_cap_kernel(const __grid_constant__ Gauss g, uint32_t limit, unsigned int* __restrict__ cap)
{
  const uint32_t d2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (d2 < limit && prob_of(g, (int32_t)d2) == 0.f)
    atomicMin(cap, d2);
}
// present[min(d2, cap)] = 1
__global__ void present_kernel(const int32_t* __restrict__ d2, uint64_t cells, uint32_t cap, uint8_t* __restrict__ present)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const int32_t v = d2[i];
    const uint32_t k = (v < 0 || (uint32_t)v > cap) ? cap : (uint32_t)v;
    if (!present[k])
      present[k] = 1;
  }
}
// lut[c] = prob_of(values[c]) with the same device function that wrote the float grid
__global__ void lut_kernel(const __grid_constant__ Gauss g, const int32_t* __restrict__ values, uint32_t n, float* __restrict__ lut)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n)
    lut[c] = prob_of(g, values[c]);
}
// one thread per four consecutive code slots (four x-neighbours of one brick row), in coded (output) order
template <typename TCode, int kRep>
__global__ void encode_kernel(const int32_t* __restrict__ d2, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t nbx, uint32_t nby,
                              uint64_t slots, uint32_t cap, const uint16_t* __restrict__ rank, TCode* __restrict__ codes)
{
  const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (t >= slots)
    return;
  uint32_t ix, iy, iz;
  coded_voxel<kRep>(t, nbx, nby, ix, iy, iz);  // ix is a multiple of 4; slots t+1..t+3 are ix+1..ix+3
  uint32_t k[4] = { cap, cap, cap, cap };      // padding voxels: the zero class (never addressed by the kernel)
  if (iy < sy && iz < sz && ix < sx)
  {
    const uint64_t base = (uint64_t)ix + (uint64_t)iy * sx + (uint64_t)iz * sx * sy;
    int32_t v[4] = { -1, -1, -1, -1 };
    if (ix + 3 < sx && (base & 3) == 0)
    {
      const int4 q = *reinterpret_cast<const int4*>(d2 + base);
      v[0] = q.x;
      v[1] = q.y;
      v[2] = q.z;
      v[3] = q.w;
    }
    else
    {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (ix + e < sx)
          v[e] = d2[base + e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      k[e] = (v[e] < 0 || (uint32_t)v[e] > cap) ? cap : (uint32_t)v[e];
  }
  TCode out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
    out[e] = (TCode)rank[k[e]];
  if constexpr (sizeof(TCode) == 1)
    *reinterpret_cast<uchar4*>(codes + t) = make_uchar4(out[0], out[1], out[2], out[3]);
  else
    *reinterpret_cast<ushort4*>(codes + t) = make_ushort4(out[0], out[1], out[2], out[3]);
}


// ====================================================================================================================
// Fused fast build (default when the exact EDT is not kept and there is a single dense lattice class)
// ====================================================================================================================
// The field only needs min(d2, cap) (cap = smallest squared lattice distance whose probability is exactly 0.0f), so the
// minimising site lies within R = floor(sqrt(cap - 1)) lattice steps along every axis and each separable pass is a
// min-plus stencil over at most 2H+1 sites.  Three launches, HBM-streaming, no int32 intermediates:
//   xpass8_kernel     occupancy bit rows -> u8  |dx| (capped), 32 outputs per thread by two register sweeps
//   ypass_dpx_kernel  u8 -> u16 min(dx^2 + dy^2, cap); two columns per lane packed as u16x2, candidates evaluated with
//                     the DPX instruction VIADDMNMX.U16x2 (min(a + b, c) on both halves), squared offsets as immediates
//   zpass_dpx_kernel  u16 -> min(d^2, cap), then in the same epilogue the dictionary code (the set of attainable d^2 of the
//                     class is known beforehand: sums of three squares of the class parity below cap), the float cell
//                     prob = lut[code] (PointCloudTools.cpp:160-162, evaluated once per dictionary entry by lut_kernel with
//                     the same device function as the legacy writer) into the linear float grid, and the code into the
//                     128-byte bricks of the page-blocked copy (staged in shared memory, written as whole bricks).
// Sites of sparse classes (e.g. the two extreme corner points that pin the bounds) are patched in afterwards over their
// (2R+1)^3 boxes.  Results are identical to the legacy passes bit for bit (test_gpu_grid.py).
struct FastArgs
{
  uint32_t sx, sy, sz;  // outputs per axis
  uint32_t w1, w2;      // site rows along y and z (size + 1)
  uint32_t px;          // x pitch of the intermediates, multiple of 64
  uint32_t cap;
};
constexpr int kFastPer = 8;  // outputs per thread and sweep of the DPX passes

__global__ void __launch_bounds__(256)
    xpass8_kernel(const uint32_t* __restrict__ bits, uint32_t words_per_row, uint64_t rows, uint32_t px, int sub, int par,
                  uint32_t d1cap, uint8_t* __restrict__ d1)
{
  const uint32_t chunks = px / 32;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * chunks)
    return;
  const uint64_t row = t / chunks;
  const uint32_t wq = (uint32_t)(t % chunks);
  const uint32_t* r = bits + row * words_per_row;
  auto word = [&](int64_t w) -> uint32_t { return (w >= 0 && w < (int64_t)words_per_row) ? r[w] : 0u; };
  const uint32_t W = word(wq), p1 = word((int64_t)wq - 1), p2 = word((int64_t)wq - 2), n1 = word((int64_t)wq + 1),
                 n2 = word((int64_t)wq + 2);
  constexpr int kFar = 1 << 20;
  // distance (in sites) from the position just left of this word to the nearest site at or before it / from the
  // position just right of it to the nearest site at or after it; two words each side cover any window (H <= 46)
  int carry_l = p1 ? __clz(p1) : (p2 ? 32 + __clz(p2) : kFar);
  int carry_r = n1 ? __ffs(n1) - 1 : (n2 ? 32 + __ffs(n2) - 1 : kFar);
  int dl[32];
  int run = carry_l;
#pragma unroll
  for (int b = 0; b < 32; ++b)
  {
    run = ((W >> b) & 1u) ? 0 : run + 1;
    dl[b] = run;
  }
  uint32_t out[8];
  run = carry_r;
#pragma unroll
  for (int b = 31; b >= 0; --b)
  {
    run = ((W >> b) & 1u) ? 0 : run + 1;
    // site at or left: |sub*dl - par|; site at or right: sub*dr + par
    int d = min(abs(sub * dl[b] - par), sub * run + par);
    d = min(d, (int)d1cap);
    if ((b & 3) == 3)
      out[b >> 2] = 0;
    out[b >> 2] |= (uint32_t)d << (8 * (b & 3));
  }
  uint4* dst = reinterpret_cast<uint4*>(d1 + row * px + (uint64_t)wq * 32);
  dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
  dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

// one sweep of the min-plus stencil for kFastPer consecutive outputs of one lane (two packed columns): rows[t] is the
// candidate site row (first output - H + t); delta = output - site is a compile-time constant of every (t, k) pair, so
// its squared offset is an immediate operand of VIADDMNMX
// explicit global-space accesses for the walking pointers of the two passes (an opaque pointer would otherwise be
// dereferenced through the generic address space)
__device__ __forceinline__ uint32_t ldg_nc_u32(const void* p)
{
  uint32_t v;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldg_nc_u16(const void* p)
{
  uint16_t v;
  asm volatile("ld.global.nc.b16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_f32x2(float* p, float a, float b)
{
  asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

template <int SUB, int PAR, int H, typename RowFn>
__device__ __forceinline__ void dpx_sweep(RowFn row, uint32_t best[kFastPer])
{
#pragma unroll
  for (int t = 0; t < kFastPer + 2 * H; ++t)
  {
    const uint32_t v = row(t);
#pragma unroll
    for (int k = 0; k < kFastPer; ++k)
    {
      const int delta = k - (t - H);
      if (delta >= -H && delta <= H)
      {
        const int d = SUB * delta - PAR;
        best[k] = __viaddmin_u16x2(v, (uint32_t)(d * d) * 0x10001u, best[k]);
      }
    }
  }
}

// in: d1 [plane z][site row j][px] u8, out: dxy [plane z][output row i][px] u16
template <int SUB, int PAR, int H>
__global__ void __launch_bounds__(256) ypass_dpx_kernel(const uint8_t* __restrict__ d1, uint16_t* __restrict__ dxy,
                                                       const __grid_constant__ FastArgs A)
{
  constexpr int kTI = 8 * kFastPer;  // 8 warps x 8 outputs
  constexpr int kRows = kTI + 2 * H;
  __shared__ uint32_t tile[kRows][32];
  const uint32_t ctiles = A.px / 64, itiles = (A.sy + kTI - 1) / kTI;
  uint32_t b = blockIdx.x;
  const uint32_t it = b % itiles;
  b /= itiles;
  const uint32_t ct = b % ctiles;
  const uint32_t plane = b / ctiles;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int i0 = (int)(it * kTI);
  const uint32_t cap = A.cap, cap2 = cap * 0x10001u;
  {
    // warp w stages site rows r = w, w + 8, ...: one pointer walking down the plane by 8 rows a step, one unsigned
    // compare per row (j < 0 wraps above w1); every load of the tile is issued before the first one is consumed
    constexpr int kLoads = (kRows + 7) / 8;
    const int j0 = i0 - H + (int)warp;
    const uint8_t* p = d1 + ((int64_t)plane * A.w1 + j0) * (int64_t)A.px + ct * 64 + 2 * lane;
    const uint64_t step = 8ull * A.px;
    uint32_t raw[kLoads];
#pragma unroll
    for (int q = 0; q < kLoads; ++q)
    {
      raw[q] = 0xFFFFu;  // 255, 255: no site
      if ((int)warp + 8 * q < kRows && (uint32_t)(j0 + 8 * q) < A.w1)
        raw[q] = ldg_nc_u16(p);
      p += step;
      asm volatile("" : "+l"(p));  // keep the walking pointer: one 64-bit add per row instead of a 64-bit multiply-add chain
    }
    uint32_t* t = &tile[warp][lane];
#pragma unroll
    for (int q = 0; q < kLoads; ++q)
    {
      const uint32_t a = raw[q] & 0xFFu, c = raw[q] >> 8;
      if ((int)warp + 8 * q < kRows)
        t[q * 8 * 32] = __vminu2(a * a | ((c * c) << 16), cap2);  // a, c <= 255: both squares fit their 16-bit half
    }
  }
  __syncthreads();
  uint32_t best[kFastPer];
#pragma unroll
  for (int k = 0; k < kFastPer; ++k)
    best[k] = cap2;
  const uint32_t base = warp * kFastPer;
  dpx_sweep<SUB, PAR, H>([&](int t) { return tile[base + t][lane]; }, best);
  const uint32_t i_first = (uint32_t)i0 + base;
  uint32_t* dst = reinterpret_cast<uint32_t*>(dxy + ((uint64_t)plane * A.sy + i_first) * A.px + ct * 64 + 2 * lane);
  const uint32_t left = A.sy > i_first ? A.sy - i_first : 0u;  // outputs of this thread that exist
  const uint32_t pitch = A.px / 2;
#pragma unroll
  for (int k = 0; k < kFastPer; ++k)
    if ((uint32_t)k < left)
      dst[(uint64_t)k * pitch] = best[k];
}

// in: dxy [site row j = z][y][px] u16; out: prob [z][y][sx] f32 and the bricked u8 codes.
// Block tile: 64 x (32 lanes x 2) x 4 y x 32 z = 8 x 1 x 8 bricks; thread = (lane, y, z group of 16) does two sweeps of 8.
constexpr int kZT = 32;
constexpr int kBrickPitch = 132;  // bytes between staged bricks (33 words: conflict-free column writes)
template <int SUB, int PAR, int H>
__global__ void __launch_bounds__(256)
    zpass_dpx_kernel(const uint16_t* __restrict__ dxy, float* __restrict__ prob, uint8_t* __restrict__ codes,
                     const __grid_constant__ FastArgs A, const uint8_t* __restrict__ rank, const float* __restrict__ lut,
                     uint32_t nbx, uint32_t nby)
{
  constexpr int kRows = kZT + 2 * H;
  static_assert(kRows % 2 == 0, "two site rows per staging step");
  extern __shared__ __align__(16) uint32_t zsm[];
  uint32_t(*tile)[4][32] = reinterpret_cast<uint32_t(*)[4][32]>(zsm);                          // [kRows][4][32]
  uint8_t* stage = reinterpret_cast<uint8_t*>(zsm + (size_t)kRows * 4 * 32);                  // 64 bricks x kBrickPitch
  // the two tables of the epilogue in shared memory: 32-bit addresses and no L1 tag lookups in the hot loop
  float* lut_s = reinterpret_cast<float*>(stage + 64 * kBrickPitch);                          // 256 dictionary values
  uint8_t* rank_s = reinterpret_cast<uint8_t*>(lut_s + 256);                                  // code of every d2 <= cap
  for (uint32_t w = threadIdx.x; w < (A.cap + 4u) / 4u; w += 256)
    reinterpret_cast<uint32_t*>(rank_s)[w] = reinterpret_cast<const uint32_t*>(rank)[w];
  lut_s[threadIdx.x] = lut[threadIdx.x];
  const uint32_t ztiles = (A.sz + kZT - 1) / kZT, ctiles = A.px / 64;
  uint32_t b = blockIdx.x;
  const uint32_t zt = b % ztiles;  // z fastest: neighbouring blocks share their halo rows in L2
  b /= ztiles;
  const uint32_t ct = b % ctiles;
  const uint32_t yt = b / ctiles;
  const uint32_t lane = threadIdx.x & 31u, ty = (threadIdx.x >> 5) & 3u, zg = threadIdx.x >> 7;
  const uint32_t x0 = ct * 64, y0 = yt * 4;
  const int z0 = (int)(zt * kZT);
  const uint32_t cap = A.cap, cap2 = cap * 0x10001u;
  {
    // warp w stages y row (w & 3) of the site rows r = (w >> 2), (w >> 2) + 2, ...: one pointer walking up by two site
    // rows a step, one unsigned compare per row (a negative j wraps above the limit, a y row outside the grid has limit 0);
    // loads in batches of 12, all issued before the first store
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t yy = warp & 3u, rb = warp >> 2;
    const int j0 = z0 - H + (int)rb;
    const uint32_t jlim = (y0 + yy < A.sy) ? A.w2 : 0u;
    const uint64_t row = (uint64_t)A.sy * A.px;  // elements between site rows
    const uint16_t* p = dxy + ((int64_t)j0 * (int64_t)A.sy + (int64_t)(y0 + yy)) * (int64_t)A.px + x0 + 2 * lane;
    uint32_t* t = &tile[rb][yy][lane];
    constexpr int kM = kRows / 2;  // staging steps of this warp
    constexpr int kBatch = 12;
#pragma unroll
    for (int m0 = 0; m0 < kM; m0 += kBatch)
    {
      uint32_t raw[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u)
      {
        const int m = m0 + u;
        raw[u] = cap2;
        if (m < kM)
        {
          if ((uint32_t)(j0 + 2 * m) < jlim)
            raw[u] = ldg_nc_u32(p);
          p += 2 * row;
          asm volatile("" : "+l"(p));  // walking pointer: one 64-bit add per staging step
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u)
      {
        const int m = m0 + u;
        if (m < kM)
          t[m * 2 * 4 * 32] = raw[u];
      }
    }
  }
  __syncthreads();
  const uint32_t x = x0 + 2 * lane, y = y0 + ty;
  // how this thread's two x neighbours go out: 0 nothing, 1 first only, 2 both as one 8-byte store, 3 both as two stores
  const uint32_t xmode = (y >= A.sy || x >= A.sx) ? 0u : (x + 1 >= A.sx ? 1u : ((A.sx & 1u) ? 3u : 2u));
  const uint64_t zstride = (uint64_t)A.sy * A.sx;
  const bool interior = (x0 + 64 <= A.sx) && (y0 + 4 <= A.sy) && !(A.sx & 1u);  // block-uniform
#pragma unroll 1
  for (int s = 0; s < 2; ++s)
  {
    const uint32_t zb = zg * 16 + s * 8;  // first output of this sweep, relative to the tile (a multiple of 8)
    uint32_t best[kFastPer];
#pragma unroll
    for (int k = 0; k < kFastPer; ++k)
      best[k] = cap2;
    dpx_sweep<SUB, PAR, H>([&](int t) { return tile[zb + t][ty][lane]; }, best);
    // staged brick (bz, bx) of the tile, voxel (z&3, y&3, x&7): two x neighbours = one 16-bit store; zb is a multiple of
    // 8, so the brick row and the z-in-brick bits of output k are compile-time offsets from one base
    uint8_t* st = stage + ((zb >> 2) * 8 + (lane >> 2)) * kBrickPitch + ((ty << 3) | ((2 * lane) & 7u));
    float* dst = prob + ((uint64_t)((uint32_t)z0 + zb) * A.sy + y) * A.sx + x;
    const uint32_t zleft = A.sz > (uint32_t)z0 + zb ? A.sz - ((uint32_t)z0 + zb) : 0u;  // outputs of this sweep inside the grid
    if (interior && zleft >= (uint32_t)kFastPer)
    {
      // the whole tile lies inside the grid and rows are 8-byte aligned: no per-output tests, one 8-byte store each
#pragma unroll
      for (int k = 0; k < kFastPer; ++k)
      {
        const uint32_t ca = rank_s[best[k] & 0xFFFFu], cb = rank_s[best[k] >> 16];
        *reinterpret_cast<uint16_t*>(st + (k >> 2) * 8 * kBrickPitch + ((k & 3) << 5)) = (uint16_t)(ca | (cb << 8));
        stg_f32x2(dst, lut_s[ca], lut_s[cb]);
        dst += zstride;
        asm volatile("" : "+l"(dst));
      }
      continue;
    }
#pragma unroll
    for (int k = 0; k < kFastPer; ++k)
    {
      const uint32_t ca = rank_s[best[k] & 0xFFFFu], cb = rank_s[best[k] >> 16];
      *reinterpret_cast<uint16_t*>(st + (k >> 2) * 8 * kBrickPitch + ((k & 3) << 5)) = (uint16_t)(ca | (cb << 8));
      if ((uint32_t)k < zleft && xmode != 0u)
      {
        const float pa = lut_s[ca];
        if (xmode == 2u)
          *reinterpret_cast<float2*>(dst) = make_float2(pa, lut_s[cb]);
        else
        {
          dst[0] = pa;
          if (xmode == 3u)
            dst[1] = lut_s[cb];
        }
      }
      dst += zstride;
    }
  }
  __syncthreads();
  // whole bricks out: thread = (brick, quarter), 32 bytes each; the 8 x-bricks of one z layer are 1 KB contiguous
  {
    const uint32_t brick = threadIdx.x >> 2, part = threadIdx.x & 3u;
    const uint32_t bz = brick >> 3, bx = brick & 7u;
    const uint32_t* srcw = reinterpret_cast<const uint32_t*>(stage + brick * kBrickPitch) + part * 8;
    const uint64_t o = coded_index<kRepCode8>(x0 + bx * 8, y0, (uint32_t)z0 + bz * 4, nbx, nby) + part * 32;
    uint4* dst = reinterpret_cast<uint4*>(codes + o);
    dst[0] = make_uint4(srcw[0], srcw[1], srcw[2], srcw[3]);
    dst[1] = make_uint4(srcw[4], srcw[5], srcw[6], srcw[7]);
  }
}

// ---- sparse sites after the fused passes ------------------------------------------------------------------------------
// thread = (sparse site s, voxel of its (2W+1)^3 box); the voxel's squared distance to the NEAREST sparse site (all of
// them are scanned, so overlapping boxes compute identical values and no atomics are needed) against what the dense
// class left there.  mark: needed[d2] = 1 for values that win; apply: rewrite code and float cell.
template <bool kApply>
__global__ void sparse_patch_kernel(const int4* __restrict__ sparse, int nsparse, int W, int sub, uint32_t sx, uint32_t sy,
                                    uint32_t sz, uint32_t nbx, uint32_t nby, uint32_t cap, const int32_t* __restrict__ values,
                                    const uint8_t* __restrict__ rank, const float* __restrict__ lut, uint8_t* __restrict__ codes,
                                    float* __restrict__ prob, uint8_t* __restrict__ needed)
{
  const int side = 2 * W + 1;
  const uint64_t per = (uint64_t)side * side * side;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per * (uint64_t)nsparse)
    return;
  const int s = (int)(t / per);
  uint64_t v = t % per;
  const int4 h = sparse[s];
  // voxel nearest to the site, then the box around it
  const int cx = (h.x + sub / 2) / sub, cy = (h.y + sub / 2) / sub, cz = (h.z + sub / 2) / sub;
  const int ix = cx - W + (int)(v % side);
  v /= side;
  const int iy = cy - W + (int)(v % side);
  const int iz = cz - W + (int)(v / side);
  if (ix < 0 || iy < 0 || iz < 0 || ix >= (int)sx || iy >= (int)sy || iz >= (int)sz)
    return;
  int best = INT32_MAX;
  for (int q = 0; q < nsparse; ++q)
  {
    const int4 g = sparse[q];
    const int dx = sub * ix - g.x, dy = sub * iy - g.y, dz = sub * iz - g.z;
    if (abs(dx) > 30000 || abs(dy) > 30000 || abs(dz) > 30000)
      continue;
    best = min(best, dx * dx + dy * dy + dz * dz);
  }
  if ((uint32_t)best >= cap)
    return;
  const uint64_t o = coded_index<kRepCode8>((uint32_t)ix, (uint32_t)iy, (uint32_t)iz, nbx, nby);
  const int32_t cur = values[codes[o]];
  if (best >= cur)
    return;
  if (!kApply)
    needed[best] = 1;
  else
  {
    const uint8_t c = rank[best];
    codes[o] = c;
    prob[((uint64_t)iz * sy + iy) * sx + ix] = lut[c];
  }
}

template <int SUB, int PAR, int H>
static cudaError_t launch_y_t(const FastArgs& A, const uint8_t* d1, uint16_t* dxy, unsigned blocks, cudaStream_t stream)
{
  ypass_dpx_kernel<SUB, PAR, H><<<blocks, 256, 0, stream>>>(d1, dxy, A);
  return cudaGetLastError();
}
template <int SUB, int PAR, int H>
static cudaError_t launch_z_t(const FastArgs& A, const uint16_t* dxy, float* prob, uint8_t* codes, const uint8_t* rank,
                              const float* lut, uint32_t nbx, uint32_t nby, unsigned blocks, cudaStream_t stream)
{
  const size_t smem = (size_t)(kZT + 2 * H) * 4 * 32 * 4 + 64 * kBrickPitch + 256 * sizeof(float) + (((size_t)A.cap + 4) & ~(size_t)3);
  // per-device attribute of the function: set on every launch (cheap), never cached process-wide
  cudaError_t e = cudaFuncSetAttribute(zpass_dpx_kernel<SUB, PAR, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess)
    return e;
  zpass_dpx_kernel<SUB, PAR, H><<<blocks, 256, smem, stream>>>(dxy, prob, codes, A, rank, lut, nbx, nby);
  return cudaGetLastError();
}
// window half-widths the passes are instantiated for (a wider window than needed is still exact: the extra candidates
// are farther than R and cannot beat the cap)
#define A3D_FOR_H(CALL)   \
  if (h <= 12)            \
    return CALL(12);      \
  if (h <= 16)            \
    return CALL(16);      \
  if (h <= 18)            \
    return CALL(18);      \
  if (h <= 20)            \
    return CALL(20);      \
  if (h <= 24)            \
    return CALL(24);      \
  if (h <= 28)            \
    return CALL(28);      \
  if (h <= 36)            \
    return CALL(36);      \
  return CALL(46);
template <int SUB, int PAR>
static cudaError_t launch_y_h(int h, const FastArgs& A, const uint8_t* d1, uint16_t* dxy, unsigned blocks, cudaStream_t stream)
{
#define A3D_Y(HH) launch_y_t<SUB, PAR, HH>(A, d1, dxy, blocks, stream)
  A3D_FOR_H(A3D_Y)
#undef A3D_Y
}
template <int SUB, int PAR>
static cudaError_t launch_z_h(int h, const FastArgs& A, const uint16_t* dxy, float* prob, uint8_t* codes, const uint8_t* rank,
                              const float* lut, uint32_t nbx, uint32_t nby, unsigned blocks, cudaStream_t stream)
{
#define A3D_Z(HH) launch_z_t<SUB, PAR, HH>(A, dxy, prob, codes, rank, lut, nbx, nby, blocks, stream)
  A3D_FOR_H(A3D_Z)
#undef A3D_Z
}
static cudaError_t launch_y(int sub, int par, int h, const FastArgs& A, const uint8_t* d1, uint16_t* dxy, unsigned blocks,
                            cudaStream_t stream)
{
  if (sub == 1)
    return launch_y_h<1, 0>(h, A, d1, dxy, blocks, stream);
  return par ? launch_y_h<2, 1>(h, A, d1, dxy, blocks, stream) : launch_y_h<2, 0>(h, A, d1, dxy, blocks, stream);
}
static cudaError_t launch_z(int sub, int par, int h, const FastArgs& A, const uint16_t* dxy, float* prob, uint8_t* codes,
                            const uint8_t* rank, const float* lut, uint32_t nbx, uint32_t nby, unsigned blocks, cudaStream_t stream)
{
  if (sub == 1)
    return launch_z_h<1, 0>(h, A, dxy, prob, codes, rank, lut, nbx, nby, blocks, stream);
  return par ? launch_z_h<2, 1>(h, A, dxy, prob, codes, rank, lut, nbx, nby, blocks, stream)
             : launch_z_h<2, 0>(h, A, dxy, prob, codes, rank, lut, nbx, nby, blocks, stream);
}
constexpr int kFastMaxH = 46;
constexpr int kFastMaxSparse = 64;

void gauss_consts(double sensor_dev, float* c1, float* c2)
{
  *c1 = static_cast<float>(1. / (sensor_dev * sqrt(2 * M_PI)));   // PointCloudTools.cpp:139
  *c2 = static_cast<float>(1. / (2. * sensor_dev * sensor_dev));  // PointCloudTools.cpp:140
}

#define CK(expr)                     \
  do                                 \
  {                                  \
    cudaError_t e__ = (expr);        \
    if (e__ != cudaSuccess)          \
    {                                \
      rc = e__;                      \
      goto done;                     \
    }                                \
  } while (0)

}  // namespace

cudaError_t launch_bounds(const float* xyz, uint32_t stride, uint64_t n, float* out6, cudaStream_t stream)
{
  // out6 is used as 6 ordered-uint keys on the device; the caller decodes them
  uint32_t init[6] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u };
  cudaError_t e = cudaMemcpyAsync(out6, init, sizeof(init), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess)
    return e;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8)
    blocks = 148 * 8;
  if (blocks == 0)
    blocks = 1;
  bounds_kernel<<<(unsigned)blocks, 256, 0, stream>>>(xyz, stride, n, reinterpret_cast<uint32_t*>(out6));
  return cudaGetLastError();
}

cudaError_t build_grid_splat(const MapDev& map, const float* d_xyz, uint32_t stride, uint64_t n, double sensor_dev,
                             float* prob, cudaStream_t stream)
{
  SplatArgs S;
  for (int a = 0; a < 3; ++a)
  {
    S.min[a] = map.min[a];
    S.size[a] = map.size[a];
  }
  S.resol = map.resol;
  S.step_y = map.size[0];
  S.step_z = map.size[0] * map.size[1];
  S.grid_size = map.size[0] * map.size[1] * map.size[2];
  gauss_consts(sensor_dev, &S.c1, &S.c2);
  cudaError_t e = cudaMemsetAsync(prob, 0, sizeof(float) * map.n_cells, stream);
  if (e != cudaSuccess)
    return e;
  if (n == 0)
    return cudaSuccess;
  const uint64_t threads = n * 27;
  splat_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(S, d_xyz, stride, n, reinterpret_cast<int*>(prob));
  return cudaGetLastError();
}

// smallest squared lattice distance whose probability is exactly 0.0f; 0xFFFFFFFF if there is none below the limit
static cudaError_t find_cap(const Gauss& G, unsigned int* h_cap, cudaStream_t stream)
{
  const uint32_t kLimit = 1u << 22;
  unsigned int* d_cap = nullptr;
  *h_cap = 0xFFFFFFFFu;
  cudaError_t rc = cudaMalloc(&d_cap, sizeof(unsigned int));
  if (rc != cudaSuccess)
    return rc;
  rc = cudaMemcpyAsync(d_cap, h_cap, sizeof(unsigned int), cudaMemcpyHostToDevice, stream);
  if (rc == cudaSuccess)
  {
    find_cap_kernel<<<kLimit / 256, 256, 0, stream>>>(G, kLimit, d_cap);
    rc = cudaGetLastError();
  }
  if (rc == cudaSuccess)
    rc = cudaMemcpyAsync(h_cap, d_cap, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
  if (rc == cudaSuccess)
    rc = cudaStreamSynchronize(stream);
  cudaFree(d_cap);
  return rc;
}


// The fused fast build for one dense class (see the kernel block above).  *done = false (and nothing kept) when the
// configuration is outside what it covers: the caller then runs the legacy passes.
static cudaError_t build_fast(const MapDev& map, const Lattice& L, const float* d_xyz, uint32_t stride, uint64_t n, const Gauss& G,
                              int cls, const int4* d_sparse, int nsparse, uint64_t pblocks, float* prob, CodedGrid* out,
                              cudaStream_t stream, cudaEvent_t ev_start, cudaEvent_t ev_end, bool* done)
{
  cudaError_t rc = cudaSuccess;
  *done = false;
  const int sub = L.sub;
  const int par[3] = { cls & 1, (cls >> 1) & 1, (cls >> 2) & 1 };
  const uint32_t sx = map.size[0], sy = map.size[1], sz = map.size[2];
  unsigned int h_cap = 0xFFFFFFFFu;
  uint32_t* d_bits = nullptr;
  uint8_t* d_d1 = nullptr;
  uint16_t* d_dxy = nullptr;
  uint8_t* d_rank = nullptr;
  uint8_t* d_needed = nullptr;
  int32_t* d_values = nullptr;
  float* d_lut = nullptr;
  void* d_codes = nullptr;
  std::vector<int32_t> values;
  std::vector<uint8_t> rank, needed;

  CK(find_cap(G, &h_cap, stream));
  if (h_cap == 0xFFFFFFFFu || h_cap < 2 || h_cap > 30000u)
    goto done;
  {
    const uint32_t cap = h_cap;
    const int R = (int)floor(sqrt((double)(cap - 1)));
    const int H = (R + 1) / sub + 1;  // candidate sites on each side of an output
    uint32_t d1cap = (uint32_t)R + 1;  // d1cap^2 >= cap
    while (d1cap * d1cap < cap)
      ++d1cap;
    if (H > kFastMaxH || d1cap > 255u || nsparse > kFastMaxSparse)
      goto done;
    // dictionary known beforehand: sums of three squares of the class parity below cap, then the zero class
    {
      std::vector<uint8_t> present(cap + 1, 0);
      auto axis = [&](int pa, std::vector<int>& o) {
        for (int d = 0; d <= R + sub; ++d)
          if (sub == 1 || (d & 1) == pa)
            o.push_back(d * d);
      };
      std::vector<int> ax, ay, az;
      axis(par[0], ax);
      axis(par[1], ay);
      axis(par[2], az);
      for (int a : ax)
        for (int b : ay)
          for (int c : az)
            if ((uint32_t)(a + b + c) < cap)
              present[a + b + c] = 1;
      rank.assign(cap + 1, 0);
      for (uint32_t k = 0; k < cap; ++k)
        if (present[k])
        {
          rank[k] = (uint8_t)(values.size() & 0xFF);
          values.push_back((int32_t)k);
        }
      if (values.size() > 253)
        goto done;  // 8-bit codes with 0xFF reserved: the legacy coder decides between u8 and u16
      rank[cap] = (uint8_t)values.size();
      values.push_back((int32_t)cap);
    }
    const uint32_t nbx = (sx + 127) / 128, nby = (sy + 127) / 128, nbz = (sz + 127) / 128;
    const uint64_t code_bytes = ((uint64_t)nbx * nby * nbz) << 21;
    FastArgs A;
    A.sx = sx;
    A.sy = sy;
    A.sz = sz;
    A.w1 = (uint32_t)L.w[1];
    A.w2 = (uint32_t)L.w[2];
    A.px = (sx + 63u) & ~63u;
    A.cap = cap;
    const uint64_t rows = (uint64_t)L.w[1] * L.w[2];
    const uint64_t bits_words = rows * L.words_per_row;
    CK(cudaMalloc(&d_bits, sizeof(uint32_t) * bits_words));
    CK(cudaMalloc(&d_d1, rows * A.px));
    CK(cudaMalloc(&d_dxy, sizeof(uint16_t) * (uint64_t)A.w2 * sy * A.px));
    CK(cudaMalloc(&d_rank, ((size_t)cap + 4) & ~(size_t)3));  // the z pass copies it to shared memory in whole words
    CK(cudaMalloc(&d_needed, cap + 1));
    CK(cudaMalloc(&d_values, sizeof(int32_t) * 256));
    CK(cudaMalloc(&d_lut, sizeof(float) * 256));
    CK(cudaMalloc(&d_codes, code_bytes + (2u << 20)));
    void* const codes_aligned =
        reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(d_codes) + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
    // the timed region of the build (amcl3d_b200_grid_last_build_ms) starts once the workspace exists and ends before it
    // is freed: cudaMalloc / cudaFree of GBs are host-side driver calls, not work of the path
    if (ev_start)
      cudaEventRecord(ev_start, stream);
    CK(cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * bits_words, stream));
    CK(cudaMemsetAsync(d_values, 0, sizeof(int32_t) * 256, stream));
    CK(cudaMemcpyAsync(d_rank, rank.data(), cap + 1, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_values, values.data(), sizeof(int32_t) * values.size(), cudaMemcpyHostToDevice, stream));
    lut_kernel<<<1, 256, 0, stream>>>(G, d_values, (uint32_t)values.size(), d_lut);
    CK(cudaGetLastError());
    voxelize_kernel<<<(unsigned)pblocks, 256, 0, stream>>>(L, d_xyz, stride, n, cls, d_bits, 0u, nullptr, nullptr);
    CK(cudaGetLastError());
    {
      const uint64_t threads = rows * (A.px / 32);
      xpass8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_bits, L.words_per_row, rows, A.px, sub, par[0], d1cap, d_d1);
      CK(cudaGetLastError());
    }
    {
      uint8_t* codes8 = reinterpret_cast<uint8_t*>(codes_aligned);
      cudaError_t e = cudaSuccess;
      // y and z may have different parities: each pass is instantiated per (sub, parity of its axis, window)
      const uint64_t by = (uint64_t)A.w2 * (A.px / 64) * ((A.sy + 8 * kFastPer - 1) / (8 * kFastPer));
      const uint64_t bz = (uint64_t)((A.sz + kZT - 1) / kZT) * (A.px / 64) * ((A.sy + 3) / 4);
      if (by > 0x7FFFFFFFull || bz > 0x7FFFFFFFull)
        goto done;
      e = launch_y(sub, par[1], H, A, d_d1, d_dxy, (unsigned)by, stream);
      if (e == cudaSuccess)
        e = launch_z(sub, par[2], H, A, d_dxy, prob, codes8, d_rank, d_lut, nbx, nby, (unsigned)bz, stream);
      CK(e);
      if (nsparse > 0)
      {
        const int W = R / sub + 1;
        const uint64_t threads = (uint64_t)(2 * W + 1) * (2 * W + 1) * (2 * W + 1) * nsparse;
        const unsigned blocks = (unsigned)((threads + 255) / 256);
        CK(cudaMemsetAsync(d_needed, 0, cap + 1, stream));
        sparse_patch_kernel<false><<<blocks, 256, 0, stream>>>(d_sparse, nsparse, W, sub, sx, sy, sz, nbx, nby, cap, d_values, d_rank,
                                                              d_lut, codes8, prob, d_needed);
        CK(cudaGetLastError());
        needed.resize(cap + 1);
        CK(cudaMemcpyAsync(needed.data(), d_needed, cap + 1, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        bool grown = false;
        for (uint32_t k = 0; k < cap; ++k)
          if (needed[k] && !(rank[k] < values.size() && values[rank[k]] == (int32_t)k))
          {
            if (values.size() >= 254)
              goto done;  // more distinct values than 8-bit codes hold: legacy passes + coder
            rank[k] = (uint8_t)values.size();
            values.push_back((int32_t)k);
            grown = true;
          }
        if (grown)
        {
          CK(cudaMemcpyAsync(d_rank, rank.data(), cap + 1, cudaMemcpyHostToDevice, stream));
          CK(cudaMemcpyAsync(d_values, values.data(), sizeof(int32_t) * values.size(), cudaMemcpyHostToDevice, stream));
          lut_kernel<<<1, 256, 0, stream>>>(G, d_values, (uint32_t)values.size(), d_lut);
          CK(cudaGetLastError());
        }
        sparse_patch_kernel<true><<<blocks, 256, 0, stream>>>(d_sparse, nsparse, W, sub, sx, sy, sz, nbx, nby, cap, d_values, d_rank,
                                                             d_lut, codes8, prob, d_needed);
        CK(cudaGetLastError());
      }
    }
    if (ev_end)
      cudaEventRecord(ev_end, stream);
    CK(cudaStreamSynchronize(stream));
    out->rep = kRepCode8;
    out->codes = codes_aligned;
    out->codes_alloc = d_codes;
    out->lut = d_lut;
    out->lut_n = (uint32_t)values.size();
    out->nbx = nbx;
    out->nby = nby;
    out->bytes = code_bytes;
    d_codes = nullptr;
    d_lut = nullptr;
    *done = true;
  }
done:
  cudaFree(d_bits);
  cudaFree(d_d1);
  cudaFree(d_dxy);
  cudaFree(d_rank);
  cudaFree(d_needed);
  cudaFree(d_values);
  cudaFree(d_codes);
  cudaFree(d_lut);
  return rc;
}

// Builds the dictionary-coded, bricked copy of the field from the linear int32 distance transform.
// On success out->rep is kRepCode8 / kRepCode16; kRepFloat (no allocation) if the field has too many
// distinct values for 16-bit codes or never underflows to zero inside the probed range.
static cudaError_t build_codes(const MapDev& map, const int32_t* d_d2, const Gauss& G, CodedGrid* out, cudaStream_t stream)
{
  cudaError_t rc = cudaSuccess;
  out->rep = kRepFloat;
  out->codes = nullptr;
  out->codes_alloc = nullptr;
  out->lut = nullptr;
  out->lut_n = 0;
  const uint32_t sx = map.size[0], sy = map.size[1], sz = map.size[2];
  const uint64_t cells = (uint64_t)sx * sy * sz;
  uint8_t* d_present = nullptr;
  uint16_t* d_rank = nullptr;
  int32_t* d_values = nullptr;
  void* d_codes = nullptr;
  float* d_lut = nullptr;
  unsigned int h_cap = 0xFFFFFFFFu;
  std::vector<uint8_t> present;
  std::vector<uint16_t> rank;
  std::vector<int32_t> values;

  CK(find_cap(G, &h_cap, stream));
  if (h_cap == 0xFFFFFFFFu)
    goto done;  // the Gaussian never reaches 0.0f: keep the float grid
  {
    const uint32_t cap = h_cap;
    CK(cudaMalloc(&d_present, cap + 1));
    CK(cudaMemsetAsync(d_present, 0, cap + 1, stream));
    present_kernel<<<148 * 16, 256, 0, stream>>>(d_d2, cells, cap, d_present);
    CK(cudaGetLastError());
    present.resize(cap + 1);
    CK(cudaMemcpyAsync(present.data(), d_present, cap + 1, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    present[cap] = 1;  // the zero class always exists (padding)
    rank.assign(cap + 1, 0);
    for (uint32_t k = 0; k <= cap; ++k)
      if (present[k])
      {
        rank[k] = (uint16_t)values.size();
        values.push_back((int32_t)k);
        if (values.size() > 65536)
          break;
      }
    if (values.size() > 65536)
      goto done;
    const uint32_t nvals = (uint32_t)values.size();
    const bool small = nvals <= 256;
    const uint32_t nbx = (sx + 127) / 128, nby = (sy + 127) / 128, nbz = (sz + 127) / 128;
    const uint64_t slots = ((uint64_t)nbx * nby * nbz) << 21;
    const uint64_t code_bytes = slots * (small ? 1 : 2);
    CK(cudaMalloc(&d_rank, sizeof(uint16_t) * (cap + 1)));
    CK(cudaMalloc(&d_values, sizeof(int32_t) * nvals));
    CK(cudaMalloc(&d_lut, sizeof(float) * nvals));
    CK(cudaMalloc(&d_codes, code_bytes + (2u << 20)));
    void* const codes_aligned = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(d_codes) + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
    CK(cudaMemcpyAsync(d_rank, rank.data(), sizeof(uint16_t) * (cap + 1), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_values, values.data(), sizeof(int32_t) * nvals, cudaMemcpyHostToDevice, stream));
    lut_kernel<<<(nvals + 255) / 256, 256, 0, stream>>>(G, d_values, nvals, d_lut);
    CK(cudaGetLastError());
    if (small)
      encode_kernel<uint8_t, kRepCode8><<<(unsigned)((slots / 4 + 255) / 256), 256, 0, stream>>>(d_d2, sx, sy, sz, nbx, nby, slots, cap,
                                                                                           d_rank, (uint8_t*)codes_aligned);
    else
      encode_kernel<uint16_t, kRepCode16><<<(unsigned)((slots / 4 + 255) / 256), 256, 0, stream>>>(d_d2, sx, sy, sz, nbx, nby, slots,
                                                                                             cap, d_rank, (uint16_t*)codes_aligned);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
    out->rep = small ? kRepCode8 : kRepCode16;
    out->codes = codes_aligned;
    out->codes_alloc = d_codes;
    out->lut = d_lut;
    out->lut_n = nvals;
    out->nbx = nbx;
    out->nby = nby;
    out->bytes = code_bytes;
    d_codes = nullptr;
    d_lut = nullptr;
  }
done:
  cudaFree(d_present);
  cudaFree(d_rank);
  cudaFree(d_values);
  cudaFree(d_codes);
  cudaFree(d_lut);
  return rc;
}

cudaError_t build_grid_edt(const MapDev& map, const float* d_xyz, uint32_t stride, uint64_t n, double sensor_dev,
                           int mode, int sub, float* prob, int32_t** keep_d2, CodedGrid* coded, int fast_passes,
                           cudaStream_t stream, cudaEvent_t* ev /* nullable, 4 events */, char* err, size_t errlen)
{
  // timed regions of amcl3d_b200_grid_last_build_ms: A = ev[0]..ev[1], B = ev[2]..ev[3] (see common.cuh)
  cudaEvent_t ev_start = ev ? ev[2] : nullptr, ev_end = ev ? ev[3] : nullptr;
  if (ev)
    for (int k = 0; k < 4; ++k)
      cudaEventRecord(ev[k], stream);
  cudaError_t rc = cudaSuccess;
  const int sx = (int)map.size[0], sy = (int)map.size[1], sz = (int)map.size[2];
  const uint64_t cells = (uint64_t)sx * sy * sz;
  const int maxdim = sx > sy ? (sx > sz ? sx : sz) : (sy > sz ? sy : sz);
  if ((long long)sub * (maxdim + 1) > 26000 || maxdim + 1 > 65535)
  {
    snprintf(err, errlen, "grid axis too long for the int32 EDT (sub*size must stay below 26000)");
    return cudaErrorInvalidValue;
  }
  Lattice L;
  for (int a = 0; a < 3; ++a)
  {
    L.min[a] = map.min[a];
    L.w[a] = (int)map.size[a] + 1;
  }
  L.pitch = map.resol / sub;
  L.rinv = 1.0 / L.pitch;
  L.sub = sub;
  L.words_per_row = (uint32_t)((L.w[0] + 31) / 32);
  const uint64_t rows = (uint64_t)L.w[1] * L.w[2];
  const int nclass = sub == 2 ? 8 : 1;

  unsigned long long* d_counts = nullptr;
  uint32_t* d_bits = nullptr;
  int4* d_sparse = nullptr;
  unsigned int* d_sparse_count = nullptr;
  uint16_t* d_d1 = nullptr;
  int32_t* d_dxy = nullptr;
  uint16_t* d_dxy16 = nullptr;
  int32_t* d_d2 = nullptr;
  int2* d_stack = nullptr;
  unsigned long long h_counts[8] = { 0 };
  unsigned int h_nsparse = 0;
  uint32_t sparse_mask = 0;
  int dense[8], ndense = 0;
  Gauss G;
  gauss_consts(sensor_dev, &G.c1, &G.c2);
  G.unit = L.pitch * L.pitch;
  G.mode = mode;

  uint64_t pblocks = (n + 255) / 256;
  if (pblocks > 148 * 16)
    pblocks = 148 * 16;
  if (pblocks == 0)
    pblocks = 1;

  CK(cudaMalloc(&d_counts, 8 * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(d_counts, 0, 8 * sizeof(unsigned long long), stream));
  class_count_kernel<<<(unsigned)pblocks, 256, 0, stream>>>(L, d_xyz, stride, n, d_counts);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h_counts, d_counts, sizeof(h_counts), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  for (int c = 0; c < nclass; ++c)
  {
    if (h_counts[c] == 0)
      continue;
    if (h_counts[c] <= (unsigned long long)kSparseMax)
      sparse_mask |= 1u << c;
    else
      dense[ndense++] = c;
  }
  CK(cudaMalloc(&d_sparse, sizeof(int4) * 8 * kSparseMax));
  CK(cudaMalloc(&d_sparse_count, sizeof(unsigned int)));
  CK(cudaMemsetAsync(d_sparse_count, 0, sizeof(unsigned int), stream));
  if (sparse_mask)
  {
    voxelize_kernel<<<(unsigned)pblocks, 256, 0, stream>>>(L, d_xyz, stride, n, -1, nullptr, sparse_mask, d_sparse,
                                                           d_sparse_count);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h_nsparse, d_sparse_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (h_nsparse > 8u * kSparseMax)
      h_nsparse = 8u * kSparseMax;
  }

  {
    const bool need_d2 = (keep_d2 != nullptr) || ndense > 1 || coded != nullptr;

    bool fast_done = false;
    if (ev)
      cudaEventRecord(ev[1], stream);  // end of region A: class count + sparse-site collection
    if (ndense == 1 && keep_d2 == nullptr && fast_passes && coded != nullptr)
      CK(build_fast(map, L, d_xyz, stride, n, G, dense[0], d_sparse, (int)h_nsparse, pblocks, prob, coded, stream, ev_start,
                    ev_end, &fast_done));
    if (fast_done)
      goto done;
    if (need_d2)
      CK(cudaMalloc(&d_d2, sizeof(int32_t) * cells));

    if (ndense == 0)
    {
      sparse_only_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(cells, sx, sy, sub, d_sparse, (int)h_nsparse,
                                                                                G, prob, d_d2);
      CK(cudaGetLastError());
    }
    else
    {
      const uint64_t bits_words = rows * L.words_per_row;
      CK(cudaMalloc(&d_bits, sizeof(uint32_t) * bits_words));
      CK(cudaMalloc(&d_d1, sizeof(uint16_t) * rows * (uint64_t)sx));
      // Fast passes when only the field is wanted (the exact transform is not kept): min(d2, cap) by direct search
      // inside the window the cap allows.  Falls back to the exact envelopes when the caller keeps the EDT, when the
      // probability never reaches 0.0f, or when the window would be wide (large sigma / fine lattice).
      unsigned int h_cap = 0xFFFFFFFFu;
      int win = 0;
      if (keep_d2 == nullptr && fast_passes)
      {
        CK(find_cap(G, &h_cap, stream));
        if (h_cap != 0xFFFFFFFFu && h_cap >= 1 && h_cap < 65535u)
        {
          const int R = (int)floor(sqrt((double)(h_cap - 1)));
          win = R / sub + 1;
          if (win > kWinMaxW)
            win = 0;
        }
      }
      const uint64_t nthreads = 148ull * 1024ull;
      const int maxlen = (L.w[1] > L.w[2] ? L.w[1] : L.w[2]);
      if (win > 0)
        CK(cudaMalloc(&d_dxy16, sizeof(uint16_t) * (uint64_t)sx * sy * L.w[2]));
      else
      {
        CK(cudaMalloc(&d_dxy, sizeof(int32_t) * (uint64_t)sx * sy * L.w[2]));
        CK(cudaMalloc(&d_stack, sizeof(int2) * nthreads * (uint64_t)maxlen));
      }

      for (int k = 0; k < ndense; ++k)
      {
        const int cls = dense[k];
        const int px = cls & 1, py = (cls >> 1) & 1, pz = (cls >> 2) & 1;
        const bool last = (k == ndense - 1);
        CK(cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * bits_words, stream));
        voxelize_kernel<<<(unsigned)pblocks, 256, 0, stream>>>(L, d_xyz, stride, n, cls, d_bits, 0u, nullptr, nullptr);
        CK(cudaGetLastError());
        {
          const uint64_t warps = rows * (uint64_t)((sx + 31) / 32);
          const uint64_t blocks = (warps * 32 + 255) / 256;
          // capped passes: a site farther than the window along x cannot matter, so the bit scan stops there
          const int scan_words = win > 0 ? (win + 31) / 32 + 1 : 0x7FFFFFF;
          xpass_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_bits, L.words_per_row, rows, sx, L.w[0], sub, px, scan_words, d_d1);
          CK(cudaGetLastError());
        }
        FinalArgs F0{};
        PassArgs Ay, Az;
        Ay.ncol = (uint64_t)sx;
        Ay.nplane = (uint32_t)L.w[2];
        Ay.len_in = L.w[1];
        Ay.len_out = sy;
        Ay.sub = sub;
        Ay.par = py;
        Ay.nthreads = nthreads;
        Az.ncol = (uint64_t)sx * sy;
        Az.nplane = 1;
        Az.len_in = L.w[2];
        Az.len_out = sz;
        Az.sub = sub;
        Az.par = pz;
        Az.nthreads = nthreads;
        FinalArgs F{};
        F.enabled = last ? 1 : 0;
        F.prev = (k > 0) ? d_d2 : nullptr;
        F.d2_out = d_d2;  // null only when single class and !keep
        F.prob = prob;
        F.sparse = d_sparse;
        F.nsparse = last ? (int)h_nsparse : 0;
        F.sx = sx;
        F.sy = sy;
        F.g = G;
        if (win > 0)
        {
          const uint64_t by = (uint64_t)Ay.nplane * ((Ay.len_out + kWinTI - 1) / kWinTI) * ((Ay.ncol + kWinTX - 1) / kWinTX);
          const uint64_t bz = (uint64_t)Az.nplane * ((Az.len_out + kWinTI - 1) / kWinTI) * ((Az.ncol + kWinTX - 1) / kWinTX);
          if (by > 0x7FFFFFFFull || bz > 0x7FFFFFFFull)
          {
            rc = cudaErrorInvalidValue;
            goto done;
          }
          window_kernel<false><<<(unsigned)by, 256, 0, stream>>>(d_d1, d_dxy16, Ay, (int32_t)h_cap, win, F0);
          CK(cudaGetLastError());
          window_kernel<true><<<(unsigned)bz, 256, 0, stream>>>(d_dxy16, nullptr, Az, (int32_t)h_cap, win, F);
          CK(cudaGetLastError());
        }
        else
        {
          envelope_kernel<uint16_t, false><<<(unsigned)(nthreads / 256), 256, 0, stream>>>(d_d1, d_dxy, Ay, d_stack, F0);
          CK(cudaGetLastError());
          envelope_kernel<int32_t, true><<<(unsigned)(nthreads / 256), 256, 0, stream>>>(d_dxy, nullptr, Az, d_stack, F);
          CK(cudaGetLastError());
        }
      }
    }
    CK(cudaStreamSynchronize(stream));
    // free the pass temporaries before the coder allocates
    cudaFree(d_bits);
    d_bits = nullptr;
    cudaFree(d_d1);
    d_d1 = nullptr;
    cudaFree(d_dxy);
    d_dxy = nullptr;
    cudaFree(d_dxy16);
    d_dxy16 = nullptr;
    cudaFree(d_stack);
    d_stack = nullptr;
    if (coded != nullptr)
      CK(build_codes(map, d_d2, G, coded, stream));
    if (ev)
      cudaEventRecord(ev[1], stream);  // legacy passes: one region from the first kernel to the last (allocations included)
    if (keep_d2 != nullptr)
    {
      *keep_d2 = d_d2;
      d_d2 = nullptr;
    }
  }

done:
  if (rc != cudaSuccess && err)
    snprintf(err, errlen, "grid build: %s", cudaGetErrorString(rc));
  cudaFree(d_counts);
  cudaFree(d_bits);
  cudaFree(d_sparse);
  cudaFree(d_sparse_count);
  cudaFree(d_d1);
  cudaFree(d_dxy);
  cudaFree(d_dxy16);
  cudaFree(d_d2);
  cudaFree(d_stack);
  return rc;
}

}  // namespace a3d
