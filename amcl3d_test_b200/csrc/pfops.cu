// pfops.cu -- K4/K5/K6: normalise, sequential-f32 chains, resample index search, mean / best (sm_100a)
//
// Stands behind ParticleFilter::particleNormalize   amcl3d/src/ParticleFilter.cpp:168-196
//               ParticleFilter::resample            amcl3d/src/ParticleFilter.cpp:230-254
//               ParticleFilter::uniformSample       amcl3d/src/ParticleFilter.cpp:256-280
//               ParticleFilter::getMean / getBest   amcl3d/src/ParticleFilter.cpp:198-228
//               ParticleFilter::predict             amcl3d/src/ParticleFilter.cpp:91-112
//               MonteCarloLocalization::addParticleWeight  amcl3d/src/amcl3d.cpp:205-217
//
// Bit-exact resampling needs the reference's *sequential* f32 rounding chain for the weight sum, the
// inclusive prefix and the threshold sequence.  The chain kernels run that recurrence on
// one lane while the other warps of the CTA stream the operands through a 3-stage shared-memory ring
// with coalesced loads/stores, so the chain costs one dependent FADD (~4 cycles) per element and no
// exposed memory latency.  Everything that is order-independent (divide, bounds mask, max, the
// lower/upper-bound searches, the particle gather) is embarrassingly parallel.
#include "common.cuh"

#include <float.h>

namespace a3d
{
namespace
{
constexpr int kChainThreads = 256;
constexpr int kChunk = 2048;  // floats per ring stage

enum ChainMode
{
  kChainSum = 0,        // acc += v
  kChainSumPositive = 1,  // acc += (v > 0 ? v : 0)        (ParticleFilter.cpp:174-175)
  kChainThreshold = 2   // out = acc; acc += inc           (ParticleFilter.cpp:267-268,275)
};

// src element k lives at src[k*stride + offset].  prefix (nullable) gets the running value.
__global__ void __launch_bounds__(kChainThreads)
    chain_kernel(const float* __restrict__ src, uint32_t stride, uint32_t offset, uint64_t n, int mode, float start,
                 float inc, float* __restrict__ prefix, float* __restrict__ total)
{
  __shared__ __align__(16) float ring[3][kChunk];
  const uint32_t tid = threadIdx.x;
  const uint64_t nchunks = (n + kChunk - 1) / kChunk;
  float acc = start;

  for (uint64_t c = 0; c < nchunks + 2; ++c)
  {
    if (tid >= 32)
    {
      // loaders: fill stage c, drain stage c-2
      if (c < nchunks && mode != kChainThreshold)
      {
        float* dst = ring[c % 3];
        const uint64_t base = c * kChunk;
        for (uint32_t k = tid - 32; k < kChunk; k += kChainThreads - 32)
        {
          const uint64_t g = base + k;
          dst[k] = (g < n) ? src[g * stride + offset] : 0.f;
        }
      }
      if (c >= 2 && prefix != nullptr)
      {
        const float* from = ring[(c - 2) % 3];
        const uint64_t base = (c - 2) * kChunk;
        for (uint32_t k = tid - 32; k < kChunk; k += kChainThreads - 32)
        {
          const uint64_t g = base + k;
          if (g < n)
            prefix[g] = from[k];
        }
      }
    }
    else if (tid == 0 && c >= 1 && c <= nchunks)
    {
      float4* buf = reinterpret_cast<float4*>(ring[(c - 1) % 3]);
      if (mode == kChainThreshold)
      {
#pragma unroll 4
        for (int k = 0; k < kChunk / 4; ++k)
        {
          float4 o;
          o.x = acc;
          acc = __fadd_rn(acc, inc);
          o.y = acc;
          acc = __fadd_rn(acc, inc);
          o.z = acc;
          acc = __fadd_rn(acc, inc);
          o.w = acc;
          acc = __fadd_rn(acc, inc);
          buf[k] = o;
        }
      }
      else if (mode == kChainSumPositive)
      {
#pragma unroll 4
        for (int k = 0; k < kChunk / 4; ++k)
        {
          float4 v = buf[k];
          acc = __fadd_rn(acc, v.x > 0.f ? v.x : 0.f);
          v.x = acc;
          acc = __fadd_rn(acc, v.y > 0.f ? v.y : 0.f);
          v.y = acc;
          acc = __fadd_rn(acc, v.z > 0.f ? v.z : 0.f);
          v.z = acc;
          acc = __fadd_rn(acc, v.w > 0.f ? v.w : 0.f);
          v.w = acc;
          buf[k] = v;
        }
      }
      else
      {
#pragma unroll 4
        for (int k = 0; k < kChunk / 4; ++k)
        {
          float4 v = buf[k];
          acc = __fadd_rn(acc, v.x);
          v.x = acc;
          acc = __fadd_rn(acc, v.y);
          v.y = acc;
          acc = __fadd_rn(acc, v.z);
          v.z = acc;
          acc = __fadd_rn(acc, v.w);
          v.w = acc;
          buf[k] = v;
        }
      }
    }
    __syncthreads();
  }
  // padding elements beyond n were zeros (sum modes), so acc is the sum over exactly n elements
  if (tid == 0 && total != nullptr)
    total[0] = acc;
}

// ---- exact parallel evaluation of the sequential f32 recurrence ---------------------------------------
// c_i = fl(c_{i-1} + w_i), w_i >= 0.  While the running sum stays inside one binade [2^e, 2^(e+1)) every
// add rounds onto the fixed grid u = 2^(e-23): with c = s*u and w = (k + f)*u the result is
//   s' = s + k            (f < 1/2)        s' = s + k + 1     (f > 1/2)
//   s' = s + k + ((s+k)&1)                 (f == 1/2, ties to even)
// i.e. an integer step that depends on the input only through its parity.  Such steps compose
// associatively as pairs (increment for even input, increment for odd input), so a whole binade is an
// integer prefix scan -- bit-identical to the serial chain.  The first element whose result reaches
// 2^(e+1) (a binade crossing: at most a few dozen per array) is redone with a real f32 add and the scan
// restarts on the coarser grid.  Negative / non-finite operands and denormal sums also take the serial add.
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 16;
constexpr unsigned long long kNoCross = ~0ull;
constexpr int kSat = 1 << 28;

struct StepFn
{
  int d0, d1;  // increment in units of u for an even / odd input state
};
__device__ __forceinline__ StepFn step_compose(StepFn f, StepFn g)  // f first, then g
{
  StepFn r;
  r.d0 = min(f.d0 + ((f.d0 & 1) ? g.d1 : g.d0), kSat);
  r.d1 = min(f.d1 + (((f.d1 + 1) & 1) ? g.d1 : g.d0), kSat);
  return r;
}
// step of one operand on the grid of binade e: k and class (0 round down, 1 round up, 2 tie)
__device__ __forceinline__ void step_classify(float w, int e, int& k, int& cls)
{
  const uint32_t b = __float_as_uint(w);
  k = 0;
  cls = 0;
  if (b == 0u)
    return;
  const uint32_t ew = (b >> 23) & 0xFFu;
  if ((b & 0x80000000u) || ew == 0xFFu)
  {
    k = kSat;  // negative, NaN, Inf: force the serial add
    return;
  }
  if (ew == 0u)
    return;  // denormal operand against a normal sum with e >= -100: far below u/2
  const uint32_t m = (b & 0x7FFFFFu) | 0x800000u;
  const int sh = e - ((int)ew - 127);
  if (sh < 0)
  {
    k = kSat;  // operand above the sum's binade: certain crossing
    return;
  }
  if (sh == 0)
  {
    k = (int)m;
    return;
  }
  if (sh >= 25)
    return;
  k = (int)(m >> sh);
  const uint32_t rem = m & ((1u << sh) - 1u);
  const uint32_t half = 1u << (sh - 1);
  cls = rem > half ? 1 : (rem == half ? 2 : 0);
}
__device__ __forceinline__ int step_apply(int s, int k, int cls)
{
  const int b = min(s + k, kSat);
  return b + (cls == 1 ? 1 : 0) + (cls == 2 ? (b & 1) : 0);
}
__device__ __forceinline__ float state_value(int e, int s)  // s in [2^23, 2^24)
{
  return __uint_as_float(((uint32_t)(e + 127) << 23) | ((uint32_t)s & 0x7FFFFFu));
}

// mode / operands as chain_kernel.  Threshold mode: out[0] = start, out[k] = value after k adds (n = S).
__global__ void __launch_bounds__(kScanThreads)
    exact_scan_kernel(const float* __restrict__ src, uint32_t stride, uint32_t offset, uint64_t n, int mode, float start,
                      float inc, float* __restrict__ prefix, float* __restrict__ total)
{
  __shared__ float sh_c;
  __shared__ unsigned long long sh_i, sh_cross;
  __shared__ int sh_s;
  __shared__ StepFn warp_fn[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // number of adds: thresholds emit S values from S-1 adds
  const uint64_t nadd = (mode == kChainThreshold) ? (n > 0 ? n - 1 : 0) : n;

  auto operand = [&](uint64_t i) -> float {
    if (mode == kChainThreshold)
      return inc;
    const float v = src[i * stride + offset];
    if (mode == kChainSumPositive)
      return v > 0.f ? v : 0.f;
    return v;
  };
  auto emit = [&](uint64_t i, float v) {
    if (prefix == nullptr)
      return;
    if (mode == kChainThreshold)
      prefix[i + 1] = v;
    else
      prefix[i] = v;
  };

  if (tid == 0)
  {
    sh_c = start;
    sh_i = 0;
    if (mode == kChainThreshold && prefix != nullptr && n > 0)
      prefix[0] = start;
  }
  __syncthreads();

  while (true)
  {
    // ---- serial phase: the crossing element (or anything the integer scheme does not cover) ----------
    if (tid == 0)
    {
      uint64_t i = sh_i;
      float c = sh_c;
      while (i < nadd)
      {
        const float w = operand(i);
        const uint32_t cb = __float_as_uint(c);
        const int ce = (int)((cb >> 23) & 0xFFu);
        const bool c_ok = !(cb & 0x80000000u) && ce > 27 && ce < 0xFF;  // positive, normal, e >= -99
        if (c_ok && w >= 0.f)
        {
          const float r = __fadd_rn(c, w);
          if (((__float_as_uint(r) >> 23) & 0xFFu) == (uint32_t)ce)
          {
            // stays in the binade.  Hand over to the scan only if the binade is expected to last a while
            // (room to the next power of two vs the current operand); otherwise a block-wide round would
            // be spent on a handful of elements -- purely a scheduling heuristic, both paths are exact.
            const float top = __uint_as_float((uint32_t)(ce + 1) << 23);
            if (__fsub_rn(top, c) > w * 256.f)
              break;
          }
        }
        c = __fadd_rn(c, w);
        emit(i, c);
        ++i;
      }
      sh_c = c;
      sh_i = i;
    }
    __syncthreads();
    uint64_t base = sh_i;
    const float c0 = sh_c;
    if (base >= nadd)
      break;
    const int e = (int)((__float_as_uint(c0) >> 23) & 0xFFu) - 127;
    int s_chunk = (int)((__float_as_uint(c0) & 0x7FFFFFu) | 0x800000u);

    // ---- parallel phase: chunks of kScanThreads * kScanItems operands inside binade e --------------------
    while (true)
    {
      const uint64_t idx0 = base + (uint64_t)tid * kScanItems;
      int k[kScanItems], cls[kScanItems];
      StepFn f{ 0, 0 };
#pragma unroll
      for (int j = 0; j < kScanItems; ++j)
      {
        k[j] = 0;
        cls[j] = 0;
        if (idx0 + j < nadd)
        {
          step_classify(operand(idx0 + j), e, k[j], cls[j]);
          StepFn g;
          g.d0 = min(k[j] + (cls[j] == 1 ? 1 : 0) + (cls[j] == 2 ? (k[j] & 1) : 0), kSat);
          g.d1 = min(k[j] + (cls[j] == 1 ? 1 : 0) + (cls[j] == 2 ? ((k[j] + 1) & 1) : 0), kSat);
          f = step_compose(f, g);
        }
      }
      // inclusive scan of the per-thread functions across the block
      StepFn incl = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        StepFn p;
        p.d0 = __shfl_up_sync(0xffffffffu, incl.d0, o);
        p.d1 = __shfl_up_sync(0xffffffffu, incl.d1, o);
        if (lane >= o)
          incl = step_compose(p, incl);
      }
      if (lane == 31)
        warp_fn[warp] = incl;
      if (tid == 0)
        sh_cross = kNoCross;
      __syncthreads();
      if (warp == 0)
      {
        StepFn wf = warp_fn[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
          StepFn p;
          p.d0 = __shfl_up_sync(0xffffffffu, wf.d0, o);
          p.d1 = __shfl_up_sync(0xffffffffu, wf.d1, o);
          if (lane >= o)
            wf = step_compose(p, wf);
        }
        warp_fn[lane] = wf;  // inclusive over warps
      }
      __syncthreads();
      // exclusive prefix function of this thread = (warps before) o (lanes before)
      StepFn lane_excl;
      lane_excl.d0 = __shfl_up_sync(0xffffffffu, incl.d0, 1);
      lane_excl.d1 = __shfl_up_sync(0xffffffffu, incl.d1, 1);
      if (lane == 0)
        lane_excl = StepFn{ 0, 0 };
      StepFn excl = lane_excl;
      if (warp > 0)
        excl = step_compose(warp_fn[warp - 1], lane_excl);
      const StepFn block_total = warp_fn[kScanThreads / 32 - 1];

      int s = min(s_chunk + ((s_chunk & 1) ? excl.d1 : excl.d0), kSat);
      unsigned long long my_cross = kNoCross;
      int s_before_cross = 0;
      if (s >= (1 << 24) && idx0 < nadd)
        my_cross = idx0;  // an earlier thread already crossed; its index is smaller and wins the min below
#pragma unroll
      for (int j = 0; j < kScanItems; ++j)
      {
        if (idx0 + j < nadd && my_cross == kNoCross)
        {
          const int s2 = step_apply(s, k[j], cls[j]);
          if (s2 >= (1 << 24))
          {
            my_cross = idx0 + j;
            s_before_cross = s;
          }
          else
          {
            emit(idx0 + j, state_value(e, s2));
            s = s2;
          }
        }
      }
      if (my_cross != kNoCross)
        atomicMin(&sh_cross, my_cross);
      __syncthreads();
      const unsigned long long cross = sh_cross;
      if (cross != kNoCross)
      {
        // the owner of the first crossing knows the state just before it (threads that only inherited an
        // overflowed state have larger indices)
        if (my_cross == cross && s_before_cross != 0)
        {
          sh_c = state_value(e, s_before_cross);
          sh_i = cross;
        }
        __syncthreads();
        break;
      }
      const uint64_t chunk = (uint64_t)kScanThreads * kScanItems;
      s_chunk = s_chunk + ((s_chunk & 1) ? block_total.d1 : block_total.d0);
      base += chunk;
      if (base >= nadd)
      {
        if (tid == 0)
        {
          sh_c = state_value(e, s_chunk);
          sh_i = nadd;
        }
        __syncthreads();
        break;
      }
      __syncthreads();  // warp_fn / sh_cross are rewritten by the next chunk
    }
  }
  if (tid == 0 && total != nullptr)
    total[0] = sh_c;
}

__device__ __forceinline__ bool into_map(const MapDev& map, float x, float y, float z)
{
  return map.has_map && ((double)x >= map.min[0]) && ((double)x < map.max[0]) && ((double)y >= map.min[1]) &&
         ((double)y < map.max[1]) && ((double)z >= map.min[2]) && ((double)z < map.max[2]);
}

// ParticleFilter.cpp:180-195
__global__ void normalize_kernel(const __grid_constant__ MapDev map, float* __restrict__ particles, uint64_t n,
                                 const float* __restrict__ total)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float* p = particles + 7 * i;
  const float wtp = total[0];
  float w = p[6];
  w = (wtp > 0.f) ? __fdiv_rn(w, wtp) : 0.f;
  if (!into_map(map, p[0], p[1], p[2]))
    w = 0.f;
  p[6] = w;
}

__global__ void fuse_range_kernel(float* __restrict__ particles, const float* __restrict__ wr, uint64_t n,
                                  const float* __restrict__ sum_p, const float* __restrict__ sum_r, float alpha)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float wtp = sum_p[0], wtr = sum_r[0];
  const float wp = wtp > 0.f ? __fdiv_rn(particles[7 * i + 6], wtp) : 0.f;
  const float wq = wtr > 0.f ? __fdiv_rn(wr[i], wtr) : 0.f;
  particles[7 * i + 6] = __fadd_rn(__fmul_rn(wp, alpha), __fmul_rn(wq, __fsub_rn(1.f, alpha)));
}

// ---- max / first arg-max ------------------------------------------------------------------------
struct MaxArg
{
  float v;
  unsigned long long i;
};
__device__ __forceinline__ MaxArg better(MaxArg a, MaxArg b)
{
  if (b.v > a.v || (b.v == a.v && b.i < a.i))
    return b;
  return a;
}
constexpr int kRedBlocks = 592;  // 4 x 148 SMs
constexpr int kRedThreads = 256;

__device__ __forceinline__ MaxArg block_reduce_maxarg(MaxArg m)
{
  __shared__ float sv[kRedThreads / 32];
  __shared__ unsigned long long si[kRedThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    MaxArg t;
    t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
    t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
    m = better(m, t);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
  {
    sv[warp] = m.v;
    si[warp] = m.i;
  }
  __syncthreads();
  if (warp == 0)
  {
    MaxArg t;
    t.v = (lane < kRedThreads / 32) ? sv[lane] : 0.f;
    t.i = (lane < kRedThreads / 32) ? si[lane] : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      MaxArg u;
      u.v = __shfl_xor_sync(0xffffffffu, t.v, o);
      u.i = __shfl_xor_sync(0xffffffffu, t.i, o);
      t = better(t, u);
    }
    m = t;
  }
  return m;  // valid in warp 0
}

// pass 1: per-block (max, lowest index attaining it) over w > 0; pass 2 (blocks == 1): final
__global__ void __launch_bounds__(kRedThreads)
    max_pass1(const float* __restrict__ particles, uint64_t n, float* __restrict__ bv, unsigned long long* __restrict__ bi)
{
  MaxArg m{ 0.f, ~0ull };
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const float w = particles[7 * i + 6];
    if (w > m.v)  // strict: keeps the first occurrence within this thread's increasing index walk
    {
      m.v = w;
      m.i = i;
    }
  }
  m = block_reduce_maxarg(m);
  if (threadIdx.x == 0)
  {
    bv[blockIdx.x] = m.v;
    bi[blockIdx.x] = m.i;
  }
}
__global__ void __launch_bounds__(kRedThreads)
    max_pass2(const float* __restrict__ bv, const unsigned long long* __restrict__ bi, int nb, float* __restrict__ out_max,
              unsigned long long* __restrict__ out_arg)
{
  MaxArg m{ 0.f, ~0ull };
  for (int k = threadIdx.x; k < nb; k += blockDim.x)
  {
    MaxArg t{ bv[k], bi[k] };
    if (t.i != ~0ull)
      m = better(m, t);
  }
  m = block_reduce_maxarg(m);
  if (threadIdx.x == 0)
  {
    out_max[0] = m.v;
    out_arg[0] = m.i;
  }
}

// amcl3d.cpp:213-216: float k = 1; if (max_w > 0) k = 1.0 / max_w;  w = k * w
__global__ void scale_kernel(float* __restrict__ particles, uint64_t n, const float* __restrict__ max_w)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float mw = max_w[0];
  float k = 1.f;
  if (mw > 0.f)
    k = __double2float_rn(__ddiv_rn(1.0, (double)mw));
  particles[7 * i + 6] = __fmul_rn(k, particles[7 * i + 6]);
}

// ---- searches -----------------------------------------------------------------------------------
// ParticleFilter.cpp:239-250: i(m) = first i with !(u_m > c_i), clamped to n-1
__global__ void resample_search_kernel(const float* __restrict__ particles, const float* __restrict__ prefix, uint64_t n,
                                       float u01, uint64_t m0, uint64_t m1, float* __restrict__ out,
                                       int32_t* __restrict__ idx)
{
  const uint64_t m = m0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= m1)
    return;
  const float factor = __fdiv_rn(1.f, __ull2float_rn(n));  // 1.f / p_.size()
  const float r = __fmul_rn(factor, u01);                  // factor * rngUniform(0,1)
  const float u = __fadd_rn(r, __fmul_rn(factor, __uint2float_rn((uint32_t)m)));  // uint32_t m -> f32
  uint64_t lo = 0, hi = n;  // first i in [0,n) with prefix[i] >= u
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (u > prefix[mid])
      lo = mid + 1;
    else
      hi = mid;
  }
  const uint64_t src = (lo >= n) ? n - 1 : lo;
  const uint64_t o = m - m0;
  if (out != nullptr)
  {
    const float* p = particles + 7 * src;
    float* q = out + 7 * o;
#pragma unroll
    for (int k = 0; k < 6; ++k)
      q[k] = p[k];
    q[6] = factor;
  }
  if (idx != nullptr)
    idx[o] = (int32_t)src;
}

// ParticleFilter.cpp:266-277: sample k takes the first i with left_sum[i] > samp_k; none -> zero particle
__global__ void uniform_search_kernel(const float* __restrict__ particles, const float* __restrict__ prefix, uint64_t n,
                                      const float* __restrict__ thr, int S, uint64_t k0, uint64_t k1,
                                      float* __restrict__ out, int32_t* __restrict__ idx)
{
  const uint64_t k = k0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= k1)
    return;
  const float inteval = __fdiv_rn(1.0f, __int2float_rn(S));
  const float t = thr[k];
  uint64_t lo = 0, hi = n;  // first i with prefix[i] > t
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (prefix[mid] > t)
      hi = mid;
    else
      lo = mid + 1;
  }
  const uint64_t o = k - k0;
  const bool hit = lo < n;
  if (out != nullptr)
  {
    float* q = out + 7 * o;
    if (hit)
    {
      const float* p = particles + 7 * lo;
#pragma unroll
      for (int c = 0; c < 6; ++c)
        q[c] = p[c];
      q[6] = inteval;
    }
    else
    {
#pragma unroll
      for (int c = 0; c < 7; ++c)
        q[c] = 0.f;
    }
  }
  if (idx != nullptr)
    idx[o] = hit ? (int32_t)lo : -1;
}

// emitted = #{k < S : thr[k] < prefix[n-1]} (thr is non-decreasing)
__global__ void count_emitted_kernel(const float* __restrict__ thr, int S, const float* __restrict__ prefix, uint64_t n,
                                     int* __restrict__ out)
{
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  if (n == 0 || S <= 0)
  {
    out[0] = 0;
    return;
  }
  const float last = prefix[n - 1];
  int lo = 0, hi = S;  // first k with !(last > thr[k])
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (last > thr[mid])
      lo = mid + 1;
    else
      hi = mid;
  }
  out[0] = lo;
}

// ---- mean ---------------------------------------------------------------------------------------
// exact: lanes 0..5 of warp 0 each run one of the six sequential f32 chains mean.c += w * p.c
// (ParticleFilter.cpp:204-209), lane 6 tracks w_max; the other warps stream particles through smem.
constexpr int kMeanChunk = 512;  // particles per ring stage (7 floats each)
__global__ void __launch_bounds__(kChainThreads) mean_exact_kernel(const float* __restrict__ particles, uint64_t n,
                                                                   float* __restrict__ out7)
{
  __shared__ float ring[2][kMeanChunk * 7];
  const uint32_t tid = threadIdx.x;
  const uint64_t nchunks = (n + kMeanChunk - 1) / kMeanChunk;
  float acc = 0.f;  // lanes 0..5: the sum; lane 6: the max
  for (uint64_t c = 0; c < nchunks + 1; ++c)
  {
    if (tid >= 32)
    {
      if (c < nchunks)
      {
        float* dst = ring[c & 1];
        const uint64_t base = c * kMeanChunk * 7;
        const uint64_t lim = n * 7;
        for (uint32_t k = tid - 32; k < kMeanChunk * 7; k += kChainThreads - 32)
          dst[k] = (base + k < lim) ? particles[base + k] : 0.f;
      }
    }
    else if (tid < 7 && c >= 1)
    {
      const float* buf = ring[(c - 1) & 1];
      const uint64_t base = (c - 1) * kMeanChunk;
      const int cntc = (int)((n - base) < (uint64_t)kMeanChunk ? (n - base) : (uint64_t)kMeanChunk);
      if (tid < 6)
      {
#pragma unroll 4
        for (int k = 0; k < cntc; ++k)
          acc = __fadd_rn(acc, __fmul_rn(buf[k * 7 + 6], buf[k * 7 + tid]));
      }
      else
      {
        for (int k = 0; k < cntc; ++k)
        {
          const float w = buf[k * 7 + 6];
          if (w > acc)
            acc = w;
        }
      }
    }
    __syncthreads();
  }
  if (tid < 7)
    out7[tid] = acc;
}

// fast: deterministic f64 tree reduction of the f32 products
__global__ void __launch_bounds__(kRedThreads) mean_pass1(const float* __restrict__ particles, uint64_t n,
                                                          double* __restrict__ part /*[blocks][7]*/)
{
  double s[7] = { 0, 0, 0, 0, 0, 0, 0 };
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const float* p = particles + 7 * i;
    const float w = p[6];
#pragma unroll
    for (int c = 0; c < 6; ++c)
      s[c] += (double)__fmul_rn(w, p[c]);
    s[6] = fmax(s[6], (double)w);
  }
  __shared__ double sh[kRedThreads / 32][7];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < 7; ++c)
  {
    double v = s[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      const double t = __shfl_xor_sync(0xffffffffu, v, o);
      v = (c == 6) ? fmax(v, t) : v + t;
    }
    if (lane == 0)
      sh[warp][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7)
  {
    double v = sh[0][threadIdx.x];
    for (int w2 = 1; w2 < kRedThreads / 32; ++w2)
      v = (threadIdx.x == 6) ? fmax(v, sh[w2][threadIdx.x]) : v + sh[w2][threadIdx.x];
    part[blockIdx.x * 7 + threadIdx.x] = v;
  }
}
__global__ void mean_pass2(const double* __restrict__ part, int nb, float* __restrict__ out7)
{
  const int c = threadIdx.x;
  if (c >= 7)
    return;
  double v = 0;
  for (int k = 0; k < nb; ++k)
    v = (c == 6) ? fmax(v, part[k * 7 + c]) : v + part[k * 7 + c];
  out7[c] = (float)v;
}

// ParticleFilter.cpp:98-111 with the ranGaussian results supplied by the caller
__global__ void predict_kernel(float* __restrict__ particles, uint64_t n, const double* __restrict__ delta,
                               const float* __restrict__ noise)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  float* p = particles + 7 * i;
  const float* g = noise + 6 * i;
  double sd, cd;
  sincos((double)p[5], &sd, &cd);
  const float sa = __double2float_rn(sd), ca = __double2float_rn(cd);
  const float rand_x = __double2float_rn(__dadd_rn(delta[0], (double)g[0]));
  const float rand_y = __double2float_rn(__dadd_rn(delta[1], (double)g[1]));
  p[0] = __fadd_rn(p[0], __fsub_rn(__fmul_rn(ca, rand_x), __fmul_rn(sa, rand_y)));
  p[1] = __fadd_rn(p[1], __fadd_rn(__fmul_rn(sa, rand_x), __fmul_rn(ca, rand_y)));
  p[2] = __double2float_rn(__dadd_rn((double)p[2], __dadd_rn(delta[2], (double)g[2])));
  p[3] = __double2float_rn(__dadd_rn((double)p[3], __dadd_rn(delta[3], (double)g[3])));
  p[4] = __double2float_rn(__dadd_rn((double)p[4], __dadd_rn(delta[4], (double)g[4])));
  p[5] = __double2float_rn(__dadd_rn((double)p[5], __dadd_rn(delta[5], (double)g[5])));
}

inline unsigned int blocks_for(uint64_t n, int threads)
{
  return (unsigned int)((n + threads - 1) / threads);
}
}  // namespace

size_t reduce_scratch_bytes()
{
  // [0,4096): block maxima, [4096,16384): block arg-max, [16384,...): block partial sums (7 doubles)
  return (size_t)16384 + (size_t)kRedBlocks * 7 * sizeof(double) + 256;
}

// weights out of the AoS particle records into a dense array (coalesced operands for the scan)
__global__ void extract_weights_kernel(const float* __restrict__ particles, uint64_t n, float* __restrict__ dense)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    dense[i] = particles[7 * i + 6];
}

// 0: exact parallel scan (default), 1: the serial chain kernel (kept for A/B and as the simple reference)
static int g_chain_mode = 0;
void set_chain_mode(int mode)
{
  g_chain_mode = mode;
}
static cudaError_t launch_chain(const float* src, uint32_t stride, uint32_t offset, uint64_t n, int mode, float start,
                                float inc, float* prefix, float* total, cudaStream_t stream)
{
  if (g_chain_mode == 1)
    chain_kernel<<<1, kChainThreads, 0, stream>>>(src, stride, offset, n, mode, start, inc, prefix, total);
  else
    exact_scan_kernel<<<1, kScanThreads, 0, stream>>>(src, stride, offset, n, mode, start, inc, prefix, total);
  return cudaGetLastError();
}

cudaError_t launch_chain_prefix(const float* particles, uint64_t n, int positive_only, float* prefix, float* total,
                                float* dense_scratch, cudaStream_t stream)
{
  const int mode = positive_only ? kChainSumPositive : kChainSum;
  if (dense_scratch != nullptr && n > 0 && g_chain_mode == 0)
  {
    extract_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(particles, n, dense_scratch);
    return launch_chain(dense_scratch, 1, 0, n, mode, 0.f, 0.f, prefix, total, stream);
  }
  return launch_chain(particles, 7, 6, n, mode, 0.f, 0.f, prefix, total, stream);
}

cudaError_t launch_chain_sum_array(const float* v, uint64_t n, float* total, cudaStream_t stream)
{
  return launch_chain(v, 1, 0, n, kChainSum, 0.f, 0.f, nullptr, total, stream);
}

cudaError_t launch_chain_thresholds(int S, float* thr, cudaStream_t stream)
{
  if (S <= 0)
    return cudaSuccess;
  const float inteval = 1.0f / S;  // ParticleFilter.cpp:267 (IEEE f32 divide on the host)
  const float samp0 = 0.5f / S;    // ParticleFilter.cpp:268
  return launch_chain(nullptr, 1, 0, (uint64_t)S, kChainThreshold, samp0, inteval, thr, nullptr, stream);
}

cudaError_t launch_normalize(const MapDev& map, float* particles, uint64_t n, const float* total, cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  normalize_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(map, particles, n, total);
  return cudaGetLastError();
}

cudaError_t launch_fuse_range(float* particles, const float* wr, uint64_t n, const float* sum_p, const float* sum_r,
                              float alpha, cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  fuse_range_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(particles, wr, n, sum_p, sum_r, alpha);
  return cudaGetLastError();
}

cudaError_t launch_max_weight(const float* particles, uint64_t n, float* out_max, unsigned long long* out_arg,
                              void* scratch, cudaStream_t stream)
{
  float* bv = reinterpret_cast<float*>(scratch);
  unsigned long long* bi = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(scratch) + 4096);
  int nb = (int)blocks_for(n ? n : 1, kRedThreads);
  if (nb > kRedBlocks)
    nb = kRedBlocks;
  max_pass1<<<nb, kRedThreads, 0, stream>>>(particles, n, bv, bi);
  max_pass2<<<1, kRedThreads, 0, stream>>>(bv, bi, nb, out_max, out_arg);
  return cudaGetLastError();
}

cudaError_t launch_scale_by_max(float* particles, uint64_t n, const float* max_w, cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  scale_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(particles, n, max_w);
  return cudaGetLastError();
}

cudaError_t launch_resample_search(const float* particles, const float* prefix, uint64_t n, float u01, uint64_t m0,
                                   uint64_t m1, float* out, int32_t* idx, cudaStream_t stream)
{
  if (m1 <= m0)
    return cudaSuccess;
  resample_search_kernel<<<blocks_for(m1 - m0, 256), 256, 0, stream>>>(particles, prefix, n, u01, m0, m1, out, idx);
  return cudaGetLastError();
}

cudaError_t launch_uniform_search(const float* particles, const float* prefix, uint64_t n, const float* thr, int S,
                                  uint64_t k0, uint64_t k1, float* out, int32_t* idx, cudaStream_t stream)
{
  if (k1 <= k0)
    return cudaSuccess;
  uniform_search_kernel<<<blocks_for(k1 - k0, 256), 256, 0, stream>>>(particles, prefix, n, thr, S, k0, k1, out, idx);
  return cudaGetLastError();
}

cudaError_t launch_count_emitted(const float* thr, int S, const float* prefix, uint64_t n, int* out, cudaStream_t stream)
{
  count_emitted_kernel<<<1, 32, 0, stream>>>(thr, S, prefix, n, out);
  return cudaGetLastError();
}

cudaError_t launch_mean(const float* particles, uint64_t n, int exact, float* out7, void* scratch, cudaStream_t stream)
{
  if (exact)
  {
    mean_exact_kernel<<<1, kChainThreads, 0, stream>>>(particles, n, out7);
    return cudaGetLastError();
  }
  double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 16384);
  int nb = (int)blocks_for(n ? n : 1, kRedThreads);
  if (nb > kRedBlocks)
    nb = kRedBlocks;
  mean_pass1<<<nb, kRedThreads, 0, stream>>>(particles, n, part);
  mean_pass2<<<1, 32, 0, stream>>>(part, nb, out7);
  return cudaGetLastError();
}

cudaError_t launch_predict(float* particles, uint64_t n, const double* delta6_dev, const float* noise6,
                           cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  predict_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(particles, n, delta6_dev, noise6);
  return cudaGetLastError();
}

}  // namespace a3d
