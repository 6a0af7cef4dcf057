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
constexpr int kScanItems = 8;
constexpr float kHandOver = 128.f;  // serial run-up while fewer than about this many adds remain in the binade
// operands / results of one chunk pass through shared memory so that global traffic is coalesced while every
// thread still owns kScanItems consecutive elements; one pad word per kScanItems keeps both views conflict-free
__device__ __forceinline__ int stage_slot(int i)
{
  return i + (i >> 3);
}
constexpr int kStageWords = kScanThreads * kScanItems + kScanThreads;
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
// Step of one operand on the grid of binade e, branch-free.  With w = (k + f)*u the code packs
//   k (bits 2..), h = [f >= 1/2] (bit 1), st = [f is neither 0 nor 1/2] (bit 0);
// the add then is  s' = s + k + (h & (st | (s + k))), the last term being round-to-nearest-even.
// Negative / non-finite operands and operands above the binade get kBadCode (k saturates: a forced "crossing"
// that hands the element to the serial add).  Zeros and denormals (far below u/2 for e >= -99) code as 0.
constexpr uint32_t kBadCode = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t step_code(uint32_t b, int e)
{
  const uint32_t ew = (b >> 23) & 0xFFu;
  const int sh = e + 127 - (int)ew;
  const uint32_t m2 = ew ? (((b & 0x7FFFFFu) | 0x800000u) << 1) : 0u;
  const uint32_t shc = (uint32_t)min(max(sh, 0), 31);
  const uint32_t t = m2 >> shc;
  const uint32_t low = m2 & ((1u << shc) - 1u);
  const bool bad = (b & 0x80000000u) || ew == 0xFFu || sh < 0;
  return bad ? kBadCode : ((t << 1) | (low != 0u ? 1u : 0u));
}
__device__ __forceinline__ int step_apply(int s, uint32_t code)
{
  const int t = min(s + (int)(code >> 2), kSat);
  return t + (int)((code >> 1) & (code | (uint32_t)t) & 1u);
}
// the StepFn of a run of codes: follow an even and an odd virtual state through it
struct StepPair
{
  int a = 0, b = 1;
  __device__ __forceinline__ void push(uint32_t code)
  {
    a = step_apply(a, code);
    b = step_apply(b, code);
  }
  __device__ __forceinline__ StepFn fn() const
  {
    return StepFn{ min(a, kSat), min(b - 1, kSat) };
  }
};
__device__ __forceinline__ float state_value(int e, int s)  // s in [2^23, 2^24)
{
  return __uint_as_float(((uint32_t)(e + 127) << 23) | ((uint32_t)s & 0x7FFFFFu));
}

// The operands of one chain: element i lives at src[i*stride + offset] (sum modes) or is the constant inc
// (threshold mode, where the value after add i goes to prefix[i+1]).
struct ChainOps
{
  const float* src;
  uint32_t stride, offset;
  int mode;
  float inc;
  float* prefix;
  __device__ __forceinline__ float operand(uint64_t i) const
  {
    if (mode == kChainThreshold)
      return inc;
    const float v = src[i * stride + offset];
    if (mode == kChainSumPositive)
      return v > 0.f ? v : 0.f;
    return v;
  }
  __device__ __forceinline__ void emit(uint64_t i, float v) const
  {
    if (prefix == nullptr)
      return;
    if (mode == kChainThreshold)
      prefix[i + 1] = v;
    else
      prefix[i] = v;
  }
};

// Block-wide (kScanThreads threads): performs adds [begin, nadd) of the chain starting from state `start`,
// emits the running values and returns the final state to every thread.
//
// One chunk of kScanThreads*kScanItems operands at a time is staged in shared memory (coalesced), turned into
// running values in place, and written back coalesced.  Inside a chunk three kinds of pass alternate until every
// slot holds its result:
//   serial  lane 0 adds element by element (real f32 adds, operands from shared memory) -- the crossing element,
//           anything the integer scheme does not cover, and the run-up while the binade is about to end anyway;
//   zero    while the state is +0, leading +0 operands leave it there: find the first other operand in parallel;
//   binade  the integer step scan of the remaining slots on the grid of the current binade, up to the first
//           element whose result leaves it.
__device__ float scan_range(const ChainOps& ops, uint64_t begin, uint64_t nadd, float start, float* stage /* [kStageWords] shared */)
{
  __shared__ float sh_c;
  __shared__ int sh_pos, sh_cross;
  __shared__ StepFn warp_fn[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kChunkLen = kScanThreads * kScanItems;
  constexpr int kNone = 0x7fffffff;

  float c = start;
  for (uint64_t base = begin; base < nadd; base += kChunkLen)
  {
    const int len = (int)min((uint64_t)kChunkLen, nadd - base);
    __syncthreads();  // the previous chunk's (or call's) readers are done with stage / sh_*
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
    {
      const int i = j * kScanThreads + tid;
      if (i < len)
        stage[stage_slot(i)] = ops.operand(base + i);
    }
    if (tid == 0)
    {
      sh_pos = 0;
      sh_c = c;
    }
    __syncthreads();

    while (true)
    {
      // ---- serial pass ---------------------------------------------------------------------------------
      if (tid == 0)
      {
        int i = sh_pos;
        float cc = sh_c;
        while (i < len)
        {
          const float w = stage[stage_slot(i)];
          const uint32_t cb = __float_as_uint(cc);
          const int ce = (int)((cb >> 23) & 0xFFu);
          const bool c_ok = !(cb & 0x80000000u) && ce > 27 && ce < 0xFF;  // positive, normal, e >= -99
          if (c_ok && w >= 0.f)
          {
            const float top = __uint_as_float((uint32_t)(ce + 1) << 23);
            if (__fadd_rn(cc, w) < top)
            {
              // stays in the binade.  Hand over to the scan only if the binade is expected to last a while
              // (room to the next power of two vs the current operand); otherwise a block-wide pass would
              // be spent on a handful of elements -- purely a scheduling heuristic, both paths are exact.
              if (__fsub_rn(top, cc) > w * kHandOver)
                break;
              // run-up to the end of the binade: plain adds until the sum would reach `top` (or an operand
              // needs the general step below), then look again
              const int limit = min(len, i + 2 * (int)kHandOver);
              for (; i < limit; ++i)
              {
                const float v = stage[stage_slot(i)];
                const float r = __fadd_rn(cc, v);
                if (!(v >= 0.f) || !(r < top))
                  break;
                cc = r;
                stage[stage_slot(i)] = r;
              }
              continue;
            }
          }
          else if (cb == 0u && __float_as_uint(w) == 0u)
            break;  // zero pass
          cc = __fadd_rn(cc, w);
          stage[stage_slot(i)] = cc;
          ++i;
        }
        sh_pos = i;
        sh_c = cc;
        sh_cross = kNone;
      }
      __syncthreads();
      const int pos = sh_pos;
      const float c0 = sh_c;
      if (pos >= len)
        break;
      const int i0 = tid * kScanItems;

      if (__float_as_uint(c0) == 0u)
      {
        // ---- zero pass: slots [pos, first other operand) already hold their result (+0) ----------------
        int first = kNone;
#pragma unroll
        for (int j = kScanItems - 1; j >= 0; --j)
          if (i0 + j >= pos && i0 + j < len && __float_as_uint(stage[stage_slot(i0 + j)]) != 0u)
            first = i0 + j;
        if (first != kNone)
          atomicMin(&sh_cross, first);
        __syncthreads();
        if (tid == 0)
          sh_pos = min(sh_cross, len);
        __syncthreads();
        continue;
      }

      // ---- binade pass ---------------------------------------------------------------------------------
      const int e = (int)((__float_as_uint(c0) >> 23) & 0xFFu) - 127;
      const int s0 = (int)((__float_as_uint(c0) & 0x7FFFFFu) | 0x800000u);
      uint32_t code[kScanItems];
      StepPair run;
#pragma unroll
      for (int j = 0; j < kScanItems; ++j)
      {
        const bool live = i0 + j >= pos && i0 + j < len;
        code[j] = live ? step_code(__float_as_uint(stage[stage_slot(i0 + j)]), e) : 0u;
        run.push(code[j]);
      }
      const StepFn f = run.fn();
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

      int s = min(s0 + ((s0 & 1) ? excl.d1 : excl.d0), kSat);
      int my_cross = kNone, s_before_cross = 0;
      if (s >= (1 << 24) && i0 + kScanItems > pos && i0 < len)
        my_cross = max(i0, pos);  // an earlier thread already crossed; its index is smaller and wins the min below
#pragma unroll
      for (int j = 0; j < kScanItems; ++j)
      {
        if (i0 + j >= pos && i0 + j < len && my_cross == kNone)
        {
          const int s2 = step_apply(s, code[j]);
          if (s2 >= (1 << 24))
          {
            my_cross = i0 + j;
            s_before_cross = s;
          }
          else
          {
            stage[stage_slot(i0 + j)] = state_value(e, s2);  // own slot: nobody else reads it
            s = s2;
          }
        }
      }
      if (my_cross != kNone)
        atomicMin(&sh_cross, my_cross);
      __syncthreads();
      const int cross = sh_cross;
      if (cross != kNone)
      {
        // the owner of the first crossing knows the state just before it (threads that only inherited an
        // overflowed state have larger indices); lane 0 redoes that element with a real add
        if (my_cross == cross && s_before_cross != 0)
        {
          sh_c = state_value(e, s_before_cross);
          sh_pos = cross;
        }
      }
      else if (tid == 0)
      {
        sh_c = state_value(e, s0 + ((s0 & 1) ? block_total.d1 : block_total.d0));
        sh_pos = len;
      }
      __syncthreads();
    }
    c = sh_c;
    if (ops.prefix != nullptr)
    {
#pragma unroll
      for (int j = 0; j < kScanItems; ++j)
      {
        const int i = j * kScanThreads + tid;
        if (i < len)
          ops.emit(base + i, stage[stage_slot(i)]);
      }
    }
  }
  __syncthreads();
  return c;
}

// single-CTA form (chain mode 2, and arrays of a tile or two).  Threshold mode: out[0] = start, out[k] = value
// after k adds (n = S).
__global__ void __launch_bounds__(kScanThreads)
    exact_scan_kernel(const float* __restrict__ src, uint32_t stride, uint32_t offset, uint64_t n, int mode, float start,
                      float inc, float* __restrict__ prefix, float* __restrict__ total)
{
  const ChainOps ops{ src, stride, offset, mode, inc, prefix };
  const uint64_t nadd = (mode == kChainThreshold) ? (n > 0 ? n - 1 : 0) : n;  // S thresholds from S-1 adds
  if (threadIdx.x == 0 && mode == kChainThreshold && prefix != nullptr && n > 0)
    prefix[0] = start;
  __shared__ float stage[kStageWords];
  const float c = scan_range(ops, 0, nadd, start, stage);
  if (threadIdx.x == 0 && total != nullptr)
    total[0] = c;
}

// ---- the same recurrence over many CTAs -----------------------------------------------------------------
// A run of operands acts on the incoming state as one StepFn as long as the state stays inside a binade whose
// exponent is known.  tile_sum / tile_fn evaluate that function in parallel for every sub-tile of kSubLen operands
// (one warp each) and, composed, for every tile of kTile operands -- for the three binades around a guess (the
// f64 sum of everything before the sub-tile).  resolve_kernel then walks the tiles serially in O(1) each; in a
// tile where that fails (the state leaves the binade inside it: a few dozen tiles per array) it walks the
// sub-tiles in O(1) each and redoes the one that holds the crossing with real f32 adds.  emit_kernel finally
// derives every sub-tile's exact incoming state the same way and turns the operands into running values with
// one warp-wide integer scan per sub-tile.  Nothing here is approximate: a guess that misses only costs time
// (serial adds, or the block-wide scan_range when a whole tile is off).
constexpr uint64_t kTile = (uint64_t)kScanThreads * kScanItems;
constexpr int kSubTiles = kScanThreads / 32;
constexpr int kSubLen = 32 * kScanItems;
constexpr int kCand = 3;
constexpr int kNoGuess = -100000;
constexpr int kMaxSerialSubTiles = 8;
struct __align__(16) TileFn
{
  int g;         // candidate c covers binade g - 1 + c; kNoGuess: look at the sub-tiles
  int identity;  // every operand is +0: the state passes through whatever it is
  int d[kCand][2];
};
struct __align__(16) TileSubFn
{
  StepFn f[kCand][kSubTiles];
  int g[kSubTiles];
};
constexpr int kSubWords = sizeof(TileSubFn) / 4;

__global__ void __launch_bounds__(kScanThreads) tile_sum_kernel(ChainOps ops, uint64_t nadd, double* __restrict__ tile_sum)
{
  const uint64_t begin = (uint64_t)blockIdx.x * kTile;
  double s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    const uint64_t i = begin + (uint64_t)j * kScanThreads + threadIdx.x;
    if (i < nadd)
      s += (double)ops.operand(i);
  }
  __shared__ double sh[kScanThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0)
    sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double v = 0;
    for (int w = 0; w < kScanThreads / 32; ++w)
      v += sh[w];
    tile_sum[blockIdx.x] = v;
  }
}

__device__ __forceinline__ int guess_binade(double v)
{
  const uint32_t gb = __float_as_uint((float)v);
  const int ge = (int)((gb >> 23) & 0xFFu);
  return (!(gb & 0x80000000u) && ge > 28 && ge < 0xFE) ? ge - 127 : kNoGuess;
}

__global__ void __launch_bounds__(kScanThreads)
    tile_fn_kernel(ChainOps ops, uint64_t nadd, float start, const double* __restrict__ tile_sum, TileFn* __restrict__ fns,
                   TileSubFn* __restrict__ subs)
{
  __shared__ double sh_red[kSubTiles], sh_wsum[kSubTiles];
  __shared__ int sh_g[kSubTiles], sh_nonzero;
  __shared__ StepFn sh_fn[kCand][kSubTiles];
  __shared__ float stage[kStageWords];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t t = blockIdx.x, begin = t * kTile;
  int nonzero = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    const int i = j * kScanThreads + tid;
    const float w = (begin + i < nadd) ? ops.operand(begin + i) : 0.f;
    nonzero |= (__float_as_uint(w) != 0u);
    stage[stage_slot(i)] = w;
  }
  // f64 sum of everything before this tile
  double before = 0;
  if (ops.mode == kChainThreshold)
    before = (tid == 0) ? (double)ops.inc * (double)begin : 0.0;
  else
    for (uint64_t k = tid; k < t; k += kScanThreads)
      before += tile_sum[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0)
    sh_red[warp] = before;
  if (tid == 0)
    sh_nonzero = 0;
  __syncthreads();
  // f64 sum of each sub-tile (warp w owns operands [w*kSubLen, (w+1)*kSubLen) of the tile)
  uint32_t wb[kScanItems];
  double mine = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    const float w = stage[stage_slot(tid * kScanItems + j)];
    wb[j] = __float_as_uint(w);
    mine += (double)w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0)
    sh_wsum[warp] = mine;
  if (nonzero)
    sh_nonzero = 1;
  __syncthreads();
  if (warp == 0)
  {
    // guess per sub-tile: the binade of (start + everything before it) evaluated in f64
    double base = (double)start;
    for (int w = 0; w < kSubTiles; ++w)
      base += sh_red[w];
    const double v = sh_wsum[lane];
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const double p = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o)
        incl += p;
    }
    sh_g[lane] = guess_binade(base + (incl - v));
  }
  __syncthreads();
  const int g = sh_g[warp];

  StepPair run[kCand];
  if (g != kNoGuess)
  {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
    {
#pragma unroll
      for (int c = 0; c < kCand; ++c)
        run[c].push(step_code(wb[j], g - 1 + c));  // slots beyond nadd hold +0: no step
    }
  }
  // ordered reduction: lane i ends with the composition of lanes [i, i + 2o)
#pragma unroll
  for (int c = 0; c < kCand; ++c)
  {
    StepFn v = run[c].fn();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      StepFn r;
      r.d0 = __shfl_down_sync(0xffffffffu, v.d0, o);
      r.d1 = __shfl_down_sync(0xffffffffu, v.d1, o);
      if (lane + o < 32)
        v = step_compose(v, r);
    }
    if (lane == 0)
    {
      sh_fn[c][warp] = v;
      subs[t].f[c][warp] = v;
    }
  }
  if (lane == 0)
    subs[t].g[warp] = g;
  __syncthreads();
  if (warp == 0)
  {
    // the tile as a whole has a function only if all its sub-tiles were evaluated for the same binades
    const int g0 = sh_g[0];
    const bool uniform = __all_sync(0xffffffffu, sh_g[lane] == g0) && g0 != kNoGuess;
    TileFn out;
    out.g = uniform ? g0 : kNoGuess;
    out.identity = sh_nonzero ? 0 : 1;
#pragma unroll
    for (int c = 0; c < kCand; ++c)
    {
      StepFn v = sh_fn[c][lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        StepFn r;
        r.d0 = __shfl_down_sync(0xffffffffu, v.d0, o);
        r.d1 = __shfl_down_sync(0xffffffffu, v.d1, o);
        if (lane + o < 32)
          v = step_compose(v, r);
      }
      out.d[c][0] = v.d0;
      out.d[c][1] = v.d1;
    }
    if (lane == 0)
      fns[t] = out;
  }
}

// Warp-cooperative application of up to 32 consecutive runs to the state c.  Lane v holds run v: the binade guess
// g its three candidate functions d were evaluated around, and whether it is all +0 (identity for ANY state).
// Runs [u, m) are live.  Returns the first live run that cannot be applied in O(1) -- the state would leave its
// binade inside it, an operand needs the real add, or no candidate covers the state's binade -- or 32 if all
// went through.  `before` (per lane) = state in front of run v, valid for u <= v <= min(returned, m - 1);
// c_out (uniform) = state in front of the returned run, or behind the last live run.
__device__ __forceinline__ int warp_walk(float c, int u, int m, int g, int d00, int d01, int d10, int d11, int d20, int d21,
                                         bool identity, float& before, float& c_out)
{
  const int lane = threadIdx.x & 31;
  const uint32_t cb = __float_as_uint(c);
  const int ce = (int)((cb >> 23) & 0xFFu);
  const bool c_ok = !(cb & 0x80000000u) && ce > 27 && ce < 0xFF;  // positive, normal, e >= -99
  const int s = (int)((cb & 0x7FFFFFu) | 0x800000u);
  StepFn f{ 0, 0 };
  if (lane >= u && lane < m && !identity)
  {
    const int idx = (ce - 127) - (g - 1);
    if (c_ok && g != kNoGuess && idx >= 0 && idx < kCand)
      f = StepFn{ idx == 0 ? d00 : (idx == 1 ? d10 : d20), idx == 0 ? d01 : (idx == 1 ? d11 : d21) };
    else
      f = StepFn{ kSat, kSat };
  }
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
  StepFn excl;
  excl.d0 = __shfl_up_sync(0xffffffffu, incl.d0, 1);
  excl.d1 = __shfl_up_sync(0xffffffffu, incl.d1, 1);
  if (lane == 0)
    excl = StepFn{ 0, 0 };
  const int d_in = (s & 1) ? incl.d1 : incl.d0;
  const int d_ex = (s & 1) ? excl.d1 : excl.d0;
  // d == 0 leaves the state as it is whatever it is (runs of +0 in front of a zero / denormal state included)
  const bool ok_in = d_in == 0 || (d_in < kSat && s + d_in < (1 << 24));
  const uint32_t hi = cb & 0x7F800000u;
  before = d_ex == 0 ? c : __uint_as_float(hi | ((uint32_t)(s + d_ex) & 0x7FFFFFu));
  const float after = d_in == 0 ? c : __uint_as_float(hi | ((uint32_t)(s + d_in) & 0x7FFFFFu));
  const uint32_t fail = __ballot_sync(0xffffffffu, !ok_in);  // increments only grow: once a lane fails all later do
  const int ff = fail ? __ffs(fail) - 1 : 32;
  c_out = ff < 32 ? __shfl_sync(0xffffffffu, before, ff) : __shfl_sync(0xffffffffu, after, 31);
  return ff;
}

// real f32 adds over operands [i0, i1) of a staged tile (i0 a multiple of 8); keep: the running values replace them
__device__ __forceinline__ float serial_adds(float* stage, int i0, int i1, bool keep, float c)
{
  float cc = c;
  for (int i = i0; i < i1; i += 8)  // a multiple of 8 operands share one pad-free run of slots
  {
    float v[8];
    const int slot = stage_slot(i);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = (i + k < i1) ? stage[slot + k] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      if (i + k < i1)
        cc = __fadd_rn(cc, v[k]);
      v[k] = cc;
    }
    if (keep)
    {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (i + k < i1)
          stage[slot + k] = v[k];
    }
  }
  return cc;
}

// Warp 0 walks the sub-tiles of a staged tile from state c.  Sub-tiles whose function applies cost O(1) for the
// whole run of them (one warp scan); a sub-tile where it does not is redone with real adds by lane 0 (keep: the
// running values replace its operands in `stage` and it is marked finished).  sub_state / sub_done (nullable,
// shared memory) receive every sub-tile's incoming state.  Returns the sub-tile where the walk gave up (more than
// kMaxSerialSubTiles needed the adds: the caller switches to the block-wide scan from there), or kSubTiles.
__device__ int walk_sub_tiles(const TileSubFn& sub, float* stage, int len, bool keep, float& c, float* sub_state, int* sub_done)
{
  const int lane = threadIdx.x & 31;
  const int nsub = (len + kSubLen - 1) / kSubLen;
  const int g = sub.g[lane];
  const StepFn f0 = sub.f[0][lane], f1 = sub.f[1][lane], f2 = sub.f[2][lane];
  if (sub_done)
    sub_done[lane] = 0;
  int serial_left = kMaxSerialSubTiles;
  int su = 0;
  while (su < nsub)
  {
    float before, c_out;
    const int ff = warp_walk(c, su, nsub, g, f0.d0, f0.d1, f1.d0, f1.d1, f2.d0, f2.d1, false, before, c_out);
    if (sub_state && lane >= su && lane <= min(ff, nsub - 1))
      sub_state[lane] = before;
    c = c_out;
    if (ff >= nsub)
      break;
    if (serial_left-- == 0)
      return ff;
    float cc = c;
    if (lane == 0)
    {
      cc = serial_adds(stage, ff * kSubLen, min((ff + 1) * kSubLen, len), keep, c);
      if (sub_done)
        sub_done[ff] = 1;
    }
    c = __shfl_sync(0xffffffffu, cc, 0);
    su = ff + 1;
  }
  return kSubTiles;
}

__device__ __forceinline__ void stage_tile(const ChainOps& ops, uint64_t tbeg, int len, const TileSubFn* src, float* stage,
                                           TileSubFn* sh_sub)
{
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    const int i = j * kScanThreads + tid;
    if (i < len)
      stage[stage_slot(i)] = ops.operand(tbeg + i);
  }
  if (tid < kSubWords)
    reinterpret_cast<int*>(sh_sub)[tid] = reinterpret_cast<const int*>(src)[tid];
}

__global__ void __launch_bounds__(kScanThreads)
    resolve_kernel(ChainOps ops, uint64_t nadd, float start, const TileFn* __restrict__ fns,
                   const TileSubFn* __restrict__ subs, uint64_t ntiles, float* __restrict__ state_in,
                   float* __restrict__ total)
{
  __shared__ TileSubFn sh_sub;
  __shared__ float stage[kStageWords];
  __shared__ float sh_state;
  __shared__ int sh_ff, sh_u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ChainOps quiet = ops;  // the prefix is written by emit_kernel
  quiet.prefix = nullptr;
  float c = start;
  if (tid == 0 && ops.mode == kChainThreshold && ops.prefix != nullptr)
    ops.prefix[0] = start;
  // 32 tiles at a time, one per lane of warp 0; the next group's functions are fetched while this one is walked
  TileFn cur{}, nxt{};
  if (warp == 0 && (uint64_t)lane < ntiles)
    cur = fns[lane];
  for (uint64_t tb = 0; tb < ntiles; tb += 32)
  {
    const int m = (int)min((uint64_t)32, ntiles - tb);
    if (warp == 0 && tb + 32 + lane < ntiles)
      nxt = fns[tb + 32 + lane];
    int u = 0;
    while (true)  // block-uniform
    {
      if (warp == 0)
      {
        float before, c_out;
        const int ff = warp_walk(c, u, m, cur.g, cur.d[0][0], cur.d[0][1], cur.d[1][0], cur.d[1][1], cur.d[2][0],
                                 cur.d[2][1], cur.identity != 0, before, c_out);
        if (lane >= u && lane <= min(ff, m - 1))
          state_in[tb + lane] = before;
        if (lane == 0)
        {
          sh_ff = ff;
          sh_state = c_out;
        }
      }
      __syncthreads();
      const int ff = sh_ff;
      c = sh_state;
      if (ff >= m)
        break;
      // ---- tile tb + ff needs a closer look: stage it and walk its sub-tiles ----------------------------
      const uint64_t tbeg = (tb + ff) * kTile;
      const int len = (int)min(kTile, nadd - tbeg);
      stage_tile(ops, tbeg, len, subs + tb + ff, stage, &sh_sub);
      __syncthreads();
      if (warp == 0)
      {
        const int stop = walk_sub_tiles(sh_sub, stage, len, false, c, nullptr, nullptr);
        if (lane == 0)
        {
          sh_u = stop;
          sh_state = c;
        }
      }
      __syncthreads();
      c = sh_state;
      const int stop = sh_u;
      if (stop < kSubTiles)
        c = scan_range(quiet, tbeg + (uint64_t)stop * kSubLen, tbeg + len, c, stage);  // the guess is of no use here
      u = ff + 1;
      __syncthreads();
    }
    cur = nxt;
  }
  if (tid == 0 && total != nullptr)
    total[0] = c;
}

__global__ void __launch_bounds__(kScanThreads)
    emit_kernel(ChainOps ops, uint64_t nadd, const TileFn* __restrict__ fns, const TileSubFn* __restrict__ subs,
                const float* __restrict__ state_in)
{
  __shared__ TileSubFn sh_sub;
  __shared__ float stage[kStageWords];
  __shared__ float sub_state[kSubTiles];
  __shared__ int sub_done[kSubTiles];
  __shared__ int sh_u;
  __shared__ float sh_state;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t t = blockIdx.x, tbeg = t * kTile;
  const int len = (int)min(kTile, nadd - tbeg);
  float c = state_in[t];
  if (fns[t].identity)
  {
    // a tile of +0 operands: every running value is the incoming state
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
    {
      const int i = j * kScanThreads + tid;
      if (i < len)
        ops.emit(tbeg + i, c);
    }
    return;
  }
  stage_tile(ops, tbeg, len, subs + t, stage, &sh_sub);
  __syncthreads();
  if (warp == 0)
  {
    const int stop = walk_sub_tiles(sh_sub, stage, len, true, c, sub_state, sub_done);
    if (lane == 0)
    {
      sh_u = stop;
      sh_state = c;
    }
  }
  __syncthreads();
  const int u_stop = sh_u;
  // one warp per sub-tile: the state provably stays inside its binade, so a single integer scan gives all values
  if (warp < u_stop && warp * kSubLen < len && !sub_done[warp])
  {
    const float c0 = sub_state[warp];
    const int e = (int)((__float_as_uint(c0) >> 23) & 0xFFu) - 127;
    const int s0 = (int)((__float_as_uint(c0) & 0x7FFFFFu) | 0x800000u);
    const int i0 = tid * kScanItems;
    uint32_t code[kScanItems];
    StepPair run;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
    {
      code[j] = (i0 + j < len) ? step_code(__float_as_uint(stage[stage_slot(i0 + j)]), e) : 0u;
      run.push(code[j]);
    }
    StepFn incl = run.fn();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      StepFn p;
      p.d0 = __shfl_up_sync(0xffffffffu, incl.d0, o);
      p.d1 = __shfl_up_sync(0xffffffffu, incl.d1, o);
      if (lane >= o)
        incl = step_compose(p, incl);
    }
    StepFn excl;
    excl.d0 = __shfl_up_sync(0xffffffffu, incl.d0, 1);
    excl.d1 = __shfl_up_sync(0xffffffffu, incl.d1, 1);
    if (lane == 0)
      excl = StepFn{ 0, 0 };
    int s = s0 + ((s0 & 1) ? excl.d1 : excl.d0);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
    {
      s = step_apply(s, code[j]);
      if (i0 + j < len)
        stage[stage_slot(i0 + j)] = state_value(e, s);
    }
  }
  __syncthreads();
  const int done_len = min(u_stop * kSubLen, len);
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
  {
    const int i = j * kScanThreads + tid;
    if (i < done_len)
      ops.emit(tbeg + i, stage[stage_slot(i)]);
  }
  if (u_stop < kSubTiles)
    scan_range(ops, tbeg + (uint64_t)u_stop * kSubLen, tbeg + len, sh_state, stage);  // the guess is of no use here
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
// (ParticleFilter.cpp:204-209), lane 6 tracks w_max.  The particle records arrive in a two-stage shared-memory ring by
// TMA bulk copies; the other seven warps turn every stage into component-major rows of the ROUNDED products w * p.c (the
// reference rounds the product before the add: no FMA on baseline x86-64) one chunk ahead of the chain warp, whose step
// is then a quarter of a conflict-free LDS.128 and one dependent FADD.
constexpr int kMeanChunk = 384;        // particles per stage: 384 x 28 B = 10752 B (a multiple of 16, as TMA wants)
constexpr int kMeanPitch = kMeanChunk + 4;  // floats between product rows: 16-byte aligned, 4 banks apart
__global__ void __launch_bounds__(kChainThreads) mean_exact_kernel(const float* __restrict__ particles, uint64_t n,
                                                                   const float* carry7, float* out7)
{
  __shared__ __align__(128) float raw[2][kMeanChunk * 7];
  __shared__ __align__(16) float prod[2][7][kMeanPitch];
  __shared__ __align__(8) uint64_t full_bar[2];
  const uint32_t tid = threadIdx.x;
  // bulk copies move multiples of 16 bytes = 4 records; up to three records at the end are read directly
  const bool tma_ok = (reinterpret_cast<uintptr_t>(particles) & 15u) == 0;
  const uint64_t n4 = tma_ok ? (n & ~(uint64_t)3) : 0;
  const uint64_t nchunks = (n4 + kMeanChunk - 1) / kMeanChunk;
  if (tid == 0)
  {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto count_of = [&](uint64_t c) { return (uint32_t)((n4 - c * kMeanChunk) < (uint64_t)kMeanChunk ? (n4 - c * kMeanChunk) : (uint64_t)kMeanChunk); };
  auto issue = [&](uint64_t c) {
    const uint32_t s = (uint32_t)(c & 1u);
    mbar_expect_tx(&full_bar[s], count_of(c) * 28u);
    tma_bulk_g2s(&raw[s][0], particles + c * kMeanChunk * 7, count_of(c) * 28u, &full_bar[s]);
  };
  if (tid == 0)
    for (uint64_t c = 0; c < nchunks && c < 2; ++c)
      issue(c);
  // lanes 0..5: the sum; lane 6: the max.  carry7: the state the chains had after the particles before this block (a set
  // sharded over several handles: the previous rank's out7, possibly peer memory), NULL = start of the set
  float acc = (carry7 != nullptr && tid < 7) ? carry7[tid] : 0.f;
  for (uint64_t c = 0; c < nchunks + 1; ++c)
  {
    if (tid >= 32)
    {
      if (c < nchunks)  // products of chunk c, one chunk ahead of the chain
      {
        const uint32_t s = (uint32_t)(c & 1u), cnt = count_of(c);
        mbar_wait(&full_bar[s], (uint32_t)((c >> 1) & 1u));
        for (uint32_t k = tid - 32; k < cnt; k += kChainThreads - 32)
        {
          const float* p = &raw[s][k * 7];
          const float w = p[6];
#pragma unroll
          for (int e = 0; e < 6; ++e)
            prod[s][e][k] = __fmul_rn(w, p[e]);
          prod[s][6][k] = w;
        }
      }
    }
    else if (tid < 7 && c >= 1)
    {
      const uint32_t s = (uint32_t)((c - 1) & 1u);
      const int cnt = (int)count_of(c - 1);
      const float* row = prod[s][tid];
      int k = 0;
      if (tid < 6)
      {
        for (; k + 16 <= cnt; k += 16)
        {
          float4 t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            t[u] = *reinterpret_cast<const float4*>(row + k + 4 * u);
#pragma unroll
          for (int u = 0; u < 4; ++u)
          {
            acc = __fadd_rn(acc, t[u].x);
            acc = __fadd_rn(acc, t[u].y);
            acc = __fadd_rn(acc, t[u].z);
            acc = __fadd_rn(acc, t[u].w);
          }
        }
        for (; k < cnt; ++k)
          acc = __fadd_rn(acc, row[k]);
      }
      else
      {
        for (; k + 4 <= cnt; k += 4)
        {
          const float4 t = *reinterpret_cast<const float4*>(row + k);
          acc = fmaxf(fmaxf(acc, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));  // w_max starts at 0 and only grows (ParticleFilter.cpp:210-211)
        }
        for (; k < cnt; ++k)
          acc = fmaxf(acc, row[k]);
      }
    }
    __syncthreads();  // products of chunk c are complete, raw[c & 1] and prod[(c - 1) & 1] are free
    if (tid == 0 && c + 2 < nchunks)
      issue(c + 2);
  }
  if (tid < 7)
  {
    const uint32_t col = tid < 6 ? tid : 6u;
    for (uint64_t i = n4; i < n; ++i)
    {
      const float w = particles[i * 7 + 6];
      if (tid < 6)
        acc = __fadd_rn(acc, __fmul_rn(w, particles[i * 7 + col]));
      else
        acc = fmaxf(acc, w);
    }
    out7[tid] = acc;
  }
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
// one warp per component: lane-strided partial sums in a fixed order, then a fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(7 * 32) mean_pass2(const double* __restrict__ part, int nb, float* __restrict__ out7)
{
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = 0;
  for (int k = lane; k < nb; k += 32)
    v = (c == 6) ? fmax(v, part[k * 7 + c]) : v + part[k * 7 + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    v = (c == 6) ? fmax(v, t) : v + t;
  }
  if (lane == 0)
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

// chain_mode (per filter handle) -- 0: exact scan over many CTAs (default), 1: the serial chain kernel, 2: the exact
// scan on one CTA.  All three give the reference's bits; 1 and 2 are kept for A/B timing and cross-checking.
static thread_local uint32_t g_chain_extra = 0;
uint32_t chain_extra_launches_take()
{
  const uint32_t v = g_chain_extra;
  g_chain_extra = 0;
  return v;
}
uint64_t chain_scratch_floats(uint64_t n)
{
  const uint64_t ntiles = (n + kTile - 1) / kTile + 1;
  const uint64_t meta = ntiles * (sizeof(TileSubFn) + sizeof(TileFn) + sizeof(double) + sizeof(float)) + 256;
  return ((n + 63) / 64) * 64 + meta / sizeof(float);
}
static cudaError_t launch_chain(const float* src, uint32_t stride, uint32_t offset, uint64_t n, int mode, float start,
                                float inc, float* prefix, float* total, void* meta, int chain_mode, cudaStream_t stream)
{
  const uint64_t nadd = (mode == kChainThreshold) ? (n > 0 ? n - 1 : 0) : n;
  if (chain_mode == 1)
    chain_kernel<<<1, kChainThreads, 0, stream>>>(src, stride, offset, n, mode, start, inc, prefix, total);
  else if (chain_mode == 2 || meta == nullptr || nadd <= 2 * kTile)
    exact_scan_kernel<<<1, kScanThreads, 0, stream>>>(src, stride, offset, n, mode, start, inc, prefix, total);
  else
  {
    const uint64_t ntiles = (nadd + kTile - 1) / kTile;
    char* m = reinterpret_cast<char*>(meta);
    TileSubFn* subs = reinterpret_cast<TileSubFn*>(m);
    m += ntiles * sizeof(TileSubFn);
    TileFn* fns = reinterpret_cast<TileFn*>(m);
    m += ntiles * sizeof(TileFn);
    double* sums = reinterpret_cast<double*>(m);
    m += ntiles * sizeof(double);
    float* state_in = reinterpret_cast<float*>(m);
    const ChainOps ops{ src, stride, offset, mode, inc, prefix };
    if (mode != kChainThreshold)
      tile_sum_kernel<<<(unsigned)ntiles, kScanThreads, 0, stream>>>(ops, nadd, sums);
    tile_fn_kernel<<<(unsigned)ntiles, kScanThreads, 0, stream>>>(ops, nadd, start, sums, fns, subs);
    resolve_kernel<<<1, kScanThreads, 0, stream>>>(ops, nadd, start, fns, subs, ntiles, state_in, total);
    if (prefix != nullptr)
      emit_kernel<<<(unsigned)ntiles, kScanThreads, 0, stream>>>(ops, nadd, fns, subs, state_in);
    g_chain_extra += (mode != kChainThreshold ? 2 : 1) + (prefix != nullptr ? 1 : 0);
  }
  return cudaGetLastError();
}
static void* chain_meta(float* dense_scratch, uint64_t n)
{
  return dense_scratch ? reinterpret_cast<void*>(dense_scratch + ((n + 63) / 64) * 64) : nullptr;
}

cudaError_t launch_chain_prefix(const float* particles, uint64_t n, int positive_only, float* prefix, float* total,
                                float* dense_scratch, int chain_mode, cudaStream_t stream)
{
  const int mode = positive_only ? kChainSumPositive : kChainSum;
  if (dense_scratch != nullptr && n > 0 && chain_mode != 1)
  {
    extract_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(particles, n, dense_scratch);
    g_chain_extra += 1;
    return launch_chain(dense_scratch, 1, 0, n, mode, 0.f, 0.f, prefix, total, chain_meta(dense_scratch, n), chain_mode,
                        stream);
  }
  return launch_chain(particles, 7, 6, n, mode, 0.f, 0.f, prefix, total, nullptr, chain_mode, stream);
}

cudaError_t launch_chain_sum_array(const float* v, uint64_t n, float* total, float* scratch, int chain_mode,
                                   cudaStream_t stream)
{
  return launch_chain(v, 1, 0, n, kChainSum, 0.f, 0.f, nullptr, total, chain_meta(scratch, n), chain_mode, stream);
}

cudaError_t launch_chain_dense(const float* v, uint64_t n, int positive_only, float* prefix, float* total, float* scratch,
                               int chain_mode, cudaStream_t stream)
{
  return launch_chain(v, 1, 0, n, positive_only ? kChainSumPositive : kChainSum, 0.f, 0.f, prefix, total,
                      chain_meta(scratch, n), chain_mode, stream);
}

cudaError_t launch_chain_thresholds(int S, float* thr, float* scratch, int chain_mode, cudaStream_t stream)
{
  if (S <= 0)
    return cudaSuccess;
  const float inteval = 1.0f / S;  // ParticleFilter.cpp:267 (IEEE f32 divide on the host)
  const float samp0 = 0.5f / S;    // ParticleFilter.cpp:268
  return launch_chain(nullptr, 1, 0, (uint64_t)S, kChainThreshold, samp0, inteval, thr, nullptr,
                      chain_meta(scratch, (uint64_t)S), chain_mode, stream);
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
    mean_exact_kernel<<<1, kChainThreads, 0, stream>>>(particles, n, nullptr, out7);
    return cudaGetLastError();
  }
  double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 16384);
  int nb = (int)blocks_for(n ? n : 1, kRedThreads);
  if (nb > kRedBlocks)
    nb = kRedBlocks;
  mean_pass1<<<nb, kRedThreads, 0, stream>>>(particles, n, part);
  mean_pass2<<<1, 7 * 32, 0, stream>>>(part, nb, out7);
  return cudaGetLastError();
}

cudaError_t launch_mean_exact_carry(const float* particles, uint64_t n, const float* carry7, float* out7, cudaStream_t stream)
{
  mean_exact_kernel<<<1, kChainThreads, 0, stream>>>(particles, n, carry7, out7);
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
