// shard.cu -- particles sharded over the GPUs of one NVLink / NVSwitch box: the exchange step of the global
// particleNormalize + resample done by our own kernels over peer memory instead of an NCCL all-gather of
// the 28-byte records.
//
// Every rank owns one exchange region (one cudaMalloc, mapped into the peers through CUDA IPC handles or, inside one
// process, through peer access):
//     P[2][n_max]        the rank's particle block, ping-pong (resample reads the peers' P[cur] and writes its own P[cur^1])
//     w[2][C]            (C = the capacity the region was sized for, N <= C) dense weights of ALL particles, ping-pong by epoch: every rank STORES its own slice into every
//                        peer's copy over NVLink (4 B per particle instead of 28)
//     wr[2][N]           raw range-beacon weights (row R), same
//     inmap[2][N]        isIntoMap of every particle (particleNormalize zeroes the weight outside the map, ParticleFilter.cpp:187-191)
//     mean[2][world][8]  per-rank partial sums of getMean
//     flags[2][world]    epoch counters: flags[0] "weights of epoch e are in your copy", flags[1] "partial mean of epoch e"
// Sequence of one frame on every rank (all on the rank's stream, no host wait):
//     update (local, no exchange) -> publish_kernel (P2P stores + release flags) -> wait_kernel (acquire) ->
//     the reference's sequential f32 chains over the full dense weight array, replicated on every rank (bit-identical to
//     1 GPU and to the CPU: ParticleFilter.cpp:171-176, 235-247) -> shard_search_kernel: this rank's output slice
//     [lo, hi), the selected source particles are LOADED from the owning peer's block over NVLink -> partial mean ->
//     publish / wait / combine.
// The ping-pong of P, w and flags by epoch parity makes the scheme safe without any further barrier: a rank can only be
// one exchange ahead of its slowest peer (it cannot pass wait(e+1) before that peer published e+1, which the peer does
// only after it finished reading everything of epoch e).
#include "common.cuh"

#include <algorithm>

namespace a3d
{
namespace
{
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer(const float* p)
{
  // peer memory: never through L1 (a line cached there could be a previous epoch's)
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ bool into_map(const MapDev& map, float x, float y, float z)
{
  return map.has_map && ((double)x >= map.min[0]) && ((double)x < map.max[0]) && ((double)y >= map.min[1]) &&
         ((double)y < map.max[1]) && ((double)z >= map.min[2]) && ((double)z < map.max[2]);
}

// This rank's weights (and range weights, and isIntoMap flags) into every rank's dense copy at [lo, lo + n); the last
// block to finish raises flags[rank] = epoch on every peer.
__global__ void __launch_bounds__(256)
    publish_kernel(const __grid_constant__ MapDev map, const __grid_constant__ ShardPeers peers, const float* __restrict__ particles,
                   const float* __restrict__ wr_local, uint32_t n, uint32_t lo, uint32_t dense_stride, uint32_t parity, uint32_t epoch,
                   uint32_t rank, unsigned int* __restrict__ done_counter)
{
  const uint32_t world = peers.world;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const float* p = particles + (size_t)7 * i;
    const float w = p[6];
    const uint8_t in = into_map(map, p[0], p[1], p[2]) ? 1 : 0;
    const float wr = wr_local ? wr_local[i] : 0.f;
    for (uint32_t r = 0; r < world; ++r)
    {
      char* base = peers.base[r];
      reinterpret_cast<float*>(base + peers.off_w)[(size_t)parity * dense_stride + lo + i] = w;
      reinterpret_cast<uint8_t*>(base + peers.off_inmap)[(size_t)parity * dense_stride + lo + i] = in;
      if (wr_local)
        reinterpret_cast<float*>(base + peers.off_wr)[(size_t)parity * dense_stride + lo + i] = wr;
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0)
    last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last)
  {
    __threadfence_system();
    if (threadIdx.x < world)
      st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + peers.off_flags) + (size_t)0 * kMaxShardWorld + rank, epoch);
    if (threadIdx.x == 0)
      *done_counter = 0;  // ready for the next launch (stream order)
  }
}

// spins until every rank's flag of `kind` reached `epoch`.  A peer that never arrives (crashed process) must not hang
// the GPU: after ~4 s of spinning the error word is set and the kernel returns.
__global__ void wait_kernel(const uint32_t* __restrict__ flags, uint32_t kind, uint32_t world, uint32_t epoch,
                            uint32_t* __restrict__ error)
{
  const uint32_t r = threadIdx.x;
  if (r >= world)
    return;
  const uint32_t* f = flags + (size_t)kind * kMaxShardWorld + r;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(f) - epoch) < 0)
  {
    if (clock64() - t0 > 8000000000ll)
    {
      atomicExch(error, 1u + r);
      return;
    }
    __nanosleep(200);
  }
}

// particleNormalize on the dense copy (ParticleFilter.cpp:180-195)
__global__ void dense_normalize_kernel(float* __restrict__ w, const uint8_t* __restrict__ inmap, uint64_t n,
                                       const float* __restrict__ total)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float wtp = total[0];
  float v = (wtp > 0.f) ? __fdiv_rn(w[i], wtp) : 0.f;
  if (!inmap[i])
    v = 0.f;
  w[i] = v;
}
// row R fusion on the dense copies: w = alpha * wp / sum(wp) + (1 - alpha) * wr / sum(wr)
__global__ void dense_fuse_kernel(float* __restrict__ w, const float* __restrict__ wr, uint64_t n, const float* __restrict__ sum_p,
                                  const float* __restrict__ sum_r, float alpha)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float wtp = sum_p[0], wtr = sum_r[0];
  const float wp = wtp > 0.f ? __fdiv_rn(w[i], wtp) : 0.f;
  const float wq = wtr > 0.f ? __fdiv_rn(wr[i], wtr) : 0.f;
  w[i] = __fadd_rn(__fmul_rn(wp, alpha), __fmul_rn(wq, __fsub_rn(1.f, alpha)));
}

// ParticleFilter.cpp:239-250 for this rank's outputs m in [m0, m1): i(m) = first i with !(u_m > c_i), clamped to N-1;
// the source particle is read from the block of the rank that owns it (peer memory), w = 1/N
__global__ void __launch_bounds__(256)
    shard_search_kernel(const __grid_constant__ ShardPeers peers, const float* __restrict__ prefix, uint64_t n, float u01,
                        uint64_t m0, uint64_t m1, uint32_t src_buf, float* __restrict__ out, int32_t* __restrict__ idx)
{
  const uint64_t m = m0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= m1)
    return;
  const float factor = __fdiv_rn(1.f, __ull2float_rn(n));
  const float r = __fmul_rn(factor, u01);
  const float u = __fadd_rn(r, __fmul_rn(factor, __uint2float_rn((uint32_t)m)));
  uint64_t lo = 0, hi = n;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (u > prefix[mid])
      lo = mid + 1;
    else
      hi = mid;
  }
  const uint64_t src = (lo >= n) ? n - 1 : lo;
  // owner of global index src under the contiguous block partition (first n % world ranks hold one more)
  const uint64_t base_n = n / peers.world, rem = n % peers.world;
  const uint64_t split = rem * (base_n + 1);
  uint32_t owner;
  uint64_t local;
  if (src < split)
  {
    owner = (uint32_t)(src / (base_n + 1));
    local = src - (uint64_t)owner * (base_n + 1);
  }
  else
  {
    owner = (uint32_t)(rem + (src - split) / base_n);
    local = (src - split) - (uint64_t)(owner - rem) * base_n;
  }
  const float* p = reinterpret_cast<const float*>(peers.base[owner] + peers.off_p + (size_t)src_buf * peers.p_bytes) + 7 * local;
  float* q = out + 7 * (m - m0);
#pragma unroll
  for (int k = 0; k < 6; ++k)
    q[k] = ld_peer(p + k);
  q[6] = factor;
  if (idx != nullptr)
    idx[m - m0] = (int32_t)src;
}

// ---- fast (non-exact) mode --------------------------------------------------------------------------------------
// north_star's "allreduce of the weight sum + allgather of prefix offsets": no rank ever sees another rank's weights.
// Two scalar exchanges (local weight sum; local total of the normalised weights -> base offset of the rank's stretch of
// the global prefix), then the SOURCE side places its particles: the thresholds u_m that fall into (base_r, base_r+1] are
// searched in the rank's own prefix and the selected particle is stored straight into the output block of the rank that
// owns slot m.  The sums are sequential f32 chains inside a rank but combined across ranks in rank order, so a threshold
// sitting within an ulp of a prefix value can pick the neighbouring particle: indices are NOT guaranteed bit-identical
// to one GPU (tests report the mismatch count); the exact mode above is the default.
__global__ void publish_scalar_kernel(const __grid_constant__ ShardPeers peers, const float* __restrict__ value, uint32_t step,
                                      uint32_t parity, uint32_t epoch, uint32_t rank)
{
  const uint32_t r = threadIdx.x;
  if (r >= peers.world)
    return;
  float* dst = reinterpret_cast<float*>(peers.base[r] + peers.off_sums) + ((size_t)step * 2 + parity) * kMaxShardWorld + rank;
  *dst = value[0];
  __threadfence_system();
  st_release_sys(reinterpret_cast<uint32_t*>(peers.base[r] + peers.off_flags) + (size_t)(2 + step) * kMaxShardWorld + rank, epoch);
}
__global__ void combine_scalar_kernel(const float* __restrict__ sums_all, uint32_t world, uint32_t rank, float* __restrict__ out2)
{
  if (threadIdx.x != 0)
    return;
  float total = 0.f, before = 0.f;
  for (uint32_t r = 0; r < world; ++r)
  {
    if (r == rank)
      before = total;
    total = __fadd_rn(total, sums_all[r]);
  }
  out2[0] = total;
  out2[1] = before;
}

__device__ __forceinline__ float threshold(float r, float factor, uint32_t m)
{
  return __fadd_rn(r, __fmul_rn(factor, __uint2float_rn(m)));  // ParticleFilter.cpp:241
}
// first m in [0, n] with u_m > c (n if none): u_m is non-decreasing in m
__device__ __forceinline__ uint32_t first_above(float r, float factor, uint32_t n, float c)
{
  uint32_t lo = 0, hi = n;
  while (lo < hi)
  {
    const uint32_t mid = (lo + hi) >> 1;
    if (threshold(r, factor, mid) > c)
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}
__global__ void __launch_bounds__(256)
    shard_push_kernel(const __grid_constant__ ShardPeers peers, const float* __restrict__ particles, const float* __restrict__ prefix,
                      uint32_t n_local, uint64_t lo, uint64_t n_total, const float* __restrict__ base2, float u01, uint32_t rank,
                      uint32_t dst_buf, uint32_t epoch, unsigned int* __restrict__ done_counter)
{
  const uint32_t N = (uint32_t)n_total;
  const float factor = __fdiv_rn(1.f, __ull2float_rn(n_total));
  const float r = __fmul_rn(factor, u01);
  const float base = base2[1];
  __shared__ uint32_t range[2];
  if (threadIdx.x == 0)
  {
    // this rank's stretch of the global prefix is (base, base + local total]; rank 0 also takes u_m <= 0, the last rank
    // everything above the global total (clamped to the last particle of the set like the exact path)
    const float end = n_local ? __fadd_rn(base, prefix[n_local - 1]) : base;
    const bool first = rank == 0, last = rank + 1 == peers.world;
    range[0] = first ? 0u : first_above(r, factor, N, base);
    range[1] = (n_local == 0) ? range[0] : (last ? N : first_above(r, factor, N, end));
    if (n_local == 0)
      range[1] = range[0];
  }
  __syncthreads();
  const uint32_t m_a = range[0], m_b = range[1];
  const uint64_t base_n = n_total / peers.world, rem = n_total % peers.world;
  const uint64_t split = rem * (base_n + 1);
  for (uint64_t m = (uint64_t)m_a + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; m < m_b; m += (uint64_t)gridDim.x * blockDim.x)
  {
    const float u = threshold(r, factor, (uint32_t)m);
    uint32_t a = 0, b = n_local;
    while (a < b)
    {
      const uint32_t mid = (a + b) >> 1;
      if (u > __fadd_rn(base, prefix[mid]))
        a = mid + 1;
      else
        b = mid;
    }
    const uint32_t src = (a >= n_local) ? n_local - 1 : a;
    uint32_t owner;
    uint64_t slot;
    if (m < split)
    {
      owner = (uint32_t)(m / (base_n + 1));
      slot = m - (uint64_t)owner * (base_n + 1);
    }
    else
    {
      owner = (uint32_t)(rem + (m - split) / base_n);
      slot = (m - split) - (uint64_t)(owner - rem) * base_n;
    }
    const float* p = particles + (size_t)7 * src;
    float* q = reinterpret_cast<float*>(peers.base[owner] + peers.off_p + (size_t)dst_buf * peers.p_bytes) + 7 * slot;
#pragma unroll
    for (int k = 0; k < 6; ++k)
      q[k] = p[k];
    q[6] = factor;
    reinterpret_cast<int32_t*>(peers.base[owner] + peers.off_idx)[slot] = (int32_t)(lo + src);
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last_block;
  if (threadIdx.x == 0)
    last_block = atomicAdd(done_counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last_block)
  {
    __threadfence_system();
    if (threadIdx.x < peers.world)
      st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + peers.off_flags) + (size_t)4 * kMaxShardWorld + rank, epoch);
    if (threadIdx.x == 0)
      *done_counter = 0;
  }
}

// this rank's 7 partial floats (sums of w*x.., max w) into every rank's mean[parity][rank], then flags[1][rank] = epoch
__global__ void publish_mean_kernel(const __grid_constant__ ShardPeers peers, const float* __restrict__ part7, uint32_t parity,
                                    uint32_t epoch, uint32_t rank)
{
  const uint32_t r = threadIdx.x;
  if (r >= peers.world)
    return;
  float* dst = reinterpret_cast<float*>(peers.base[r] + peers.off_mean) + ((size_t)parity * kMaxShardWorld + rank) * 8;
#pragma unroll
  for (int k = 0; k < 7; ++k)
    dst[k] = part7[k];
  __threadfence_system();
  st_release_sys(reinterpret_cast<uint32_t*>(peers.base[r] + peers.off_flags) + (size_t)1 * kMaxShardWorld + rank, epoch);
}
// getMean over all ranks: sums of the per-rank partial sums in rank order (f64), w = max (ParticleFilter.cpp:198-216)
__global__ void combine_mean_kernel(const float* __restrict__ mean_all, uint32_t world, float* __restrict__ out7)
{
  const uint32_t k = threadIdx.x;
  if (k >= 7)
    return;
  if (k < 6)
  {
    double s = 0;
    for (uint32_t r = 0; r < world; ++r)
      s += (double)mean_all[(size_t)r * 8 + k];
    out7[k] = (float)s;
  }
  else
  {
    float mx = 0.f;
    for (uint32_t r = 0; r < world; ++r)
      mx = fmaxf(mx, mean_all[(size_t)r * 8 + 6]);
    out7[6] = mx;
  }
}
}  // namespace

cudaError_t launch_shard_publish(const MapDev& map, const ShardPeers& peers, const float* particles, const float* wr_local,
                                 uint32_t n, uint32_t lo, uint32_t dense_stride, uint32_t parity, uint32_t epoch, uint32_t rank,
                                 unsigned int* done_counter, cudaStream_t stream)
{
  uint32_t blocks = (n + 255) / 256;
  if (blocks > 148 * 4)
    blocks = 148 * 4;
  if (blocks == 0)
    blocks = 1;
  publish_kernel<<<blocks, 256, 0, stream>>>(map, peers, particles, wr_local, n, lo, dense_stride, parity, epoch, rank, done_counter);
  return cudaGetLastError();
}
cudaError_t launch_shard_wait(const uint32_t* flags, uint32_t kind, uint32_t world, uint32_t epoch, uint32_t* error,
                              cudaStream_t stream)
{
  wait_kernel<<<1, 32, 0, stream>>>(flags, kind, world, epoch, error);
  return cudaGetLastError();
}
cudaError_t launch_dense_normalize(float* w, const uint8_t* inmap, uint64_t n, const float* total, cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  dense_normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, inmap, n, total);
  return cudaGetLastError();
}
cudaError_t launch_dense_fuse(float* w, const float* wr, uint64_t n, const float* sum_p, const float* sum_r, float alpha,
                              cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  dense_fuse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, wr, n, sum_p, sum_r, alpha);
  return cudaGetLastError();
}
cudaError_t launch_shard_search(const ShardPeers& peers, const float* prefix, uint64_t n, float u01, uint64_t m0, uint64_t m1,
                                uint32_t src_buf, float* out, int32_t* idx, cudaStream_t stream)
{
  if (m1 <= m0)
    return cudaSuccess;
  shard_search_kernel<<<(unsigned)((m1 - m0 + 255) / 256), 256, 0, stream>>>(peers, prefix, n, u01, m0, m1, src_buf, out, idx);
  return cudaGetLastError();
}
cudaError_t launch_shard_publish_scalar(const ShardPeers& peers, const float* value, uint32_t step, uint32_t parity, uint32_t epoch,
                                        uint32_t rank, cudaStream_t stream)
{
  publish_scalar_kernel<<<1, 32, 0, stream>>>(peers, value, step, parity, epoch, rank);
  return cudaGetLastError();
}
cudaError_t launch_shard_combine_scalar(const float* sums_all, uint32_t world, uint32_t rank, float* out2, cudaStream_t stream)
{
  combine_scalar_kernel<<<1, 32, 0, stream>>>(sums_all, world, rank, out2);
  return cudaGetLastError();
}
cudaError_t launch_shard_push(const ShardPeers& peers, const float* particles, const float* prefix_local, uint32_t n_local,
                              uint64_t lo, uint64_t n_total, const float* base2, float u01, uint32_t rank, uint32_t dst_buf,
                              uint32_t epoch, unsigned int* done_counter, cudaStream_t stream)
{
  // sized for a rank that places about twice its share; a rank holding most of the weight loops (grid-stride)
  uint64_t want = 2 * (n_total / (peers.world ? peers.world : 1)) + 256;
  uint32_t blocks = (uint32_t)std::min<uint64_t>((want + 255) / 256, 148 * 16);
  if (blocks == 0)
    blocks = 1;
  shard_push_kernel<<<blocks, 256, 0, stream>>>(peers, particles, prefix_local, n_local, lo, n_total, base2, u01, rank, dst_buf, epoch,
                                               done_counter);
  return cudaGetLastError();
}
cudaError_t launch_shard_publish_mean(const ShardPeers& peers, const float* part7, uint32_t parity, uint32_t epoch, uint32_t rank,
                                      cudaStream_t stream)
{
  publish_mean_kernel<<<1, 32, 0, stream>>>(peers, part7, parity, epoch, rank);
  return cudaGetLastError();
}
cudaError_t launch_shard_combine_mean(const float* mean_all, uint32_t world, float* out7, cudaStream_t stream)
{
  combine_mean_kernel<<<1, 32, 0, stream>>>(mean_all, world, out7);
  return cudaGetLastError();
}

}  // namespace a3d
