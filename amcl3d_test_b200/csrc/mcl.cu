// mcl.cu -- the direct callers of the path, device side.
//
// Stands behind MonteCarloLocalization::computeSampleNum     amcl3d/src/amcl3d.cpp:242-288 (+ computeBoundary :222-241)
//               MonteCarloLocalization::updateSampleHeight   amcl3d/src/amcl3d.cpp:310-321
// so that one PFResample step (addParticleWeight -> computeSampleNum -> uniformSample -> updateSampleHeight,
// amcl3d.cpp:192-203) runs without moving the particle set to the host.
#include "common.cuh"

namespace a3d
{
namespace
{
__device__ __forceinline__ uint32_t f2key(float f)
{
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// min / max of x, y, z over the particle set as ordered-uint keys: keys[0..2] = min, keys[3..5] = max
__global__ void particle_bounds_kernel(const float* __restrict__ particles, uint64_t n, uint32_t* __restrict__ keys)
{
  uint32_t lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0, 0, 0 };
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const float* p = particles + 7 * i;
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      const uint32_t k = f2key(p[a]);
      lo[a] = min(lo[a], k);
      hi[a] = max(hi[a], k);
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

// amcl3d.cpp:261-271: one bit per (x, y, z, yaw) bin, count of newly set bits
__global__ void sample_bins_kernel(const float* __restrict__ particles, uint64_t n, float xmin, float ymin, float zmin,
                                   int xwidth, int ywidth, int zwidth, int theta, uint32_t* __restrict__ bits,
                                   int* __restrict__ count)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float* p = particles + 7 * i;
  const float xyz_step = 0.2f;
  const int x = (int)__fdiv_rn(__fsub_rn(p[0], xmin), xyz_step);  // int x = (_p.x - xmin)/xyz_step
  const int y = (int)__fdiv_rn(__fsub_rn(p[1], ymin), xyz_step);
  const int z = (int)__fdiv_rn(__fsub_rn(p[2], zmin), xyz_step);
  // int t = (_p.yaw - (-M_PI))/theta_step : f32 - f64 -> f64, / (double)8.0f
  const int t = (int)__ddiv_rn(__dsub_rn((double)p[5], -3.14159265358979323846), 8.0);
  if (x >= 0 && x < xwidth && y >= 0 && y < ywidth && z >= 0 && z < zwidth && t >= 0 && y < theta /* sic, :268 */)
  {
    if (t >= theta)
      return;  // the reference writes out of bounds here (UB); not counted
    const uint32_t cell = (uint32_t)(((x * ywidth + y) * zwidth + z) * theta + t);
    const uint32_t mask = 1u << (cell & 31u);
    const uint32_t old = atomicOr(&bits[cell >> 5], mask);
    if (!(old & mask))
      atomicAdd(count, 1);
  }
}

// amcl3d.cpp:316-319
__global__ void shift_z_kernel(float* __restrict__ particles, uint64_t n, float delta_h)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    particles[7 * i + 2] = __fadd_rn(particles[7 * i + 2], delta_h);
}
}  // namespace

cudaError_t launch_particle_bounds(const float* particles, uint64_t n, uint32_t* keys6, cudaStream_t stream)
{
  const uint32_t init[6] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u };
  cudaError_t e = cudaMemcpyAsync(keys6, init, sizeof(init), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess || n == 0)
    return e;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8)
    blocks = 148 * 8;
  particle_bounds_kernel<<<(unsigned)blocks, 256, 0, stream>>>(particles, n, keys6);
  return cudaGetLastError();
}

cudaError_t launch_sample_bins(const float* particles, uint64_t n, float xmin, float ymin, float zmin, int xwidth, int ywidth,
                               int zwidth, int theta, uint32_t* bits, size_t bits_words, int* count, cudaStream_t stream)
{
  cudaError_t e = cudaMemsetAsync(bits, 0, bits_words * sizeof(uint32_t), stream);
  if (e == cudaSuccess)
    e = cudaMemsetAsync(count, 0, sizeof(int), stream);
  if (e != cudaSuccess || n == 0)
    return e;
  sample_bins_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(particles, n, xmin, ymin, zmin, xwidth, ywidth, zwidth,
                                                                      theta, bits, count);
  return cudaGetLastError();
}

cudaError_t launch_shift_z(float* particles, uint64_t n, float delta_h, cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  shift_z_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(particles, n, delta_h);
  return cudaGetLastError();
}

}  // namespace a3d
