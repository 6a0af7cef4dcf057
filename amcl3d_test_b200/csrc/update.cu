// update.cu -- K3: fused ParticleFilter::update  (sm_100a)
//
// Stands behind ParticleFilter::update(cloud)            amcl3d/src/ParticleFilter.cpp:114-139
//               Grid3d::isIntoMap                        amcl3d/src/Grid3d.cpp:365-372
//               Grid3d::computeCloudWeight (7-arg)       amcl3d/src/Grid3d.cpp:234-301
// plus the range-only beacon weight of SURVEY 8(a) row R (not present in the reference tree).
//
// Mapping: one thread = one particle, so the f32 weight is accumulated in cloud order exactly like
// the reference's serial loop (Grid3d.cpp:292) -- a shuffle/tree reduction would be *more* accurate
// than the reference and miss the 1e-5 criterion at >= 10 k points.  A CTA streams the packed cloud
// through a double-buffered shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier);
// every lane reads the same point (smem broadcast), applies its particle's 3x3 rotation in registers
// with the reference's rounding sequence (f32 mul/add, no FMA, f64 offset add), and issues kUnroll
// independent voxel gathers before the in-order adds.  The kernel is a gather: one 32-byte sector per
// in-bounds evaluation, no reuse across lanes, no tensor-core work.
//
// Exact arithmetic notes:
//  * transform: ((x*r00 + y*r01) + z*r02) in f32, then + f64 offset, rounded to f32.  When the f64
//    offset is exactly representable in f32 the f64 add is replaced by one f32 add (innocuous double
//    rounding: 53 >= 2*24+2); otherwise the f64 path is taken.  Both are bit-identical to the reference.
//  * bounds: f32 value against f64 size is done against size_up = smallest f32 >= size (equivalent).
//  * voxel index floor(fl64(v / resol)): f32 quotient v*rinv with an error band; inside the band an
//    f64 product with a 1e-9 band; inside that the real f64 division.  Result identical to the DDIV.
//
// Two code paths (bit-identical weights):
//  * update_kernel<rep>  -- one launch, cloud swept in the caller's order, weight accumulated in registers.
//  * pose_kernel + sweep_kernel + ordered_sum_kernel -- the large-set path on u8-coded grids: the rotation and
//    offsets are computed once per particle (pose_kernel), the Morton-sorted cloud is swept chunk-major so that
//    all particles gather from the same neighbourhood of the map at the same time, 1-byte codes are parked at
//    the point's original row (sweep_kernel), and the ordered sum adds them in the caller's cloud order.
#include "common.cuh"

#include <math_constants.h>

#include <cub/device/device_radix_sort.cuh>

namespace a3d
{
namespace
{
#ifndef AMCL3D_THREADS
#define AMCL3D_THREADS 256
#endif
#ifndef AMCL3D_STAGE_TILE
#define AMCL3D_STAGE_TILE 128
#endif
#ifndef AMCL3D_TILE
#define AMCL3D_TILE 512
#endif
#ifndef AMCL3D_UNROLL
#define AMCL3D_UNROLL 4
#endif
constexpr int kThreads = AMCL3D_THREADS;  // particles per CTA
constexpr int kTile = AMCL3D_TILE;        // points per shared-memory stage (16 B each), single-kernel path
constexpr int kStageTile = AMCL3D_STAGE_TILE;  // points per block of the staging sweep = its lock-step granularity
constexpr int kStages = 2;
constexpr int kUnroll = AMCL3D_UNROLL;    // independent gathers per batch (two batches in flight)
static_assert(kTile % (2 * kUnroll) == 0 && kStageTile % (2 * kUnroll) == 0 && 32 % (2 * kUnroll) == 0,
              "tiles and the 32-point padding must hold whole batch pairs");

// Grid3d.cpp:240-250
__device__ __forceinline__ void rotation(float roll, float pitch, float yaw, int trig_mode, float r[9])
{
  if (trig_mode == AMCL3D_B200_TRIG_F64)
  {
    double sr, cr, sp, cp, sy, cy;
    sincos((double)roll, &sr, &cr);
    sincos((double)pitch, &sp, &cp);
    sincos((double)yaw, &sy, &cy);
    r[0] = __double2float_rn(__dmul_rn(cy, cp));
    r[1] = __double2float_rn(__dsub_rn(__dmul_rn(__dmul_rn(cy, sp), sr), __dmul_rn(sy, cr)));
    r[2] = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(cy, sp), cr), __dmul_rn(sy, sr)));
    r[3] = __double2float_rn(__dmul_rn(sy, cp));
    r[4] = __double2float_rn(__dadd_rn(__dmul_rn(__dmul_rn(sy, sp), sr), __dmul_rn(cy, cr)));
    r[5] = __double2float_rn(__dsub_rn(__dmul_rn(__dmul_rn(sy, sp), cr), __dmul_rn(cy, sr)));
    r[6] = __double2float_rn(-sp);
    r[7] = __double2float_rn(__dmul_rn(cp, sr));
    r[8] = __double2float_rn(__dmul_rn(cp, cr));
  }
  else
  {
    const float sr = sinf(roll), cr = cosf(roll);
    const float sp = sinf(pitch), cp = cosf(pitch);
    const float sy = sinf(yaw), cy = cosf(yaw);
    r[0] = __fmul_rn(cy, cp);
    r[1] = __fsub_rn(__fmul_rn(__fmul_rn(cy, sp), sr), __fmul_rn(sy, cr));
    r[2] = __fadd_rn(__fmul_rn(__fmul_rn(cy, sp), cr), __fmul_rn(sy, sr));
    r[3] = __fmul_rn(sy, cp);
    r[4] = __fadd_rn(__fmul_rn(__fmul_rn(sy, sp), sr), __fmul_rn(cy, cr));
    r[5] = __fsub_rn(__fmul_rn(__fmul_rn(sy, sp), cr), __fmul_rn(cy, sr));
    r[6] = -sp;
    r[7] = __fmul_rn(cp, sr);
    r[8] = __fmul_rn(cp, cr);
  }
}

// Voxel index floor(fl64(v / resol)) for 0 <= v < size_up (see header comment).  Fast tier: k = floor(v*rinv_f) in
// f32, valid whenever the f32 quotient is farther than eps from an integer (eps bounds its error).  The caller
// tests the three axes with ONE predicate; voxel_exact() is the rare out-of-line tier: f64 product with a 1e-9
// band, then the real f64 division.
__device__ __noinline__ uint32_t voxel_exact(float v, double rinv, double resol)
{
  const double vd = (double)v;
  const double qd = __dmul_rn(vd, rinv);
  double kd = floor(qd);
  const double frd = qd - kd;
  if (frd < 1e-9 || frd > 1.0 - 1e-9)
    kd = floor(__ddiv_rn(vd, resol));
  return (uint32_t)kd;
}

template <bool kF32Offset>
__device__ __forceinline__ float add_offset(float s, float off_f, double off_d)
{
  if (kF32Offset)
    return __fadd_rn(s, off_f);
  return __double2float_rn(__dadd_rn((double)s, off_d));
}

// element index of voxel (ix,iy,iz) in the representation's layout
template <int kRep>
__device__ __forceinline__ uint64_t voxel_address(const MapDev& map, uint32_t ix, uint32_t iy, uint32_t iz)
{
  if (kRep != kRepFloat)
    return coded_index<kRep>(ix, iy, iz, map.nbx, map.nby);
  return (uint64_t)ix + (uint64_t)iy * map.size[0] + (uint64_t)iz * map.step_z;  // Grid3d.cpp:288
}

// The gathers use one useful element of every line they touch.  AMCL3D_LOAD_MODE selects the cache
// operator: 0 ld.global.nc (L1-allocating), 1 ld.global.cg (L2 only), 2 ld.global.nc.L1::no_allocate,
// 3 ld.global.L1::no_allocate.L2::evict_first, 4 ld.global.nc.L2::64B (the L2 fetches 64 instead of 128 bytes from DRAM
// on a miss: profiles/r2_l2_prefetch_size_probe.txt), 5 ld.global.nc.L2::256B.
#ifndef AMCL3D_LOAD_MODE
#define AMCL3D_LOAD_MODE 0
#endif
#if AMCL3D_LOAD_MODE == 1
#define A3D_LD(suffix) "ld.global.cg." suffix
#elif AMCL3D_LOAD_MODE == 2
#define A3D_LD(suffix) "ld.global.nc.L1::no_allocate." suffix
#elif AMCL3D_LOAD_MODE == 3
#define A3D_LD(suffix) "ld.global.L1::no_allocate." suffix
#elif AMCL3D_LOAD_MODE == 4
#define A3D_LD(suffix) "ld.global.nc.L2::64B." suffix
#elif AMCL3D_LOAD_MODE == 5
#define A3D_LD(suffix) "ld.global.nc.L2::256B." suffix
#else
#define A3D_LD(suffix) "ld.global.nc." suffix
#endif
__device__ __forceinline__ float gather_f32(const float* p)
{
  float v;
  asm volatile(A3D_LD("f32") " %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t gather_u8(const uint8_t* p)
{
  uint32_t v;
  asm volatile(A3D_LD("u8") " %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t gather_u16(const uint16_t* p)
{
  uint32_t v;
  asm volatile(A3D_LD("u16") " %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---- one particle x kUnroll consecutive points ------------------------------------------------------
// The hot loop is software-pipelined by hand: addresses of batch k+1 are computed and their gathers
// issued before the values of batch k are consumed, so two batches of independent loads are in flight
// per thread and the gather latency overlaps the (branchy, exact) index arithmetic.
template <int kRep>
struct Raw
{
  typedef uint32_t type;  // dictionary code
};
template <>
struct Raw<kRepFloat>
{
  typedef float type;  // the cell itself
};

// returns the bit mask of evaluations that pass every bounds test; addr[u] = element index of their voxel.
// Branch-free on the fast path: everything is computed for every point and masked; the only branch is the rare
// "some axis is within eps of a voxel face" fix-up.
template <bool kF32Offset, int kRep>
__device__ __forceinline__ uint32_t batch_addresses(const MapDev& map, const float4* __restrict__ pts, const float r[9],
                                                    const float off_f[3], const double off_d[3], uint64_t addr[kUnroll],
                                                    uint32_t orow[kUnroll])
{
  float4 pt[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u)
  {
    pt[u] = pts[u];  // broadcast LDS.128, issued back to back
    orow[u] = __float_as_uint(pt[u].w);  // staging sweep: index of this point in the caller's cloud order
  }
  const float rinv_f = map.rinv_f;
  const float band = map.band;  // 0.5 - max eps
  uint32_t mask = 0;
#pragma unroll
  for (int u = 0; u < kUnroll; ++u)
  {
    // Grid3d.cpp:275-277
    const float sx = __fadd_rn(__fadd_rn(__fmul_rn(pt[u].x, r[0]), __fmul_rn(pt[u].y, r[1])), __fmul_rn(pt[u].z, r[2]));
    const float sy = __fadd_rn(__fadd_rn(__fmul_rn(pt[u].x, r[3]), __fmul_rn(pt[u].y, r[4])), __fmul_rn(pt[u].z, r[5]));
    const float sz = __fadd_rn(__fadd_rn(__fmul_rn(pt[u].x, r[6]), __fmul_rn(pt[u].y, r[7])), __fmul_rn(pt[u].z, r[8]));
    const float nx = add_offset<kF32Offset>(sx, off_f[0], off_d[0]);
    const float ny = add_offset<kF32Offset>(sy, off_f[1], off_d[1]);
    const float nz = add_offset<kF32Offset>(sz, off_f[2], off_d[2]);
    // Grid3d.cpp:279-280 (NaN fails every comparison, like the reference)
    bool in = (nx >= 0.f) & (nx < map.size_up[0]) & (ny >= 0.f) & (ny < map.size_up[1]) & (nz >= 0.f) &
              (nz < map.size_up[2]);
    // Grid3d.cpp:282-284, fast tier on all three axes at once
    const float qx = __fmul_rn(nx, rinv_f), qy = __fmul_rn(ny, rinv_f), qz = __fmul_rn(nz, rinv_f);
    const float kx = floorf(qx), ky = floorf(qy), kz = floorf(qz);
    const float dev = fmaxf(fmaxf(fabsf(__fsub_rn(__fsub_rn(qx, kx), 0.5f)), fabsf(__fsub_rn(__fsub_rn(qy, ky), 0.5f))),
                            fabsf(__fsub_rn(__fsub_rn(qz, kz), 0.5f)));
    uint32_t ix = (uint32_t)kx, iy = (uint32_t)ky, iz = (uint32_t)kz;
    if (in && !(dev < band))  // rare: within eps of a voxel face on some axis -> exact tiers
    {
      ix = voxel_exact(nx, map.rinv, map.resol);
      iy = voxel_exact(ny, map.rinv, map.resol);
      iz = voxel_exact(nz, map.rinv, map.resol);
    }
    in = in & (ix < map.size[0]) & (iy < map.size[1]) & (iz < map.size[2]);  // Grid3d.cpp:286 (and :290)
    addr[u] = voxel_address<kRep>(map, ix, iy, iz);
    mask |= in ? (1u << u) : 0u;
  }
  return mask;
}

template <int kRep>
__device__ __forceinline__ void batch_issue(const MapDev& map, const uint64_t addr[kUnroll], uint32_t mask,
                                            typename Raw<kRep>::type raw[kUnroll])
{
#pragma unroll
  for (int u = 0; u < kUnroll; ++u)
  {
    raw[u] = 0;
    if ((mask >> u) & 1u)
    {
      if (kRep == kRepFloat)
        raw[u] = gather_f32(map.prob + addr[u]);
      else if (kRep == kRepCode8)
        raw[u] = gather_u8(reinterpret_cast<const uint8_t*>(map.codes) + addr[u]);
      else
        raw[u] = gather_u16(reinterpret_cast<const uint16_t*>(map.codes) + addr[u]);
    }
  }
}

template <int kRep>
__device__ __forceinline__ void batch_consume(const typename Raw<kRep>::type raw[kUnroll], uint32_t mask,
                                              const float* __restrict__ lut_s, float& weight, int& cnt)
{
  float v[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u)
  {
    if (kRep == kRepFloat)
      v[u] = raw[u];
    else
      v[u] = lut_s[(uint32_t)raw[u]];  // the same f32 the float grid holds for this voxel
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u)
    if ((mask >> u) & 1u)
      weight = __fadd_rn(weight, v[u]);  // Grid3d.cpp:292, cloud order
  cnt += __popc(mask);
}

// The single-kernel path accumulates in registers, in cloud order: the reference's f32 rounding chain (Grid3d.cpp:292).
template <int kRep>
struct AccumSink
{
  const float* __restrict__ lut_s;
  float weight;
  int cnt;
  __device__ __forceinline__ void put(const typename Raw<kRep>::type raw[kUnroll], uint32_t mask, const uint32_t* /*orow*/)
  {
    batch_consume<kRep>(raw, mask, lut_s, weight, cnt);
  }
};
constexpr uint32_t kStageSkip = 0xFFu;  // staged code of an evaluation that loaded nothing (large-set path)

// one tile of `count` points (a multiple of 2*kUnroll) for one particle; row0 = index of tile[0] in the sweep
template <bool kF32Offset, int kRep, typename Sink>
__device__ __forceinline__ void tile_loop(const MapDev& map, const float4* __restrict__ tile, uint32_t count, uint32_t /*row0*/,
                                          const float r[9], const float off_f[3], const double off_d[3], Sink& sink)
{
  typedef typename Raw<kRep>::type raw_t;
  uint64_t addr[kUnroll];
  raw_t raw_a[kUnroll], raw_b[kUnroll];
  uint32_t row_a[kUnroll], row_b[kUnroll];
  uint32_t mask_a, mask_b;
  mask_a = batch_addresses<kF32Offset, kRep>(map, tile, r, off_f, off_d, addr, row_a);
  batch_issue<kRep>(map, addr, mask_a, raw_a);
#pragma unroll 1
  for (uint32_t base = kUnroll; base < count; base += 2 * kUnroll)
  {
    mask_b = batch_addresses<kF32Offset, kRep>(map, tile + base, r, off_f, off_d, addr, row_b);
    batch_issue<kRep>(map, addr, mask_b, raw_b);
    sink.put(raw_a, mask_a, row_a);
    if (base + kUnroll < count)
    {
      mask_a = batch_addresses<kF32Offset, kRep>(map, tile + base + kUnroll, r, off_f, off_d, addr, row_a);
      batch_issue<kRep>(map, addr, mask_a, raw_a);
    }
    sink.put(raw_b, mask_b, row_b);
  }
}

// per-particle epilogue shared by the one- and two-phase paths
__device__ __forceinline__ void finish_particle(float* __restrict__ particles, uint32_t i, bool inmap, float weight, int cnt,
                                                uint32_t npts, float px, float py, float pz,
                                                const BeaconSet* __restrict__ beacons, float* __restrict__ wr_out)
{
  // Grid3d.cpp:299: (n < 10) ? 0 : weight / cloud->size()   (f32 / (float)size_t)
  float w = 0.f;
  if (inmap && cnt >= 10)
    w = __fdiv_rn(weight, __uint2float_rn(npts));
  particles[(size_t)7 * i + 6] = w;
  if (wr_out != nullptr)
  {
    // row R: wr = prod_b k1*exp(-k2*(|p - a_b| - r_b)^2); 0 outside the map or without beacons
    float wr = 0.f;
    const uint32_t nb = beacons->n;
    if (inmap && nb > 0)
    {
      wr = 1.f;
      const float k1 = beacons->k1, k2 = beacons->k2;
      for (uint32_t b = 0; b < nb; ++b)
      {
        const float dx = __fsub_rn(px, beacons->ax[b]);
        const float dy = __fsub_rn(py, beacons->ay[b]);
        const float dz = __fsub_rn(pz, beacons->az[b]);
        const float rr = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        const float e = __fsub_rn(rr, beacons->r[b]);
        wr = __fmul_rn(wr, __fmul_rn(k1, expf(__fmul_rn(__fmul_rn(-k2, e), e))));
      }
    }
    wr_out[i] = wr;
  }
}
__device__ __forceinline__ bool particle_gate(const MapDev& map, float px, float py, float pz)
{
  // Grid3d.cpp:365-372 -- f32 promoted against the f64 bounds; update() skips particles outside
  // (has_map == 2: Grid3d::computeCloudWeight called directly, which has no such gate)
  const bool gate = (map.has_map == 2) ||
                    (((double)px >= map.min[0]) && ((double)px < map.max[0]) && ((double)py >= map.min[1]) &&
                     ((double)py < map.max[1]) && ((double)pz >= map.min[2]) && ((double)pz < map.max[2]));
  return map.has_map && map.has_grid && gate;
}

#ifndef AMCL3D_MINBLOCKS
#define AMCL3D_MINBLOCKS 3
#endif
// The whole update in one kernel: one block sweeps the whole cloud, in the caller's order, for its 256 particles.
template <int kRep>
__global__ void __launch_bounds__(kThreads, AMCL3D_MINBLOCKS)
    update_kernel(const __grid_constant__ MapDev map, float* __restrict__ particles, uint32_t n,
                  const float4* __restrict__ cloud, uint32_t npts, uint32_t npad, const BeaconSet* __restrict__ beacons,
                  float* __restrict__ wr_out, int trig_mode, unsigned long long* __restrict__ counter)
{
  constexpr int kT = kTile;
  __shared__ __align__(128) float4 tile[kStages][kT];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  extern __shared__ float lut_s[];  // kRep != kRepFloat: the code -> f32 dictionary

  const uint32_t tid = threadIdx.x;
  const uint32_t i = blockIdx.x * kThreads + tid;
  const bool live = i < n;

  float px = 0, py = 0, pz = 0, roll = 0, pitch = 0, yaw = 0;
  if (live)
  {
    const float* p = particles + (size_t)7 * i;
    px = p[0];
    py = p[1];
    pz = p[2];
    roll = p[3];
    pitch = p[4];
    yaw = p[5];
  }
  const bool inmap = live && particle_gate(map, px, py, pz);

  float r[9];
  rotation(roll, pitch, yaw, trig_mode, r);
  // Grid3d.cpp:256-258
  double off_d[3];
  off_d[0] = __dsub_rn((double)px, map.min[0]);
  off_d[1] = __dsub_rn((double)py, map.min[1]);
  off_d[2] = __dsub_rn((double)pz, map.min[2]);
  float off_f[3];
  bool f32_exact = true;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    off_f[a] = __double2float_rn(off_d[a]);
    f32_exact = f32_exact && ((double)off_f[a] == off_d[a]);
  }
  // warp-uniform choice of the loop flavour; lanes that are not in the map do not vote it down
  const bool warp_f32 = __all_sync(0xffffffffu, f32_exact || !inmap);

  const uint32_t ntiles = (npad + kT - 1) / kT;
  constexpr uint32_t tile0 = 0;
  if (tid == 0)
  {
#pragma unroll
    for (int s = 0; s < kStages; ++s)
      mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (kRep != kRepFloat)
    for (uint32_t k = tid; k < map.lut_n; k += kThreads)
      lut_s[k] = map.lut[k];
  __syncthreads();
  if (tid == 0)
  {
    for (uint32_t t = 0; t < ntiles && t < (uint32_t)kStages; ++t)
    {
      const uint32_t count = min((uint32_t)kT, npad - (tile0 + t) * kT);
      mbar_expect_tx(&full_bar[t], count * 16u);
      tma_bulk_g2s(&tile[t][0], cloud + (size_t)(tile0 + t) * kT, count * 16u, &full_bar[t]);
    }
  }

  AccumSink<kRep> acc{ lut_s, 0.f, 0 };
  for (uint32_t t = 0; t < ntiles; ++t)
  {
    const uint32_t s = t % kStages;
    const uint32_t parity = (t / kStages) & 1u;
    mbar_wait(&full_bar[s], parity);
    const uint32_t row0 = (tile0 + t) * kT;
    const uint32_t count = min((uint32_t)kT, npad - row0);
    if (inmap)
    {
      if (warp_f32)
        tile_loop<true, kRep>(map, &tile[s][0], count, row0, r, off_f, off_d, acc);
      else
        tile_loop<false, kRep>(map, &tile[s][0], count, row0, r, off_f, off_d, acc);
    }
    __syncthreads();  // every lane is done with stage s before TMA overwrites it
    if (tid == 0 && t + kStages < ntiles)
    {
      const uint32_t tn = tile0 + t + kStages;
      const uint32_t cn = min((uint32_t)kT, npad - tn * kT);
      mbar_expect_tx(&full_bar[s], cn * 16u);
      tma_bulk_g2s(&tile[s][0], cloud + (size_t)tn * kT, cn * 16u, &full_bar[s]);
    }
  }

  if (live)
    finish_particle(particles, i, inmap, acc.weight, acc.cnt, npts, px, py, pz, beacons, wr_out);
  if (counter != nullptr)
  {
    unsigned int c = (unsigned int)acc.cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31u) == 0 && c)
      atomicAdd(counter, (unsigned long long)c);
  }
}

// Phase 2 of the two-phase path: thread = particle, block = one particle tile of the staging layout.  The rows of
// the tile are in the caller's cloud order, so the block simply STREAMS its contiguous npad x kThreads bytes front
// to back (TMA bulk copies into a two-stage shared-memory ring) and every thread adds lut[code] of its column in
// order -- the reference's rounding sequence (Grid3d.cpp:292).  Each staged byte is read exactly once.
constexpr int kSumThreads = kThreads;
constexpr int kSumRows = 64;  // rows per ring stage: 64 x 256 B = 16 KB
__global__ void __launch_bounds__(kSumThreads)
    ordered_sum_kernel(const __grid_constant__ MapDev map, float* __restrict__ particles, uint32_t n,
                       const uint8_t* __restrict__ stage, uint32_t npts, uint32_t npad,
                       const BeaconSet* __restrict__ beacons, float* __restrict__ wr_out,
                       unsigned long long* __restrict__ counter)
{
  __shared__ __align__(128) uint8_t ring[2][kSumRows * kSumThreads];
  __shared__ __align__(8) uint64_t full_bar[2];
  __shared__ float lut_s[256];
  const uint32_t tid = threadIdx.x;
  const uint32_t i = blockIdx.x * kSumThreads + tid;
  for (uint32_t k = tid; k < 256; k += kSumThreads)
    lut_s[k] = k < map.lut_n ? map.lut[k] : (k == kStageSkip ? -0.f : 0.f);
  const uint32_t lut_base = (uint32_t)__cvta_generic_to_shared(lut_s);
  if (tid == 0)
  {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* __restrict__ src = stage + (size_t)blockIdx.x * kSumThreads * npad;
  const uint32_t nchunks = (npts + kSumRows - 1) / kSumRows;
  auto issue = [&](uint32_t c) {
    const uint32_t rows = min((uint32_t)kSumRows, npad - c * kSumRows);  // npad is a multiple of 32: bytes % 16 == 0
    mbar_expect_tx(&full_bar[c & 1], rows * kSumThreads);
    tma_bulk_g2s(&ring[c & 1][0], src + (size_t)c * kSumRows * kSumThreads, rows * kSumThreads, &full_bar[c & 1]);
  };
  if (tid == 0)
  {
    if (nchunks > 0)
      issue(0);
    if (nchunks > 1)
      issue(1);
  }
  const bool live = i < n;
  float px = 0, py = 0, pz = 0;
  if (live)
  {
    const float* p = particles + (size_t)7 * i;
    px = p[0];
    py = p[1];
    pz = p[2];
  }
  const bool inmap = live && particle_gate(map, px, py, pz);
  float weight = 0.f;
  int cnt = 0;
  for (uint32_t c = 0; c < nchunks; ++c)
  {
    mbar_wait(&full_bar[c & 1], (c >> 1) & 1u);
    const uint32_t rows = min((uint32_t)kSumRows, npts - c * kSumRows);
    if (inmap)
    {
      uint32_t skipped = 0;
      const uint8_t* col = &ring[c & 1][tid];
#pragma unroll 8
      for (uint32_t rr = 0; rr < rows; ++rr)
      {
        const uint32_t code = col[rr * kSumThreads];
        float v;  // lut_s[code]; lut_s[kStageSkip] = -0.0f, and x + (-0) = x for every x this sum can hold
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(lut_base + code * 4u));
        weight = __fadd_rn(weight, v);  // Grid3d.cpp:292, cloud order
        skipped += (code + 1u) >> 8;  // kStageSkip == 0xFF is the only code that carries into bit 8
      }
      cnt += (int)(rows - skipped);
    }
    __syncthreads();
    if (tid == 0 && c + 2 < nchunks)
      issue(c + 2);
  }
  if (live)
    finish_particle(particles, i, inmap, weight, cnt, npts, px, py, pz, beacons, wr_out);
  if (counter != nullptr)
  {
    unsigned int cc = (unsigned int)cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      cc += __shfl_xor_sync(0xffffffffu, cc, o);
    if ((tid & 31u) == 0 && cc)
      atomicAdd(counter, (unsigned long long)cc);
  }
}

// ---- cloud preparation --------------------------------------------------------------------------------
// 36-bit Morton key of the sensor-frame position on a 0.25 m lattice (+-512 m); non-finite points sort last
__global__ void morton_key_kernel(const float* __restrict__ src, uint32_t stride, uint32_t npts, unsigned long long* __restrict__ keys,
                                  uint32_t* __restrict__ vals)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npts)
    return;
  const float* p = src + (size_t)j * stride;
  unsigned long long key = ~0ull >> 28;
  if (isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))
  {
    key = 0;
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      const float c = fminf(fmaxf(floorf(p[a] * 4.f) + 2048.f, 0.f), 4095.f);
      q[a] = (uint32_t)c;
    }
#pragma unroll
    for (int b = 0; b < 12; ++b)
#pragma unroll
      for (int a = 0; a < 3; ++a)
        key |= (unsigned long long)((q[a] >> b) & 1u) << (3 * b + a);
  }
  keys[j] = key;
  vals[j] = j;
}
// sorted[r] = point order[r], w = its index in the caller's order (as bits); NaN padding rows point at rows >= npts
__global__ void pack_sorted_kernel(const float* __restrict__ src, uint32_t stride, uint32_t npts, const uint32_t* __restrict__ order,
                                   float4* __restrict__ sorted, uint32_t npad)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= npad)
    return;
  if (r < npts)
  {
    const uint32_t o = order[r];
    const float* p = src + (size_t)o * stride;
    sorted[r] = make_float4(p[0], p[1], p[2], __uint_as_float(o));
  }
  else
  {
    const float nanv = CUDART_NAN_F;
    sorted[r] = make_float4(nanv, nanv, nanv, __uint_as_float(r));  // parks its skip code in a padding row (>= npts)
  }
}

__global__ void pack_cloud_kernel(const float* __restrict__ src, uint32_t stride, uint32_t npts, float4* __restrict__ dst,
                                  uint32_t npad)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad)
    return;
  float4 o;
  if (i < npts)
  {
    const float* p = src + (size_t)i * stride;
    o = make_float4(p[0], p[1], p[2], 0.f);
  }
  else
  {
    const float nanv = CUDART_NAN_F;  // a NaN point fails every bounds test: contributes nothing
    o = make_float4(nanv, nanv, nanv, 0.f);
  }
  dst[i] = o;
}

// ---- random-sector gather micro-benchmark (states the gather roofline) ---------------
__device__ __forceinline__ uint64_t splitmix(uint64_t& s)
{
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
    gather_bench_kernel(const float* __restrict__ cells, uint64_t window, uint32_t iters, int mode,
                        float* __restrict__ sink)
{
  const int coherent = mode & 1;
  const int flavor = mode >> 1;  // which load instruction (development probe)
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = gtid * 0x632BE59BD9B4E019ull + 12345;
  float acc = 0.f;
  // coherent: every thread of the grid walks the same slowly moving 4 MB neighbourhood (the access
  // pattern of clustered particles sweeping the cloud in lock-step); else uniform over the window
  const uint64_t blob = 1ull << 20;  // cells
  for (uint32_t it = 0; it < iters; ++it)
  {
    uint64_t idx[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      const uint64_t h = splitmix(s);
      if (coherent)
      {
        const uint64_t centre = ((uint64_t)it * 8u + u) * 4099ull * 32ull;
        idx[u] = (centre + (h % blob)) % window;
      }
      else
        idx[u] = h % window;
      // flavor >= 16: sparse probe -- only every (flavor-th) 128-byte line of the window is touched, so the
      // touched bytes fit L2 while the address range (pages) stays large: separates TLB reach from DRAM
      if (flavor >= 16)
        idx[u] = ((idx[u] >> 5) / (uint64_t)flavor) * (uint64_t)flavor * 32ull + (idx[u] & 31ull);
    }
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      const float* a = cells + idx[u];
      switch (flavor)
      {
        case 1: asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 2: asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 3: asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 4: asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 5: asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 6: asm volatile("ld.global.L1::evict_first.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 7: asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 8: asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;   // L2 prefetch-size hints
        case 9: asm volatile("ld.global.L2::128B.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 10: asm volatile("ld.global.L2::256B.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        case 11: asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(v[u]) : "l"(a)); break;
        default: v[u] = __ldg(a); break;
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      acc += v[u];
  }
  if (acc == 123.456f)
    sink[0] = acc;
}
}  // namespace

// ---- large-set path: pose -> sweep -> ordered sum ---------------------------------------------------------
namespace
{
constexpr int kPoseFloats = 13;  // r[9], off_f[3], flags
constexpr uint32_t kPoseInMap = 1u, kPoseF32 = 2u;

// pose[k * pitch + i]: rotation (Grid3d.cpp:240-250), f32-rounded offsets t - min (Grid3d.cpp:256-258) and the
// flags "passes isIntoMap" / "all three offsets are exactly representable in f32".  Once per particle per update
// instead of once per (particle, cloud chunk).
__global__ void __launch_bounds__(256)
    pose_kernel(const __grid_constant__ MapDev map, const float* __restrict__ particles, uint32_t n, int trig_mode,
                float* __restrict__ pose, uint32_t pitch)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pitch)
    return;
  float px = 0, py = 0, pz = 0, roll = 0, pitch_a = 0, yaw = 0;
  const bool live = i < n;
  if (live)
  {
    const float* p = particles + (size_t)7 * i;
    px = p[0];
    py = p[1];
    pz = p[2];
    roll = p[3];
    pitch_a = p[4];
    yaw = p[5];
  }
  const bool inmap = live && particle_gate(map, px, py, pz);
  float r[9];
  rotation(roll, pitch_a, yaw, trig_mode, r);
  const double off_d[3] = { __dsub_rn((double)px, map.min[0]), __dsub_rn((double)py, map.min[1]),
                            __dsub_rn((double)pz, map.min[2]) };
  bool f32_exact = true;
#pragma unroll
  for (int k = 0; k < 9; ++k)
    pose[(size_t)k * pitch + i] = r[k];
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    // +0.0f turns an offset of -0 into +0: the sum s + off is then never -0 (x + (-x) = +0 in RN), which lets the
    // sweep test 0 <= v < size with one unsigned compare of the bit patterns; -0 and +0 index the same voxel.
    const float f = __fadd_rn(__double2float_rn(off_d[a]), 0.f);
    f32_exact = f32_exact && ((double)f == off_d[a]);
    pose[(size_t)(9 + a) * pitch + i] = f;
  }
  pose[(size_t)12 * pitch + i] = __uint_as_float((inmap ? kPoseInMap : 0u) | (f32_exact ? kPoseF32 : 0u));
}

// what the sweep needs of the map, with the constants of its index arithmetic
struct SweepMap
{
  double min[3];
  double rinv, resol;
  uint32_t su_bits[3];  // bit patterns of size_up: for a float v that is never -0, 0 <= v < size_up  <=>  bits(v) < su_bits
  uint32_t size[3];
  float rinv_f, band;
  uint32_t mul_z, mul_y;  // 128^3-voxel blocks: block = bz * mul_z + by * mul_y + bx
  uint32_t bias;          // the three block coordinates are taken from magic-number floats and each carry the constant
                          // kMagicHi; bias = kMagicHi * (mul_z + mul_y + 1) mod 2^32 removes it
  const uint8_t* codes;
};
constexpr float kMagic = 12582912.f;               // 1.5 * 2^23: u + kMagic rounds u to the nearest integer
constexpr uint32_t kMagicHi = 0x4B400000u >> 7;    // bits(kMagic) >> 7
constexpr uint32_t kQueueCap = 1024;               // deferred exact-tier evaluations per block (mean ~50 at C3)

// Branch-free voxel address of one point for one particle.  The quotient v / resol is formed as
// u = fma(v, rinv_f, -0.5); t = u + 1.5*2^23 rounds it to the nearest integer, which is floor(v / resol) whenever the
// quotient is farther than eps from an integer; the low 22 mantissa bits of t ARE that integer, so the brick / block
// fields are cut straight out of bits(t).  |u - round(u)| >= band flags the rare evaluation that needs the exact tier.
template <bool kF32Offset>
__device__ __forceinline__ void point_address(const SweepMap& sm, const float4 pt, const float r[9], const float off_f[3],
                                              const double off_d[3], uint32_t& off_lo, uint32_t& blk, bool& in, bool& rare)
{
  // Grid3d.cpp:275-277
  const float sx = __fadd_rn(__fadd_rn(__fmul_rn(pt.x, r[0]), __fmul_rn(pt.y, r[1])), __fmul_rn(pt.z, r[2]));
  const float sy = __fadd_rn(__fadd_rn(__fmul_rn(pt.x, r[3]), __fmul_rn(pt.y, r[4])), __fmul_rn(pt.z, r[5]));
  const float sz = __fadd_rn(__fadd_rn(__fmul_rn(pt.x, r[6]), __fmul_rn(pt.y, r[7])), __fmul_rn(pt.z, r[8]));
  const float nx = add_offset<kF32Offset>(sx, off_f[0], off_d[0]);
  const float ny = add_offset<kF32Offset>(sy, off_f[1], off_d[1]);
  const float nz = add_offset<kF32Offset>(sz, off_f[2], off_d[2]);
  // Grid3d.cpp:279-280 (a NaN has a bit pattern above every finite size: fails like in the reference)
  in = (__float_as_uint(nx) < sm.su_bits[0]) & (__float_as_uint(ny) < sm.su_bits[1]) & (__float_as_uint(nz) < sm.su_bits[2]);
  // Grid3d.cpp:282-284, fast tier
  const float ux = __fmaf_rn(nx, sm.rinv_f, -0.5f), uy = __fmaf_rn(ny, sm.rinv_f, -0.5f), uz = __fmaf_rn(nz, sm.rinv_f, -0.5f);
  const float tx = __fadd_rn(ux, kMagic), ty = __fadd_rn(uy, kMagic), tz = __fadd_rn(uz, kMagic);
  const float dev = fmaxf(fmaxf(fabsf(__fsub_rn(ux, __fsub_rn(tx, kMagic))), fabsf(__fsub_rn(uy, __fsub_rn(ty, kMagic)))),
                          fabsf(__fsub_rn(uz, __fsub_rn(tz, kMagic))));
  rare = in & !(dev < sm.band);
  const uint32_t bx = __float_as_uint(tx), by = __float_as_uint(ty), bz = __float_as_uint(tz);
  // within the 2 MB block: brick (z6..2 | y6..2 | x6..3) << 7 | voxel (z1..0 | y1..0 | x2..0), see coded_index()
  const uint32_t wx = (bx & 0x7u) | ((bx << 4) & 0x780u);
  const uint32_t wy = ((by << 3) & 0x18u) | ((by << 9) & 0xF800u);
  const uint32_t wz = ((bz << 5) & 0x60u) | ((bz << 14) & 0x1F0000u);
  off_lo = wx | wy | wz;
  blk = (bz >> 7) * sm.mul_z + ((by >> 7) * sm.mul_y + ((bx >> 7) - sm.bias));
}

// exact evaluation of one (particle, point) pair: every tier of the index, all bounds tests -- the deferred path
__device__ __noinline__ uint32_t exact_code(const SweepMap& sm, const float4 pt, const float* __restrict__ pose, uint32_t pitch,
                                            uint32_t i, const float* __restrict__ particles)
{
  float r[9];
#pragma unroll
  for (int k = 0; k < 9; ++k)
    r[k] = pose[(size_t)k * pitch + i];
  const float sx = __fadd_rn(__fadd_rn(__fmul_rn(pt.x, r[0]), __fmul_rn(pt.y, r[1])), __fmul_rn(pt.z, r[2]));
  const float sy = __fadd_rn(__fadd_rn(__fmul_rn(pt.x, r[3]), __fmul_rn(pt.y, r[4])), __fmul_rn(pt.z, r[5]));
  const float sz = __fadd_rn(__fadd_rn(__fmul_rn(pt.x, r[6]), __fmul_rn(pt.y, r[7])), __fmul_rn(pt.z, r[8]));
  // the f64 offset add of the reference itself (Grid3d.cpp:256-258, 275-277); equals the f32 add when that is exact
  const float* p = particles + (size_t)7 * i;
  const float nx = __double2float_rn(__dadd_rn((double)sx, __dsub_rn((double)p[0], sm.min[0])));
  const float ny = __double2float_rn(__dadd_rn((double)sy, __dsub_rn((double)p[1], sm.min[1])));
  const float nz = __double2float_rn(__dadd_rn((double)sz, __dsub_rn((double)p[2], sm.min[2])));
  const bool in = (nx >= 0.f) & (__float_as_uint(__fadd_rn(nx, 0.f)) < sm.su_bits[0]) & (ny >= 0.f) &
                  (__float_as_uint(__fadd_rn(ny, 0.f)) < sm.su_bits[1]) & (nz >= 0.f) &
                  (__float_as_uint(__fadd_rn(nz, 0.f)) < sm.su_bits[2]);
  if (!in)
    return kStageSkip;
  const uint32_t ix = voxel_exact(nx, sm.rinv, sm.resol), iy = voxel_exact(ny, sm.rinv, sm.resol),
                 iz = voxel_exact(nz, sm.rinv, sm.resol);
  if (!((ix < sm.size[0]) & (iy < sm.size[1]) & (iz < sm.size[2])))  // Grid3d.cpp:286 (and :290)
    return kStageSkip;
  return gather_u8(sm.codes + coded_index<kRepCode8>(ix, iy, iz, sm.mul_y, sm.mul_z / sm.mul_y));
}

template <bool kF32Offset, bool kIdx32>
__device__ __forceinline__ void sweep_tile(const SweepMap& sm, const float4* __restrict__ tile, uint32_t count, const float r[9],
                                           const float off_f[3], const double off_d[3], uint8_t* __restrict__ column,
                                           uint32_t tid, uint32_t* __restrict__ q_count, uint32_t* __restrict__ queue)
{
  // software pipeline as in tile_loop: the gathers of batch k+1 are in flight while batch k is parked
  uint32_t raw_a[kUnroll], raw_b[kUnroll], row_a[kUnroll], row_b[kUnroll];
  auto issue = [&](uint32_t base, uint32_t raw[kUnroll], uint32_t row[kUnroll]) {
    float4 pt[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      pt[u] = tile[base + u];  // broadcast LDS.128
    uint32_t rare_mask = 0;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
    {
      uint32_t off_lo, blk;
      bool in, rare;
      point_address<kF32Offset>(sm, pt[u], r, off_f, off_d, off_lo, blk, in, rare);
      row[u] = __float_as_uint(pt[u].w);
      raw[u] = kStageSkip;
      rare_mask |= rare ? (1u << u) : 0u;
      if (in & !rare)
      {
        if (kIdx32)
          raw[u] = gather_u8(sm.codes + ((blk << 21) | off_lo));
        else
          raw[u] = gather_u8(sm.codes + (((uint64_t)blk << 21) | off_lo));
      }
    }
    if (rare_mask)  // a handful per block: parked as "skip" now, redone exactly after the sweep (one branch per batch)
    {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if ((rare_mask >> u) & 1u)
        {
          const uint32_t slot = atomicAdd(q_count, 1u);
          if (slot < kQueueCap)
            queue[slot] = (tid << 16) | (base + u);
        }
    }
  };
  auto park = [&](const uint32_t raw[kUnroll], const uint32_t row[kUnroll]) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      column[(size_t)row[u] * kThreads] = (uint8_t)raw[u];
  };
  issue(0, raw_a, row_a);
#pragma unroll 1
  for (uint32_t base = kUnroll; base < count; base += 2 * kUnroll)
  {
    issue(base, raw_b, row_b);
    park(raw_a, row_a);
    if (base + kUnroll < count)
      issue(base + kUnroll, raw_a, row_a);
    park(raw_b, row_b);
  }
}

// Phase 1 of the large-set path.  Thread = particle, block = 256 particles x one 128-point chunk of the Morton-sorted
// cloud, grid in CHUNK-MAJOR order: the hardware block scheduler keeps every resident block inside the same one or two
// chunks, i.e. all particles look up the same neighbourhood of the map at the same time (no flags, nothing to chain --
// the codes are only parked at the point's original row of stage[particle tile][row][lane]).
template <bool kIdx32>
__global__ void __launch_bounds__(kThreads, AMCL3D_MINBLOCKS)
    sweep_kernel(const __grid_constant__ SweepMap sm, const float* __restrict__ pose, uint32_t pitch,
                 const float* __restrict__ particles, const float4* __restrict__ sorted, uint32_t npad,
                 uint8_t* __restrict__ stage)
{
  __shared__ __align__(128) float4 tile[kStageTile];
  __shared__ __align__(8) uint64_t full_bar;
  __shared__ uint32_t q_count;
  __shared__ uint32_t queue[kQueueCap];
  const uint32_t tid = threadIdx.x;
  const uint32_t n_ptiles = pitch / kThreads;
  const uint32_t ptile = blockIdx.x % n_ptiles, chunk = blockIdx.x / n_ptiles;
  const uint32_t i = ptile * kThreads + tid;
  const uint32_t row0 = chunk * kStageTile;
  const uint32_t count = min((uint32_t)kStageTile, npad - row0);  // npad is a multiple of 32
  if (tid == 0)
  {
    mbar_init(&full_bar, 1);
    q_count = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0)
  {
    mbar_expect_tx(&full_bar, count * 16u);
    tma_bulk_g2s(&tile[0], sorted + row0, count * 16u, &full_bar);
  }
  // the particle's pose while the cloud chunk is in flight
  float r[9], off_f[3];
#pragma unroll
  for (int k = 0; k < 9; ++k)
    r[k] = pose[(size_t)k * pitch + i];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    off_f[a] = pose[(size_t)(9 + a) * pitch + i];
  const uint32_t flags = __float_as_uint(pose[(size_t)12 * pitch + i]);
  const bool inmap = flags & kPoseInMap;
  // warp-uniform choice of the loop flavour; lanes that are not in the map do not vote it down
  const bool warp_f32 = __all_sync(0xffffffffu, (flags & kPoseF32) || !inmap);
  uint8_t* const column = stage + (size_t)ptile * kThreads * npad + tid;
  mbar_wait(&full_bar, 0);
  if (inmap)
  {
    double off_d[3] = { 0, 0, 0 };
    if (warp_f32)
      sweep_tile<true, kIdx32>(sm, tile, count, r, off_f, off_d, column, tid, &q_count, queue);
    else
    {
      const float* p = particles + (size_t)7 * i;
#pragma unroll
      for (int a = 0; a < 3; ++a)
        off_d[a] = __dadd_rn(__dsub_rn((double)p[a], sm.min[a]), 0.0);  // Grid3d.cpp:256-258 (never -0, see pose_kernel)
      sweep_tile<false, kIdx32>(sm, tile, count, r, off_f, off_d, column, tid, &q_count, queue);
    }
  }
  __syncthreads();
  // deferred exact tier: evaluations within eps of a voxel face.  Kept out of the sweep so that its loop holds no call
  // (a call site makes the compiler reload every uniform register after it, ~7 % of the sweep's instructions).
  const uint32_t pending = q_count;
  if (pending <= kQueueCap)
  {
    for (uint32_t e = tid; e < pending; e += kThreads)
    {
      const uint32_t q = queue[e], owner = q >> 16, j = q & 0xFFFFu;
      const float4 pt = tile[j];
      const uint32_t code = exact_code(sm, pt, pose, pitch, ptile * kThreads + owner, particles);
      stage[(size_t)ptile * kThreads * npad + (size_t)__float_as_uint(pt.w) * kThreads + owner] = (uint8_t)code;
    }
  }
  else if (inmap)
  {
    // queue overflow (e.g. a cloud and poses that sit on the voxel lattice): this thread redoes its whole column exactly
    for (uint32_t j = 0; j < count; ++j)
    {
      const float4 pt = tile[j];
      column[(size_t)__float_as_uint(pt.w) * kThreads] = (uint8_t)exact_code(sm, pt, pose, pitch, i, particles);
    }
  }
}
}  // namespace

size_t pose_floats(uint32_t n)
{
  return (size_t)kPoseFloats * (((size_t)n + kThreads - 1) / kThreads * kThreads);
}

// true when the branch-free sweep index arithmetic is valid for this map: sizes below 2^21 per axis, and for the largest
// coordinate that passes the bounds test the reference's floor(v / resol) already stays below the size on every axis
// (so Grid3d.cpp:286 never rejects anything the bounds test accepted and the sweep can leave it out)
static bool sweep_map(const MapDev& map, SweepMap* sm)
{
  if (map.rep != kRepCode8 || map.codes == nullptr || map.lut_n >= kStageSkip)
    return false;
  for (int a = 0; a < 3; ++a)
  {
    if (map.size[a] == 0 || map.size[a] >= (1u << 21) || !(map.size_up[a] > 0.f) || !isfinite(map.size_up[a]))
      return false;
    const float vmax = nextafterf(map.size_up[a], 0.f);
    if (!(floor((double)vmax / map.resol) < (double)map.size[a]))
      return false;
    sm->min[a] = map.min[a];
    memcpy(&sm->su_bits[a], &map.size_up[a], 4);
    sm->size[a] = map.size[a];
  }
  sm->rinv = map.rinv;
  sm->resol = map.resol;
  sm->rinv_f = map.rinv_f;
  sm->band = map.band;
  sm->mul_y = map.nbx;
  sm->mul_z = map.nbx * map.nby;
  sm->bias = kMagicHi * (sm->mul_z + sm->mul_y + 1u);
  sm->codes = reinterpret_cast<const uint8_t*>(map.codes);
  return true;
}

bool two_phase_supported(const MapDev& map)
{
  SweepMap sm;
  return sweep_map(map, &sm);
}

cudaError_t launch_pose(const MapDev& map, const float* particles, uint32_t n, int trig_mode, float* pose, cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  const uint32_t pitch = (n + kThreads - 1) / kThreads * kThreads;
  pose_kernel<<<pitch / 256, 256, 0, stream>>>(map, particles, n, trig_mode, pose, pitch);
  return cudaGetLastError();
}

cudaError_t launch_update(const MapDev& map, float* particles, uint32_t n, const float4* cloud4, uint32_t npts,
                          uint32_t npts_padded, const BeaconSet* beacons, float* wr_out, int trig_mode,
                          unsigned long long* counter, const TwoPhase* tp, cudaStream_t stream)
{
  if (n == 0)
    return cudaSuccess;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  if (map.rep == kRepFloat || map.codes == nullptr || map.lut_n > kMaxLutSmem)
  {
    update_kernel<kRepFloat><<<blocks, kThreads, 0, stream>>>(map, particles, n, cloud4, npts, npts_padded, beacons,
                                                                     wr_out, trig_mode, counter);
    return cudaGetLastError();
  }
  SweepMap sm;
  if (tp != nullptr && tp->stage != nullptr && tp->pose != nullptr && sweep_map(map, &sm))
  {
    // phase 1: sweep the Morton-sorted cloud, park the codes; phase 2: ordered sum.  tp->pose holds this slice's poses
    // (pitch = tp->pose_pitch of the whole set, the slice starts at a multiple of the tile width).
    const uint32_t chunks = (npts_padded + kStageTile - 1) / kStageTile;
    // n_ptiles of THIS slice: the kernel derives it from the pitch it is given, so hand it the slice's own pitch and
    // address the SoA rows of the full array through a per-row stride
    const uint32_t slice_pitch = blocks * kThreads;
    if (tp->pose_pitch != slice_pitch)
      return cudaErrorInvalidValue;  // slices carry their own pose block (see update_impl)
    const uint64_t coded_bytes = ((uint64_t)map.nbx * map.nby * ((map.size[2] + 127) / 128)) << 21;
    if (coded_bytes <= (1ull << 32))
      sweep_kernel<true><<<blocks * chunks, kThreads, 0, stream>>>(sm, tp->pose, slice_pitch, particles, tp->sorted,
                                                                    npts_padded, tp->stage);
    else
      sweep_kernel<false><<<blocks * chunks, kThreads, 0, stream>>>(sm, tp->pose, slice_pitch, particles, tp->sorted,
                                                                     npts_padded, tp->stage);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      return e;
    if (tp->ev_mid)
      cudaEventRecord(tp->ev_mid, stream);
    ordered_sum_kernel<<<(n + kSumThreads - 1) / kSumThreads, kSumThreads, 0, stream>>>(
        map, particles, n, tp->stage, npts, npts_padded, beacons, wr_out, counter);
    if (tp->ev_end)
      cudaEventRecord(tp->ev_end, stream);
    return cudaGetLastError();
  }
  const size_t dyn = sizeof(float) * map.lut_n;
  if (map.rep == kRepCode8)
    update_kernel<kRepCode8><<<blocks, kThreads, dyn, stream>>>(map, particles, n, cloud4, npts, npts_padded, beacons,
                                                                       wr_out, trig_mode, counter);
  else
  {
    // the opt-in is a per-device attribute of the function: set it on every launch (cheap) rather than once per process
    cudaError_t e = cudaFuncSetAttribute(update_kernel<kRepCode16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(float) * kMaxLutSmem));
    if (e != cudaSuccess)
      return e;
    update_kernel<kRepCode16><<<blocks, kThreads, dyn, stream>>>(map, particles, n, cloud4, npts, npts_padded, beacons,
                                                                        wr_out, trig_mode, counter);
  }
  return cudaGetLastError();
}

cudaError_t launch_sort_cloud(const float* src, uint32_t stride, uint32_t npts, uint32_t npts_padded, float4* sorted,
                              void* scratch, size_t scratch_bytes, cudaStream_t stream)
{
  // scratch: keys_in, keys_out (8 B each), vals_in, vals_out (4 B each), then the CUB temporaries
  if (npts_padded == 0)
    return cudaSuccess;
  char* base = reinterpret_cast<char*>(scratch);
  const size_t n8 = ((size_t)npts * 8 + 255) & ~(size_t)255, n4 = ((size_t)npts * 4 + 255) & ~(size_t)255;
  unsigned long long* k_in = reinterpret_cast<unsigned long long*>(base);
  unsigned long long* k_out = reinterpret_cast<unsigned long long*>(base + n8);
  uint32_t* v_in = reinterpret_cast<uint32_t*>(base + 2 * n8);
  uint32_t* v_out = reinterpret_cast<uint32_t*>(base + 2 * n8 + n4);
  void* tmp = base + 2 * n8 + 2 * n4;
  size_t tmp_bytes = scratch_bytes - (2 * n8 + 2 * n4);
  if (npts > 0)
  {
    morton_key_kernel<<<(npts + 255) / 256, 256, 0, stream>>>(src, stride, npts, k_in, v_in);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int)npts, 0, 36, stream);
    if (e != cudaSuccess)
      return e;
  }
  pack_sorted_kernel<<<(npts_padded + 255) / 256, 256, 0, stream>>>(src, stride, npts, v_out, sorted, npts_padded);
  return cudaGetLastError();
}

uint32_t stage_tile_width()
{
  return (uint32_t)kThreads;
}

size_t sort_cloud_scratch_bytes(uint32_t npts)
{
  const size_t n8 = ((size_t)npts * 8 + 255) & ~(size_t)255, n4 = ((size_t)npts * 4 + 255) & ~(size_t)255;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)(npts ? npts : 1), 0, 36);
  return 2 * n8 + 2 * n4 + tmp + 256;
}

cudaError_t launch_pack_cloud(const float* src, uint32_t stride, uint32_t npts, float4* dst, uint32_t npts_padded,
                              cudaStream_t stream)
{
  if (npts_padded == 0)
    return cudaSuccess;
  pack_cloud_kernel<<<(npts_padded + 255) / 256, 256, 0, stream>>>(src, stride, npts, dst, npts_padded);
  return cudaGetLastError();
}

cudaError_t launch_gather_bench(const float* cells, uint64_t window, uint32_t blocks, uint32_t iters, int coherent,
                                float* sink, cudaStream_t stream)
{
  gather_bench_kernel<<<blocks, 256, 0, stream>>>(cells, window, iters, coherent, sink);
  return cudaGetLastError();
}

}  // namespace a3d
