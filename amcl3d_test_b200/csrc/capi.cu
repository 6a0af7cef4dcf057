// capi.cu -- the C-ABI (include/amcl3d_b200.h): opaque handles that own device memory and launch the
// sm_100a kernels of update.cu / pfops.cu / gridbuild.cu.  No CPU fallback: without a usable CUDA
// device every compute entry point returns AMCL3D_B200_ERR_CUDA.
#include "common.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

using namespace a3d;

namespace
{
thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
  g_err = msg;
  return code;
}
int fail_cuda(cudaError_t e, const char* where)
{
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  (void)cudaGetLastError();
  return AMCL3D_B200_ERR_CUDA;
}
#define CU(expr)                         \
  do                                     \
  {                                      \
    cudaError_t e__ = (expr);            \
    if (e__ != cudaSuccess)              \
      return fail_cuda(e__, #expr);      \
  } while (0)

template <typename T>
cudaError_t ensure(T*& ptr, uint64_t& cap, uint64_t need)
{
  if (need <= cap && ptr != nullptr)
    return cudaSuccess;
  if (ptr)
    cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
  const uint64_t want = std::max<uint64_t>(need, 1);
  cudaError_t e = cudaMalloc(&ptr, sizeof(T) * want);
  if (e == cudaSuccess)
    cap = want;
  return e;
}
// particle buffers: capacity counted in particles, 7 floats each
cudaError_t ensure_particles(float*& ptr, uint64_t& cap, uint64_t need)
{
  if (need <= cap && ptr != nullptr)
    return cudaSuccess;
  if (ptr)
    cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
  const uint64_t want = std::max<uint64_t>(need, 1);
  cudaError_t e = cudaMalloc(&ptr, sizeof(float) * 7 * want);
  if (e == cudaSuccess)
    cap = want;
  return e;
}
}  // namespace

// NVTX ranges named after the reference's own tic/toc timers ("Update" ParticleFilter.cpp:117,138; "PFMove" / "Resample"
// amcl3d.cpp:185-189,194-202): a timeline of the drop-in reads like the reference's log
struct NvtxRange
{
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct amcl3d_b200_grid_s
{
  int device = 0;
  MapDev map{};
  double sensor_dev = 0;
  float* d_prob = nullptr;
  uint64_t prob_cap = 0;
  int32_t* d_edt = nullptr;
  uint64_t edt_cells = 0;
  CodedGrid coded{};       // lossless dictionary-coded copy (rep == kRepFloat: absent)
  int use_codes = 1;       // AMCL3D_B200_CODES=0 or grid_set_option disables it
  int edt_fast = 1;        // AMCL3D_B200_EDT_FAST=0 or amcl3d_b200_grid_set_edt_fast: exact envelope passes always
  cudaStream_t stream = nullptr;
  amcl3d_b200_pf_t scratch_pf = nullptr;  // for the single-pose computeCloudWeight
  cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };  // two timed regions around the kernels of the last build
  bool build_timed = false;
};

struct amcl3d_b200_pf_s
{
  amcl3d_b200_grid_t grid = nullptr;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int trig_mode = AMCL3D_B200_TRIG_F64;
  // p_
  float* d_p = nullptr;
  uint64_t n = 0, cap = 0;
  bool external = false;
  float* d_alt = nullptr;
  uint64_t alt_cap = 0;
  // cloud
  float4* d_cloud = nullptr;
  uint64_t cloud_cap = 0;
  uint32_t npts = 0, npad = 0;
  float* d_stage = nullptr;
  uint64_t stage_cap = 0;
  // input downsample (voxelfilter.cu): centroids (x y z intensity) + workspace
  float4* d_filtered = nullptr;
  uint64_t filtered_cap = 0;
  char* d_vf_scratch = nullptr;
  uint64_t vf_scratch_cap = 0;
  bool filtered_valid = false;
  // two-phase sweep: Morton-sorted cloud, sort scratch, staged codes
  float4* d_sorted = nullptr;
  uint64_t sorted_cap = 0;
  char* d_sort_scratch = nullptr;
  uint64_t sort_scratch_cap = 0;
  uint8_t* d_codes_staged = nullptr;
  uint64_t codes_staged_cap = 0;
  float* d_pose = nullptr;  // per-particle rotation + offsets of the current slice (pose_kernel)
  uint64_t pose_cap = 0;
  int two_phase = 1;  // AMCL3D_B200_TWO_PHASE=0 / amcl3d_b200_pf_set_two_phase disables
  bool sorted_valid = false;  // d_sorted holds the Morton order of the cloud currently in d_cloud
  int chain_mode = 0;         // amcl3d_b200_pf_set_chain_mode
  // scratch
  float* d_prefix = nullptr;
  uint64_t prefix_cap = 0;
  float* d_thr = nullptr;
  uint64_t thr_cap = 0;
  float* d_wr = nullptr;
  uint64_t wr_cap = 0;
  float* d_dense = nullptr;  // dense copy of the weights for the scan
  uint64_t dense_cap = 0;
  int32_t* d_idx = nullptr;
  uint64_t idx_cap = 0;
  float* d_noise = nullptr;
  uint64_t noise_cap = 0;
  float* d_scalars = nullptr;  // [0] total [1] total_r [2] max [8..14] out7 [16..22] exact-mean chain state of this rank (sharded)
  unsigned long long* d_u64 = nullptr;  // [0] counter [1] argmax
  int* d_int = nullptr;
  double* d_delta = nullptr;
  BeaconSet* d_beacons = nullptr;
  void* d_scratch = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_sync = nullptr;  // orders this handle's stream against another handle's (shard gather / scatter)
  cudaEvent_t ev_phase[4] = { nullptr, nullptr, nullptr, nullptr };  // pose | sweep | ordered sum of the last large-set update
  bool phase_timed = false;
  bool timed = false;
  bool async_mode = false;  // amcl3d_b200_pf_set_async: calls that hand nothing back to the host do not synchronise
  uint64_t launches = 0;
  // sharded particle set over peer memory (shard.cu)
  struct Shard
  {
    bool on = false, attached = false;
    uint32_t rank = 0, world = 1;
    uint64_t n_total = 0, n_cap = 0, n_max = 0, lo = 0, hi = 0;  // n_cap: the total the region is sized for (n_total <= n_cap)
    char* region = nullptr;
    size_t region_bytes = 0;
    ShardPeers peers{};
    void* opened[kMaxShardWorld] = {};  // cudaIpcOpenMemHandle mappings to close
    uint32_t epoch = 0, mean_epoch = 0, fast_epoch = 0;
    uint32_t cur = 0;  // P[cur] holds p_
    unsigned int* d_done = nullptr;
    uint32_t* d_error = nullptr;
    float* d_part7 = nullptr;
  } shard;
};

namespace
{
void refresh_map(amcl3d_b200_grid_t g)
{
  MapDev& m = g->map;
  m.rinv = 1.0 / m.resol;
  m.rinv_f = (float)(1.0 / m.resol);
  for (int a = 0; a < 3; ++a)
  {
    const double S = m.max[a] - m.min[a];  // Grid3d.cpp:252-254
    float f = (float)S;
    if ((double)f < S)
      f = nextafterf(f, INFINITY);
    m.size_up[a] = f;
    // |fl32(v*rinv_f) - v/resol| <= (v/resol) * 2^-23 (1 + 2^-24); v/resol < size + 1
    float eps = ((float)m.size[a] + 2.f) * 1.25e-7f + 2e-7f;
    if (!(eps < 0.5f))
      eps = 0.5f;
    m.eps[a] = eps;
  }
  m.band = 0.5f - fmaxf(fmaxf(m.eps[0], m.eps[1]), m.eps[2]);
  m.step_z = (uint64_t)m.size[0] * (uint64_t)m.size[1];
  m.n_cells = (uint64_t)m.size[0] * (uint64_t)m.size[1] * (uint64_t)m.size[2];
  m.prob = g->d_prob;
  const bool coded = g->use_codes && g->coded.rep != kRepFloat && g->coded.codes != nullptr;
  m.rep = coded ? g->coded.rep : kRepFloat;
  m.codes = coded ? g->coded.codes : nullptr;
  m.lut = coded ? g->coded.lut : nullptr;
  m.lut_n = coded ? g->coded.lut_n : 0;
  m.nbx = coded ? g->coded.nbx : 0;
  m.nby = coded ? g->coded.nby : 0;
}

void drop_codes(amcl3d_b200_grid_t g)
{
  cudaFree(g->coded.codes_alloc);
  cudaFree(g->coded.lut);
  g->coded = CodedGrid{};
}

// PointCloudTools.cpp:120-125
void size_from_bounds(amcl3d_b200_grid_t g)
{
  MapDev& m = g->map;
  for (int a = 0; a < 3; ++a)
    m.size[a] = static_cast<uint32_t>(ceil((m.max[a] - m.min[a]) / m.resol));
}

int use_device(int device)
{
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess)
    return fail_cuda(e, "cudaSetDevice");
  return AMCL3D_B200_OK;
}

// end of an entry point that returns nothing to the host: the reference API is synchronous, so is this one unless
// the caller opted into stream order
int finish_call(amcl3d_b200_pf_t pf)
{
  if (pf->async_mode)
    return AMCL3D_B200_OK;
  cudaError_t e = cudaStreamSynchronize(pf->stream);
  if (e != cudaSuccess)
    return fail_cuda(e, "cudaStreamSynchronize");
  return AMCL3D_B200_OK;
}

struct Timed
{
  amcl3d_b200_pf_t pf;
  explicit Timed(amcl3d_b200_pf_t p) : pf(p)
  {
    cudaEventRecord(pf->ev0, pf->stream);
  }
  ~Timed()
  {
    cudaEventRecord(pf->ev1, pf->stream);
    pf->timed = true;
  }
};
}  // namespace

extern "C" {

int amcl3d_b200_abi_version(void)
{
  return AMCL3D_B200_ABI_VERSION;
}

const char* amcl3d_b200_last_error(void)
{
  return g_err.c_str();
}

int amcl3d_b200_device_info(int device, int* device_count, int* sm_count, int* cc_major, int* cc_minor,
                            uint64_t* l2_bytes, uint64_t* mem_bytes)
{
  int cnt = 0;
  CU(cudaGetDeviceCount(&cnt));
  if (device_count)
    *device_count = cnt;
  if (device < 0 || device >= cnt)
    return fail(AMCL3D_B200_ERR_ARG, "device index out of range");
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, device));
  if (sm_count)
    *sm_count = p.multiProcessorCount;
  if (cc_major)
    *cc_major = p.major;
  if (cc_minor)
    *cc_minor = p.minor;
  if (l2_bytes)
    *l2_bytes = (uint64_t)p.l2CacheSize;
  if (mem_bytes)
    *mem_bytes = (uint64_t)p.totalGlobalMem;
  return AMCL3D_B200_OK;
}

// page-locks a host buffer the caller keeps handing to set_particles / get_particles (cudaHostRegister), so that those
// copies are single DMA transfers instead of staged pageable copies
int amcl3d_b200_host_pin(void* ptr, size_t bytes)
{
  if (!ptr || !bytes)
    return fail(AMCL3D_B200_ERR_ARG, "NULL buffer");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered)
  {
    (void)cudaGetLastError();
    cudaHostUnregister(ptr);
    e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  }
  if (e != cudaSuccess)
    return fail_cuda(e, "cudaHostRegister");
  return AMCL3D_B200_OK;
}

int amcl3d_b200_host_unpin(void* ptr)
{
  if (!ptr)
    return AMCL3D_B200_OK;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess)
  {
    (void)cudaGetLastError();
    return fail_cuda(e, "cudaHostUnregister");
  }
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_set_chain_mode(amcl3d_b200_pf_t pf, int mode)
{
  if (!pf || mode < 0 || mode > 2)
    return fail(AMCL3D_B200_ERR_ARG, "chain mode must be 0, 1 or 2");
  pf->chain_mode = mode;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_set_edt_fast(amcl3d_b200_grid_t g, int enable)
{
  if (!g)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  g->edt_fast = enable != 0;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_set_l2_fetch_granularity(int device, int bytes)
{
  if (bytes != 32 && bytes != 64 && bytes != 128)
    return fail(AMCL3D_B200_ERR_ARG, "granularity must be 32, 64 or 128");
  if (int rc = use_device(device))
    return rc;
  CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_get_l2_fetch_granularity(int device, int* bytes)
{
  if (!bytes)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (int rc = use_device(device))
    return rc;
  size_t v = 0;
  CU(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
  *bytes = (int)v;
  return AMCL3D_B200_OK;
}

/* ============================================ Grid3d ============================================ */

int amcl3d_b200_grid_create(int device, amcl3d_b200_grid_t* out)
{
  if (!out)
    return fail(AMCL3D_B200_ERR_ARG, "out is NULL");
  *out = nullptr;
  int cnt = 0;
  CU(cudaGetDeviceCount(&cnt));
  if (device < 0 || device >= cnt)
    return fail(AMCL3D_B200_ERR_ARG, "device index out of range");
  if (int rc = use_device(device))
    return rc;
  amcl3d_b200_grid_t g = new (std::nothrow) amcl3d_b200_grid_s();
  if (!g)
    return fail(AMCL3D_B200_ERR_ARG, "out of host memory");
  g->device = device;
  if (const char* env = getenv("AMCL3D_B200_EDT_FAST"))
    g->edt_fast = atoi(env) != 0;
  if (const char* env = getenv("AMCL3D_B200_CODES"))
    g->use_codes = atoi(env) != 0;
  // The L2 fetch granularity is a per-device limit of the HOST application: it is only touched when asked for
  // (AMCL3D_B200_L2_FETCH=32|64|128 or amcl3d_b200_set_l2_fetch_granularity).  Measured on B200: no effect on the
  // DRAM bytes a scattered 1-byte gather costs (profiles/r2_gather_microbench.json).
  if (const char* env = getenv("AMCL3D_B200_L2_FETCH"))
  {
    const int gran = atoi(env);
    if (gran == 32 || gran == 64 || gran == 128)
      (void)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
    (void)cudaGetLastError();
  }
  cudaError_t e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
  for (int k = 0; k < 4 && e == cudaSuccess; ++k)
    e = cudaEventCreate(&g->ev[k]);
  if (e != cudaSuccess)
  {
    amcl3d_b200_grid_destroy(g);
    return fail_cuda(e, "cudaStreamCreate");
  }
  *out = g;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_destroy(amcl3d_b200_grid_t g)
{
  if (!g)
    return AMCL3D_B200_OK;
  cudaSetDevice(g->device);
  if (g->scratch_pf)
    amcl3d_b200_pf_destroy(g->scratch_pf);
  cudaFree(g->d_prob);
  cudaFree(g->d_edt);
  drop_codes(g);
  for (int k = 0; k < 4; ++k)
    if (g->ev[k])
      cudaEventDestroy(g->ev[k]);
  if (g->stream)
    cudaStreamDestroy(g->stream);
  delete g;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_get_info(amcl3d_b200_grid_t g, amcl3d_b200_grid_info_t* info)
{
  if (!g || !info)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  for (int a = 0; a < 3; ++a)
  {
    info->min[a] = g->map.min[a];
    info->max[a] = g->map.max[a];
    info->size[a] = g->map.size[a];
  }
  info->resol = g->map.resol;
  info->sensor_dev = g->sensor_dev;
  info->n_cells = g->map.n_cells;
  info->has_map = g->map.has_map;
  info->has_grid = g->map.has_grid;
  info->device = g->device;
  info->rep = g->map.rep;
  info->lut_n = g->map.lut_n;
  info->coded_bytes = g->map.rep != kRepFloat ? g->coded.bytes : 0;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_use_codes(amcl3d_b200_grid_t g, int enable)
{
  if (!g)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  g->use_codes = enable != 0;
  refresh_map(g);
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_set_map(amcl3d_b200_grid_t g, const double mn[3], const double mx[3], double resol)
{
  if (!g || !mn || !mx || !(resol > 0))
    return fail(AMCL3D_B200_ERR_ARG, "bad map bounds / resolution");
  for (int a = 0; a < 3; ++a)
  {
    g->map.min[a] = mn[a];
    g->map.max[a] = mx[a];
  }
  g->map.resol = resol;
  g->map.has_map = 1;
  if (!g->map.has_grid)
    size_from_bounds(g);
  refresh_map(g);
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_compute_bounds(amcl3d_b200_grid_t g, const float* xyz, uint32_t stride, uint64_t n, double resol,
                                    double mn[3], double mx[3])
{
  if (!g || !xyz || stride < 3 || !(resol > 0))
    return fail(AMCL3D_B200_ERR_ARG, "bad cloud");
  if (n <= 1)
    return fail(AMCL3D_B200_ERR_EMPTY, "OcTree is empty");  // PointCloudTools.cpp:89-91
  if (int rc = use_device(g->device))
    return rc;
  float* d_xyz = nullptr;
  float* d_keys = nullptr;
  CU(cudaMalloc(&d_xyz, sizeof(float) * n * stride));
  cudaError_t e = cudaMalloc(&d_keys, 6 * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_xyz, xyz, sizeof(float) * n * stride, cudaMemcpyHostToDevice, g->stream);
  if (e == cudaSuccess)
    e = launch_bounds(d_xyz, stride, n, d_keys, g->stream);
  uint32_t keys[6];
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(keys, d_keys, sizeof(keys), cudaMemcpyDeviceToHost, g->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(g->stream);
  cudaFree(d_xyz);
  cudaFree(d_keys);
  if (e != cudaSuccess)
    return fail_cuda(e, "bounds reduction");
  double lo[3], hi[3];
  for (int a = 0; a < 6; ++a)
  {
    const uint32_t k = keys[a];
    const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    float f;
    memcpy(&f, &b, 4);
    if (a < 3)
      lo[a] = f;
    else
      hi[a - 3] = f;
  }
  g->map.has_grid = 0;
  int rc = amcl3d_b200_grid_set_map(g, lo, hi, resol);
  if (rc)
    return rc;
  for (int a = 0; a < 3; ++a)
  {
    if (mn)
      mn[a] = lo[a];
    if (mx)
      mx[a] = hi[a];
  }
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_upload(amcl3d_b200_grid_t g, const float* prob, const uint32_t size[3], double sensor_dev)
{
  if (!g || !prob || !size)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (int rc = use_device(g->device))
    return rc;
  const uint64_t cells = (uint64_t)size[0] * size[1] * size[2];
  CU(ensure(g->d_prob, g->prob_cap, cells));
  CU(cudaMemcpyAsync(g->d_prob, prob, sizeof(float) * cells, cudaMemcpyHostToDevice, g->stream));
  CU(cudaStreamSynchronize(g->stream));
  for (int a = 0; a < 3; ++a)
    g->map.size[a] = size[a];
  g->sensor_dev = sensor_dev;
  g->map.has_grid = 1;
  drop_codes(g);  // an uploaded float grid is used as is (linear f32)
  refresh_map(g);
  return AMCL3D_B200_OK;
}

// replica of a built / loaded grid on another device of the box (the grid is replicated, the particles are
// sharded).  The coded cells travel as one peer copy; the linear float cells only on request (or when there are no codes).
int amcl3d_b200_grid_clone(amcl3d_b200_grid_t src, int device, int with_float, amcl3d_b200_grid_t* out)
{
  if (!src || !out)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (!src->map.has_map || !src->map.has_grid)
    return fail(AMCL3D_B200_ERR_STATE, "no grid to replicate");
  amcl3d_b200_grid_t g = nullptr;
  if (int rc = amcl3d_b200_grid_create(device, &g))
    return rc;
  cudaError_t e = cudaSuccess;
  g->map = src->map;
  g->sensor_dev = src->sensor_dev;
  g->use_codes = src->use_codes;
  g->edt_fast = src->edt_fast;
  g->d_prob = nullptr;
  g->coded = CodedGrid{};
  const bool coded = src->coded.rep != kRepFloat && src->coded.codes != nullptr;
  const uint64_t cells = src->map.n_cells;
  if (coded)
  {
    const size_t align = (size_t)2 << 20;
    void* alloc = nullptr;
    e = cudaMalloc(&alloc, src->coded.bytes + align);
    if (e == cudaSuccess)
    {
      g->coded = src->coded;
      g->coded.codes_alloc = alloc;
      g->coded.codes = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(alloc) + align - 1) & ~(uintptr_t)(align - 1));
      g->coded.lut = nullptr;
      e = cudaMemcpyPeer(g->coded.codes, device, src->coded.codes, src->device, src->coded.bytes);
    }
    if (e == cudaSuccess)
      e = cudaMalloc(&g->coded.lut, sizeof(float) * std::max<uint32_t>(src->coded.lut_n, 1));
    if (e == cudaSuccess)
      e = cudaMemcpyPeer(g->coded.lut, device, src->coded.lut, src->device, sizeof(float) * src->coded.lut_n);
  }
  if (e == cudaSuccess && (with_float || !coded) && src->d_prob)
  {
    e = ensure(g->d_prob, g->prob_cap, cells);
    if (e == cudaSuccess)
      e = cudaMemcpyPeer(g->d_prob, device, src->d_prob, src->device, sizeof(float) * cells);
  }
  if (e == cudaSuccess)
    e = cudaDeviceSynchronize();
  if (e != cudaSuccess)
  {
    amcl3d_b200_grid_destroy(g);
    return fail_cuda(e, "grid_clone");
  }
  refresh_map(g);
  *out = g;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_download(amcl3d_b200_grid_t g, float* prob, uint64_t cap)
{
  if (!g || !prob)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (!g->map.has_grid)
    return fail(AMCL3D_B200_ERR_STATE, "no grid");
  if (cap < g->map.n_cells)
    return fail(AMCL3D_B200_ERR_ARG, "output too small");
  if (!g->d_prob)
    return fail(AMCL3D_B200_ERR_STATE, "this replica holds the coded cells only (amcl3d_b200_grid_clone with_float = 0)");
  if (int rc = use_device(g->device))
    return rc;
  CU(cudaMemcpy(prob, g->d_prob, sizeof(float) * g->map.n_cells, cudaMemcpyDeviceToHost));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_device_ptr(amcl3d_b200_grid_t g, void** d_prob)
{
  if (!g || !d_prob)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  *d_prob = g->d_prob;
  return AMCL3D_B200_OK;
}

// xyz: HOST points (copied H2D first) or, with on_device, a DEVICE buffer used in place
static int grid_build_impl(amcl3d_b200_grid_t g, const float* xyz, bool on_device, uint32_t stride, uint64_t n, double sensor_dev,
                           int mode, int sub, int keep_edt)
{
  NvtxRange nvtx_range("computeGrid");
  if (!g || !xyz || stride < 3)
    return fail(AMCL3D_B200_ERR_ARG, "bad cloud");
  if (!g->map.has_map)
    return fail(AMCL3D_B200_ERR_STATE, "PointCloudInfo is NULL");  // PointCloudTools.cpp:113-114
  if (mode != AMCL3D_B200_GRID_QUIRK && mode != AMCL3D_B200_GRID_GAUSSIAN && mode != AMCL3D_B200_GRID_SPLAT)
    return fail(AMCL3D_B200_ERR_ARG, "unknown grid mode");
  if (mode != AMCL3D_B200_GRID_SPLAT && sub != 1 && sub != 2)
    return fail(AMCL3D_B200_ERR_ARG, "sub must be 1 or 2");
  if (int rc = use_device(g->device))
    return rc;
  g->map.has_grid = 0;
  g->build_timed = false;
  size_from_bounds(g);
  refresh_map(g);
  const uint64_t cells = g->map.n_cells;
  if (mode == AMCL3D_B200_GRID_SPLAT && cells >= (1ull << 32))
    return fail(AMCL3D_B200_ERR_ARG, "computeGrid2 indexes cells with uint32_t");
  CU(ensure(g->d_prob, g->prob_cap, cells));
  refresh_map(g);
  cudaFree(g->d_edt);
  g->d_edt = nullptr;
  g->edt_cells = 0;
  drop_codes(g);
  refresh_map(g);

  float* d_own = nullptr;
  const float* d_xyz = xyz;
  cudaError_t e = cudaSuccess;
  if (!on_device)
  {
    CU(cudaMalloc(&d_own, sizeof(float) * std::max<uint64_t>(n, 1) * stride));
    e = cudaMemcpyAsync(d_own, xyz, sizeof(float) * n * stride, cudaMemcpyHostToDevice, g->stream);
    d_xyz = d_own;
  }
  char msg[256] = { 0 };
  if (e == cudaSuccess)
  {
    if (mode == AMCL3D_B200_GRID_SPLAT)
    {
      for (int k = 0; k < 4; ++k)
        cudaEventRecord(g->ev[k], g->stream);
      e = build_grid_splat(g->map, d_xyz, stride, n, sensor_dev, g->d_prob, g->stream);
      cudaEventRecord(g->ev[1], g->stream);
    }
    else
      e = build_grid_edt(g->map, d_xyz, stride, n, sensor_dev, mode, sub, g->d_prob, keep_edt ? &g->d_edt : nullptr,
                         g->use_codes ? &g->coded : nullptr, g->edt_fast, g->stream, g->ev, msg, sizeof(msg));
  }
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(g->stream);
  cudaFree(d_own);
  if (e != cudaSuccess)
  {
    if (msg[0])
    {
      (void)cudaGetLastError();
      return fail(e == cudaErrorInvalidValue ? AMCL3D_B200_ERR_ARG : AMCL3D_B200_ERR_CUDA, msg);
    }
    return fail_cuda(e, "grid build");
  }
  g->build_timed = true;
  if (g->d_edt)
    g->edt_cells = cells;
  g->sensor_dev = sensor_dev;
  g->map.has_grid = 1;
  refresh_map(g);
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_build(amcl3d_b200_grid_t g, const float* xyz, uint32_t stride, uint64_t n, double sensor_dev,
                           int mode, int sub, int keep_edt)
{
  return grid_build_impl(g, xyz, false, stride, n, sensor_dev, mode, sub, keep_edt);
}

int amcl3d_b200_grid_build_device(amcl3d_b200_grid_t g, const void* d_xyz, uint32_t stride, uint64_t n, double sensor_dev,
                                  int mode, int sub, int keep_edt)
{
  return grid_build_impl(g, reinterpret_cast<const float*>(d_xyz), true, stride, n, sensor_dev, mode, sub, keep_edt);
}

int amcl3d_b200_grid_last_build_ms(amcl3d_b200_grid_t g, float* ms)
{
  if (!g || !ms)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  *ms = 0.f;
  if (!g->build_timed)
    return AMCL3D_B200_OK;
  float a = 0.f, b = 0.f;
  CU(cudaEventElapsedTime(&a, g->ev[0], g->ev[1]));
  CU(cudaEventElapsedTime(&b, g->ev[2], g->ev[3]));
  *ms = a + b;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_download_edt(amcl3d_b200_grid_t g, int32_t* d2, uint64_t cap)
{
  if (!g || !d2)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (!g->d_edt)
    return fail(AMCL3D_B200_ERR_STATE, "no distance transform kept (keep_edt = 0?)");
  if (cap < g->edt_cells)
    return fail(AMCL3D_B200_ERR_ARG, "output too small");
  if (int rc = use_device(g->device))
    return rc;
  CU(cudaMemcpy(d2, g->d_edt, sizeof(int32_t) * g->edt_cells, cudaMemcpyDeviceToHost));
  return AMCL3D_B200_OK;
}

// Grid3d.cpp:374-402
int amcl3d_b200_grid_save_file(amcl3d_b200_grid_t g, const char* path)
{
  if (!g || !path)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (!g->map.has_grid)
    return fail(AMCL3D_B200_ERR_STATE, "no grid");  // Grid3d.cpp:376-377
  if (g->map.n_cells >= (1ull << 32))
    return fail(AMCL3D_B200_ERR_ARG, "the .grid header stores u32 sizes whose product must fit u32");
  std::vector<float> host(g->map.n_cells);
  int rc = amcl3d_b200_grid_download(g, host.data(), host.size());
  if (rc)
    return rc;
  FILE* pf = fopen(path, "wb");
  if (!pf)
    return fail(AMCL3D_B200_ERR_IO, std::string("Error opening file ") + path + " for writing");
  bool ok = fwrite(&g->map.size[0], sizeof(uint32_t), 1, pf) == 1 && fwrite(&g->map.size[1], sizeof(uint32_t), 1, pf) == 1 &&
            fwrite(&g->map.size[2], sizeof(uint32_t), 1, pf) == 1 && fwrite(&g->sensor_dev, sizeof(double), 1, pf) == 1;
  ok = ok && fwrite(host.data(), sizeof(float), host.size(), pf) == host.size();
  fclose(pf);
  if (!ok)
    return fail(AMCL3D_B200_ERR_IO, "short write");
  return AMCL3D_B200_OK;
}

// Grid3d.cpp:404-442
int amcl3d_b200_grid_load_file(amcl3d_b200_grid_t g, const char* path, double sensor_dev)
{
  if (!g || !path)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  FILE* pf = fopen(path, "rb");
  if (!pf)
    return fail(AMCL3D_B200_ERR_IO, std::string("Error opening file ") + path + " for reading");
  uint32_t size[3];
  double dev = 0;
  if (fread(&size[0], sizeof(uint32_t), 1, pf) != 1 || fread(&size[1], sizeof(uint32_t), 1, pf) != 1 ||
      fread(&size[2], sizeof(uint32_t), 1, pf) != 1 || fread(&dev, sizeof(double), 1, pf) != 1)
  {
    fclose(pf);
    return fail(AMCL3D_B200_ERR_IO, "short header");
  }
  if (fabs(dev - sensor_dev) >= 2.220446049250313e-16)  // Grid3d.cpp:422
  {
    fclose(pf);
    return fail(AMCL3D_B200_ERR_MISMATCH, "Loaded sensorDev is different");
  }
  const uint64_t cells = (uint64_t)size[0] * size[1] * size[2];
  std::vector<float> host;
  try
  {
    host.resize(cells);
  }
  catch (...)
  {
    fclose(pf);
    return fail(AMCL3D_B200_ERR_IO, "grid too large for host staging");
  }
  const size_t got = fread(host.data(), sizeof(float), cells, pf);
  fclose(pf);
  if (got != cells)
    return fail(AMCL3D_B200_ERR_IO, "short read");
  return amcl3d_b200_grid_upload(g, host.data(), size, dev);
}

// Grid3d.cpp:365-372
int amcl3d_b200_grid_is_into_map(amcl3d_b200_grid_t g, float x, float y, float z, int* inside)
{
  if (!g || !inside)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  const MapDev& m = g->map;
  *inside = (m.has_map && x >= m.min[0] && x < m.max[0] && y >= m.min[1] && y < m.max[1] && z >= m.min[2] && z < m.max[2])
                ? 1
                : 0;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_grid_cloud_weight(amcl3d_b200_grid_t g, const float* cloud, uint32_t stride, uint32_t npts, float tx,
                                  float ty, float tz, float roll, float pitch, float yaw, int trig_mode, float* weight)
{
  if (!g || !weight || (npts && !cloud) || stride < 3)
    return fail(AMCL3D_B200_ERR_ARG, "bad argument");
  *weight = 0.f;
  if (!g->map.has_grid || !g->map.has_map)
    return AMCL3D_B200_OK;  // Grid3d.cpp:237-238 returns 0
  if (!g->scratch_pf)
  {
    int rc = amcl3d_b200_pf_create(g, &g->scratch_pf);
    if (rc)
      return rc;
  }
  amcl3d_b200_pf_t pf = g->scratch_pf;
  pf->trig_mode = trig_mode;
  // computeCloudWeight itself has no isIntoMap gate: evaluate the pose even when it lies outside the map
  // by running the kernel on a map copy whose particle gate is opened (bounds for points are unchanged)
  const float p7[7] = { tx, ty, tz, roll, pitch, yaw, 0.f };
  int rc = amcl3d_b200_pf_set_particles(pf, p7, 1);
  if (!rc)
    rc = amcl3d_b200_pf_set_cloud(pf, cloud, stride, npts);
  if (rc)
    return rc;
  if (int rc2 = use_device(g->device))
    return rc2;
  MapDev open = g->map;
  // the per-particle gate compares (double)p >= min && < max; widen it without touching the offsets
  // (offsets use map.min, so pass the true min through and lift only the gate via has_map semantics):
  // the kernel gates on min/max, so use a dedicated flag value 2 = "do not gate"
  open.has_map = 2;
  CU(launch_update(open, pf->d_p, 1, pf->d_cloud, pf->npts, pf->npad, pf->d_beacons, nullptr, trig_mode, nullptr, nullptr,
                   pf->stream));
  pf->launches++;
  float out7[7];
  CU(cudaMemcpyAsync(out7, pf->d_p, sizeof(out7), cudaMemcpyDeviceToHost, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  *weight = out7[6];
  return AMCL3D_B200_OK;
}

/* ========================================= ParticleFilter ======================================= */

int amcl3d_b200_pf_create(amcl3d_b200_grid_t g, amcl3d_b200_pf_t* out)
{
  if (!g || !out)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (int rc = use_device(g->device))
    return rc;
  amcl3d_b200_pf_t pf = new (std::nothrow) amcl3d_b200_pf_s();
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "out of host memory");
  pf->grid = g;
  if (const char* env = getenv("AMCL3D_B200_TWO_PHASE"))
    pf->two_phase = atoi(env) != 0;
  cudaError_t e = cudaStreamCreateWithFlags(&pf->own_stream, cudaStreamNonBlocking);
  pf->stream = pf->own_stream;
  if (e == cudaSuccess)
    e = cudaEventCreate(&pf->ev0);
  if (e == cudaSuccess)
    e = cudaEventCreate(&pf->ev1);
  if (e == cudaSuccess)
    e = cudaEventCreateWithFlags(&pf->ev_sync, cudaEventDisableTiming);
  for (int k = 0; k < 4 && e == cudaSuccess; ++k)
    e = cudaEventCreate(&pf->ev_phase[k]);
  if (e == cudaSuccess)
    e = cudaMalloc(&pf->d_scalars, 64 * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMalloc(&pf->d_u64, 8 * sizeof(unsigned long long));
  if (e == cudaSuccess)
    e = cudaMalloc(&pf->d_int, 8 * sizeof(int));
  if (e == cudaSuccess)
    e = cudaMalloc(&pf->d_delta, 8 * sizeof(double));
  if (e == cudaSuccess)
    e = cudaMalloc(&pf->d_beacons, sizeof(BeaconSet));
  if (e == cudaSuccess)
    e = cudaMemset(pf->d_beacons, 0, sizeof(BeaconSet));
  if (e == cudaSuccess)
    e = cudaMalloc(&pf->d_scratch, reduce_scratch_bytes());
  if (e != cudaSuccess)
  {
    amcl3d_b200_pf_destroy(pf);
    return fail_cuda(e, "pf_create");
  }
  *out = pf;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_destroy(amcl3d_b200_pf_t pf)
{
  if (!pf)
    return AMCL3D_B200_OK;
  cudaSetDevice(pf->grid->device);
  if (pf->shard.on)
  {
    cudaStreamSynchronize(pf->stream);
    for (uint32_t r = 0; r < kMaxShardWorld; ++r)
      if (pf->shard.opened[r])
        cudaIpcCloseMemHandle(pf->shard.opened[r]);
    cudaFree(pf->shard.region);  // p_ lives inside it (pf->external is set)
    cudaFree(pf->shard.d_done);
    cudaFree(pf->shard.d_error);
    cudaFree(pf->shard.d_part7);
  }
  if (!pf->external)
    cudaFree(pf->d_p);
  cudaFree(pf->d_alt);
  cudaFree(pf->d_cloud);
  cudaFree(pf->d_stage);
  cudaFree(pf->d_sorted);
  cudaFree(pf->d_sort_scratch);
  cudaFree(pf->d_codes_staged);
  cudaFree(pf->d_pose);
  cudaFree(pf->d_prefix);
  cudaFree(pf->d_thr);
  cudaFree(pf->d_wr);
  cudaFree(pf->d_dense);
  cudaFree(pf->d_filtered);
  cudaFree(pf->d_vf_scratch);
  cudaFree(pf->d_idx);
  cudaFree(pf->d_noise);
  cudaFree(pf->d_scalars);
  cudaFree(pf->d_u64);
  cudaFree(pf->d_int);
  cudaFree(pf->d_delta);
  cudaFree(pf->d_beacons);
  cudaFree(pf->d_scratch);
  if (pf->ev0)
    cudaEventDestroy(pf->ev0);
  if (pf->ev1)
    cudaEventDestroy(pf->ev1);
  if (pf->ev_sync)
    cudaEventDestroy(pf->ev_sync);
  for (int k = 0; k < 4; ++k)
    if (pf->ev_phase[k])
      cudaEventDestroy(pf->ev_phase[k]);
  if (pf->own_stream)
    cudaStreamDestroy(pf->own_stream);
  delete pf;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_set_trig_mode(amcl3d_b200_pf_t pf, int trig_mode)
{
  if (!pf || (trig_mode != AMCL3D_B200_TRIG_F64 && trig_mode != AMCL3D_B200_TRIG_F32))
    return fail(AMCL3D_B200_ERR_ARG, "bad trig mode");
  pf->trig_mode = trig_mode;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_set_async(amcl3d_b200_pf_t pf, int enable)
{
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  pf->async_mode = enable != 0;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_sync(amcl3d_b200_pf_t pf)
{
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  if (int rc = use_device(pf->grid->device))
    return rc;
  CU(cudaStreamSynchronize(pf->stream));
  if (pf->shard.on)
  {  // a sharded exchange that gave up on a peer (shard.cu wait_kernel) is reported here in stream-order mode
    uint32_t err = 0;
    CU(cudaMemcpy(&err, pf->shard.d_error, sizeof(err), cudaMemcpyDeviceToHost));
    if (err)
      return fail(AMCL3D_B200_ERR_STATE, "sharded exchange timed out waiting for rank " + std::to_string(err - 1));
  }
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_set_two_phase(amcl3d_b200_pf_t pf, int enable)
{
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  pf->two_phase = enable != 0;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_set_stream(amcl3d_b200_pf_t pf, void* cuda_stream)
{
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  pf->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : pf->own_stream;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_set_particles(amcl3d_b200_pf_t pf, const float* aos7, uint64_t n)
{
  if (!pf || (n && !aos7))
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (n >= (1ull << 31))
    return fail(AMCL3D_B200_ERR_ARG, "too many particles");
  if (int rc = use_device(pf->grid->device))
    return rc;
  if (pf->external)
  {
    if (n > pf->cap)
      return fail(AMCL3D_B200_ERR_ARG, "bound device buffer too small");
  }
  else
    CU(ensure_particles(pf->d_p, pf->cap, n));
  if (n)
    CU(cudaMemcpyAsync(pf->d_p, aos7, sizeof(float) * 7 * n, cudaMemcpyHostToDevice, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  pf->n = n;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_get_particles(amcl3d_b200_pf_t pf, float* aos7, uint64_t cap, uint64_t* n)
{
  if (!pf || (!aos7 && cap))
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (n)
    *n = pf->n;
  if (cap < pf->n)
    return fail(AMCL3D_B200_ERR_ARG, "output too small");
  if (int rc = use_device(pf->grid->device))
    return rc;
  if (pf->n)
    CU(cudaMemcpyAsync(aos7, pf->d_p, sizeof(float) * 7 * pf->n, cudaMemcpyDeviceToHost, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_size(amcl3d_b200_pf_t pf, uint64_t* n)
{
  if (!pf || !n)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  *n = pf->n;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_bind_device(amcl3d_b200_pf_t pf, void* d_aos7, uint64_t n, uint64_t capacity)
{
  if (!pf || !d_aos7 || n > capacity || capacity >= (1ull << 31))
    return fail(AMCL3D_B200_ERR_ARG, "bad device buffer");
  if (pf->shard.on)
    return fail(AMCL3D_B200_ERR_STATE, "a sharded handle keeps p_ inside its exchange region");
  if (!pf->external)
    cudaFree(pf->d_p);
  pf->d_p = reinterpret_cast<float*>(d_aos7);
  pf->n = n;
  pf->cap = capacity;
  pf->external = true;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_device_ptr(amcl3d_b200_pf_t pf, void** d_aos7, uint64_t* n)
{
  if (!pf || !d_aos7)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  *d_aos7 = pf->d_p;
  if (n)
    *n = pf->n;
  return AMCL3D_B200_OK;
}

// Morton order of the cloud in d_cloud for the two-phase sweep (keys + radix sort + pack: a few microseconds).
// Runs at set_cloud when the two-phase path is on, else lazily in update() the first time it is needed.
static int sort_cloud(amcl3d_b200_pf_t pf)
{
  CU(ensure(pf->d_sorted, pf->sorted_cap, pf->npad));
  const size_t need = sort_cloud_scratch_bytes(pf->npts);
  CU(ensure(pf->d_sort_scratch, pf->sort_scratch_cap, need));
  CU(launch_sort_cloud(reinterpret_cast<const float*>(pf->d_cloud), 4, pf->npts, pf->npad, pf->d_sorted, pf->d_sort_scratch,
                       need, pf->stream));
  pf->launches += 3;
  pf->launches += chain_extra_launches_take();
  pf->sorted_valid = true;
  return AMCL3D_B200_OK;
}

static int set_cloud_common(amcl3d_b200_pf_t pf, const float* d_src, uint32_t stride, uint32_t npts)
{
  const uint32_t npad = (npts + 31u) & ~31u;
  CU(ensure(pf->d_cloud, pf->cloud_cap, npad));
  CU(launch_pack_cloud(d_src, stride, npts, pf->d_cloud, npad, pf->stream));
  pf->launches += npad ? 1 : 0;
  pf->npts = npts;
  pf->npad = npad;
  pf->filtered_valid = false;
  pf->sorted_valid = false;  // d_sorted (if any) belongs to the previous cloud
  if (pf->two_phase && npts > 0)
    return sort_cloud(pf);
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_set_cloud(amcl3d_b200_pf_t pf, const float* cloud, uint32_t stride, uint32_t npts)
{
  if (!pf || (npts && !cloud) || stride < 3)
    return fail(AMCL3D_B200_ERR_ARG, "bad cloud");
  if (int rc = use_device(pf->grid->device))
    return rc;
  const uint64_t floats = (uint64_t)npts * stride;
  CU(ensure(pf->d_stage, pf->stage_cap, floats));
  if (floats)
    CU(cudaMemcpyAsync(pf->d_stage, cloud, sizeof(float) * floats, cudaMemcpyHostToDevice, pf->stream));
  int rc = set_cloud_common(pf, pf->d_stage, stride, npts);
  if (rc)
    return rc;
  CU(cudaStreamSynchronize(pf->stream));  // the host buffer may be reused by the caller
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_set_cloud_device(amcl3d_b200_pf_t pf, const void* d_cloud, uint32_t stride, uint32_t npts)
{
  if (!pf || (npts && !d_cloud) || stride < 3)
    return fail(AMCL3D_B200_ERR_ARG, "bad cloud");
  if (int rc = use_device(pf->grid->device))
    return rc;
  return set_cloud_common(pf, reinterpret_cast<const float*>(d_cloud), stride, npts);
}

int amcl3d_b200_pf_set_cloud_downsampled(amcl3d_b200_pf_t pf, const float* cloud, uint32_t stride, int intensity_offset,
                                         uint32_t npts, float leaf, int skip_nonfinite, uint32_t* npts_out)
{
  if (!pf || (npts && !cloud) || stride < 3 || intensity_offset >= (int)stride || !(leaf > 0.f))
    return fail(AMCL3D_B200_ERR_ARG, "bad cloud / leaf size");
  if (int rc = use_device(pf->grid->device))
    return rc;
  const uint64_t floats = (uint64_t)npts * stride;
  CU(ensure(pf->d_stage, pf->stage_cap, floats));
  CU(ensure(pf->d_filtered, pf->filtered_cap, (uint64_t)(npts ? npts : 1)));
  const size_t need = voxel_filter_scratch_bytes(npts);
  CU(ensure(pf->d_vf_scratch, pf->vf_scratch_cap, need));
  if (floats)
    CU(cudaMemcpyAsync(pf->d_stage, cloud, sizeof(float) * floats, cudaMemcpyHostToDevice, pf->stream));
  int64_t cells = 0;
  uint32_t launched = 0;
  {
    Timed t(pf);
    CU(run_voxel_filter(pf->d_stage, stride, intensity_offset, npts, leaf, skip_nonfinite, pf->d_filtered,
                        pf->d_vf_scratch, need, &cells, &launched, pf->stream));
    pf->launches += launched;
  }
  int rc;
  if (cells < 0)  // PCL: "Leaf size is too small for the input dataset" -> the input passes through
    rc = set_cloud_common(pf, pf->d_stage, stride, npts);
  else
  {
    rc = set_cloud_common(pf, reinterpret_cast<const float*>(pf->d_filtered), 4, (uint32_t)cells);
    pf->filtered_valid = (rc == AMCL3D_B200_OK);
  }
  if (rc)
    return rc;
  CU(cudaStreamSynchronize(pf->stream));
  if (npts_out)
    *npts_out = pf->npts;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_get_cloud(amcl3d_b200_pf_t pf, float* xyzi, uint32_t cap, uint32_t* npts)
{
  if (!pf || (!xyzi && cap))
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (npts)
    *npts = pf->npts;
  if (!xyzi)
    return AMCL3D_B200_OK;
  if (cap < pf->npts)
    return fail(AMCL3D_B200_ERR_ARG, "output too small");
  if (int rc = use_device(pf->grid->device))
    return rc;
  if (pf->npts)
  {
    const float4* src = pf->filtered_valid ? pf->d_filtered : pf->d_cloud;  // d_cloud: x y z 0
    CU(cudaMemcpyAsync(xyzi, src, sizeof(float4) * pf->npts, cudaMemcpyDeviceToHost, pf->stream));
    CU(cudaStreamSynchronize(pf->stream));
  }
  return AMCL3D_B200_OK;
}

static int fuse_range_async(amcl3d_b200_pf_t pf, const float* d_wr, float alpha)
{
  // w = alpha*wp/sum(wp) + (1-alpha)*wr/sum(wr): both sums as sequential f32 chains
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(pf->n)));
  CU(launch_chain_prefix(pf->d_p, pf->n, 0, nullptr, pf->d_scalars + 0, pf->d_dense, pf->chain_mode, pf->stream));
  CU(launch_chain_sum_array(d_wr, pf->n, pf->d_scalars + 1, pf->d_dense, pf->chain_mode, pf->stream));
  CU(launch_fuse_range(pf->d_p, d_wr, pf->n, pf->d_scalars + 0, pf->d_scalars + 1, alpha, pf->stream));
  pf->launches += 4;
  pf->launches += chain_extra_launches_take();
  return AMCL3D_B200_OK;
}

// d_wr_ext != nullptr: raw range weights go there and the fusion is left to the caller (multi-GPU: the sums
// of the fusion are global, so it runs after the all-gather)
static int update_impl(amcl3d_b200_pf_t pf, const amcl3d_b200_beacon_t* beacons, uint32_t nb, float alpha, float sigma,
                       uint64_t* in_bounds, float* d_wr_ext)
{
  NvtxRange nvtx_range("Update");
  if (!pf || (nb && !beacons) || nb > (uint32_t)kMaxBeacons)
    return fail(AMCL3D_B200_ERR_ARG, "bad beacons");
  if (int rc = use_device(pf->grid->device))
    return rc;
  const MapDev& map = pf->grid->map;
  float* d_wr = nullptr;
  if (nb)
  {
    BeaconSet bs{};
    bs.n = nb;
    bs.k1 = 1.f / (sigma * sqrt(2 * M_PI));
    bs.k2 = 0.5f / (sigma * sigma);
    for (uint32_t b = 0; b < nb; ++b)
    {
      bs.ax[b] = (float)beacons[b].position[0];
      bs.ay[b] = (float)beacons[b].position[1];
      bs.az[b] = (float)beacons[b].position[2];
      bs.r[b] = (float)beacons[b].range;
    }
    CU(cudaMemcpyAsync(pf->d_beacons, &bs, sizeof(bs), cudaMemcpyHostToDevice, pf->stream));
    CU(cudaStreamSynchronize(pf->stream));  // bs lives on this stack frame
    if (d_wr_ext != nullptr)
      d_wr = d_wr_ext;
    else
    {
      CU(ensure(pf->d_wr, pf->wr_cap, pf->n));
      d_wr = pf->d_wr;
    }
  }
  unsigned long long* d_counter = nullptr;
  if (in_bounds)
  {
    d_counter = pf->d_u64;
    CU(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), pf->stream));
  }
  // two-phase sweep when it pays: enough particles to fill the machine, a u8-coded grid, staging that fits
  // The staging buffer holds one byte per particle x point; a particle set too large for the budget is swept in
  // slices of whole particle tiles, one after the other through the same buffer (particles are independent).
  TwoPhase tp{};
  const TwoPhase* tpp = nullptr;
  uint64_t slice_n = pf->n;
  // The sweep parallelises over (particle tile, cloud chunk) pairs, so it also wins for SMALL sets: the single-kernel path
  // gives one thread the whole cloud of its particle, i.e. 600 particles occupy 3 blocks of a 148-SM machine and every
  // thread walks 2 000 dependent-latency gathers (measured at BASELINE config 1, 600 x 2 000: 0.337 -> 0.070 ms; a single
  // pose of Grid3d::computeCloudWeight likewise).  The single kernel keeps the clouds of less than two chunks.
  if (pf->two_phase && pf->n >= 1 && pf->npts >= 256 && two_phase_supported(map))
  {
    if (!pf->sorted_valid)  // the cloud was set while the two-phase path was off: sort it now
      if (int rc = sort_cloud(pf))
        return rc;
    const uint64_t pitch = stage_tile_width();  // bytes per staged row of one particle tile (= particles per block)
    const uint64_t tile_bytes = pitch * pf->npad;
    const uint64_t tiles = (pf->n + pitch - 1) / pitch;
    uint64_t budget = 4ull << 30;
    if (const char* env = getenv("AMCL3D_B200_STAGE_MB"))
      budget = (uint64_t)atoll(env) << 20;
    uint64_t slice_tiles = budget / tile_bytes;
    if (slice_tiles > tiles)
      slice_tiles = tiles;
    if (slice_tiles >= 16 || slice_tiles == tiles)  // a set cut into slices of less than 16 tiles does not fill the machine
    {
      CU(ensure(pf->d_codes_staged, pf->codes_staged_cap, slice_tiles * tile_bytes));
      CU(ensure(pf->d_pose, pf->pose_cap, (uint64_t)pose_floats((uint32_t)(slice_tiles * pitch))));
      tp.sorted = pf->d_sorted;
      tp.stage = pf->d_codes_staged;
      tp.pose = pf->d_pose;
      tpp = &tp;
      slice_n = slice_tiles * pitch;
    }
  }
  {
    Timed t(pf);
    const uint64_t pitch_w = stage_tile_width();
    for (uint64_t off = 0; off < pf->n || off == 0; off += slice_n)
    {
      const uint64_t n_s = pf->n - off < slice_n ? pf->n - off : slice_n;
      if (tpp && n_s)
      {
        // phase 0: rotation + offsets once per particle of this slice
        cudaEventRecord(pf->ev_phase[0], pf->stream);
        CU(launch_pose(map, pf->d_p + 7 * off, (uint32_t)n_s, pf->trig_mode, pf->d_pose, pf->stream));
        cudaEventRecord(pf->ev_phase[1], pf->stream);
        tp.pose_pitch = (uint32_t)((n_s + pitch_w - 1) / pitch_w * pitch_w);
        tp.ev_mid = pf->ev_phase[2];
        tp.ev_end = pf->ev_phase[3];
        pf->phase_timed = true;
      }
      else
        pf->phase_timed = false;
      CU(launch_update(map, pf->d_p + 7 * off, (uint32_t)n_s, pf->d_cloud, pf->npts, pf->npad, pf->d_beacons,
                       d_wr ? d_wr + off : nullptr, pf->trig_mode, d_counter, tpp, pf->stream));
      pf->launches += n_s ? (tpp ? 3 : 1) : 0;
      if (pf->n == 0)
        break;
    }
    if (nb && pf->n && d_wr_ext == nullptr)
      if (int rc = fuse_range_async(pf, pf->d_wr, alpha))
        return rc;
  }
  if (in_bounds)
  {
    unsigned long long c = 0;
    CU(cudaMemcpyAsync(&c, d_counter, sizeof(c), cudaMemcpyDeviceToHost, pf->stream));
    CU(cudaStreamSynchronize(pf->stream));
    *in_bounds = c;
    return AMCL3D_B200_OK;
  }
  return finish_call(pf);
}

int amcl3d_b200_pf_update(amcl3d_b200_pf_t pf, const amcl3d_b200_beacon_t* beacons, uint32_t nb, float alpha, float sigma,
                          uint64_t* in_bounds)
{
  return update_impl(pf, beacons, nb, alpha, sigma, in_bounds, nullptr);
}

int amcl3d_b200_pf_update_split(amcl3d_b200_pf_t pf, const amcl3d_b200_beacon_t* beacons, uint32_t nb, float sigma,
                                void* d_wr_out, uint64_t* in_bounds)
{
  if (nb && !d_wr_out)
    return fail(AMCL3D_B200_ERR_ARG, "d_wr_out is NULL");
  return update_impl(pf, beacons, nb, 1.f, sigma, in_bounds, reinterpret_cast<float*>(d_wr_out));
}

int amcl3d_b200_pf_fuse_range(amcl3d_b200_pf_t pf, const void* d_wr, float alpha)
{
  if (!pf || !d_wr)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (int rc = use_device(pf->grid->device))
    return rc;
  if (pf->n)
  {
    Timed t(pf);
    if (int rc = fuse_range_async(pf, reinterpret_cast<const float*>(d_wr), alpha))
      return rc;
  }
  return finish_call(pf);
}

int amcl3d_b200_pf_update_cloud(amcl3d_b200_pf_t pf, const float* cloud, uint32_t stride, uint32_t npts)
{
  int rc = amcl3d_b200_pf_set_cloud(pf, cloud, stride, npts);
  if (rc)
    return rc;
  return amcl3d_b200_pf_update(pf, nullptr, 0, 1.f, 1.f, nullptr);
}

static int normalize_async(amcl3d_b200_pf_t pf)
{
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(pf->n)));
  CU(launch_chain_prefix(pf->d_p, pf->n, 1, nullptr, pf->d_scalars + 0, pf->d_dense, pf->chain_mode, pf->stream));
  CU(launch_normalize(pf->grid->map, pf->d_p, pf->n, pf->d_scalars + 0, pf->stream));
  pf->launches += pf->n ? 2 : 1;
  pf->launches += chain_extra_launches_take();
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_normalize(amcl3d_b200_pf_t pf, float* wtp)
{
  NvtxRange nvtx_range("particleNormalize");
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  if (int rc = use_device(pf->grid->device))
    return rc;
  {
    Timed t(pf);
    if (int rc = normalize_async(pf))
      return rc;
  }
  if (!wtp)
    return finish_call(pf);
  float total = 0;
  CU(cudaMemcpyAsync(&total, pf->d_scalars, sizeof(float), cudaMemcpyDeviceToHost, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  *wtp = total;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_scale_by_max(amcl3d_b200_pf_t pf)
{
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  if (int rc = use_device(pf->grid->device))
    return rc;
  {
    Timed t(pf);
    CU(launch_max_weight(pf->d_p, pf->n, pf->d_scalars + 2, pf->d_u64 + 1, pf->d_scratch, pf->stream));
    CU(launch_scale_by_max(pf->d_p, pf->n, pf->d_scalars + 2, pf->stream));
    pf->launches += 3;
    pf->launches += chain_extra_launches_take();
  }
  return finish_call(pf);
}

// replaces p_ by the first `count` particles of d_alt
static int adopt_alt(amcl3d_b200_pf_t pf, uint64_t count)
{
  if (pf->external)
  {
    if (count > pf->cap)
      return fail(AMCL3D_B200_ERR_ARG, "bound device buffer too small for the resampled set");
    if (count)
      CU(cudaMemcpyAsync(pf->d_p, pf->d_alt, sizeof(float) * 7 * count, cudaMemcpyDeviceToDevice, pf->stream));
  }
  else
  {
    std::swap(pf->d_p, pf->d_alt);
    std::swap(pf->cap, pf->alt_cap);
  }
  pf->n = count;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_resample(amcl3d_b200_pf_t pf, float u01, uint64_t m0, uint64_t m1, void* d_out, int32_t* idx)
{
  NvtxRange nvtx_range("Resample");
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  const uint64_t n = pf->n;
  if (m1 == 0)
    m1 = n;
  if (m0 > m1 || m1 > n)
    return fail(AMCL3D_B200_ERR_ARG, "bad output range");
  if (!d_out && (m0 != 0 || m1 != n))
    return fail(AMCL3D_B200_ERR_ARG, "replacing p_ needs the full output range");
  if (n == 0)
    return AMCL3D_B200_OK;
  if (int rc = use_device(pf->grid->device))
    return rc;
  const uint64_t cnt = m1 - m0;
  CU(ensure(pf->d_prefix, pf->prefix_cap, n));
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(n)));
  float* out = reinterpret_cast<float*>(d_out);
  if (!out)
  {
    CU(ensure_particles(pf->d_alt, pf->alt_cap, n));
    out = pf->d_alt;
  }
  int32_t* d_idx = nullptr;
  if (idx)
  {
    CU(ensure(pf->d_idx, pf->idx_cap, cnt));
    d_idx = pf->d_idx;
  }
  {
    Timed t(pf);
    CU(launch_chain_prefix(pf->d_p, n, 0, pf->d_prefix, nullptr, pf->d_dense, pf->chain_mode, pf->stream));
    CU(launch_resample_search(pf->d_p, pf->d_prefix, n, u01, m0, m1, out, d_idx, pf->stream));
    pf->launches += 2;
    pf->launches += chain_extra_launches_take();
  }
  if (!d_out)
    if (int rc = adopt_alt(pf, n))
      return rc;
  if (idx && cnt)
  {
    CU(cudaMemcpyAsync(idx, d_idx, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, pf->stream));
    CU(cudaStreamSynchronize(pf->stream));
    return AMCL3D_B200_OK;
  }
  return finish_call(pf);
}

int amcl3d_b200_pf_uniform_sample(amcl3d_b200_pf_t pf, int sample_num, uint64_t k0, uint64_t k1, void* d_out,
                                  int32_t* idx, int* emitted)
{
  NvtxRange nvtx_range("uniformSample");
  if (!pf || sample_num < 0)
    return fail(AMCL3D_B200_ERR_ARG, "bad sample_num");
  const uint64_t S = (uint64_t)sample_num;
  const uint64_t n = pf->n;
  if (k1 == 0)
    k1 = S;
  if (k0 > k1 || k1 > S)
    return fail(AMCL3D_B200_ERR_ARG, "bad output range");
  if (!d_out && (k0 != 0 || k1 != S))
    return fail(AMCL3D_B200_ERR_ARG, "replacing p_ needs the full output range");
  if (int rc = use_device(pf->grid->device))
    return rc;
  const uint64_t cnt = k1 - k0;
  CU(ensure(pf->d_prefix, pf->prefix_cap, n));
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(n > (uint64_t)S ? n : (uint64_t)S)));
  CU(ensure(pf->d_thr, pf->thr_cap, S));
  float* out = reinterpret_cast<float*>(d_out);
  if (!out)
  {
    CU(ensure_particles(pf->d_alt, pf->alt_cap, S));
    out = pf->d_alt;
  }
  int32_t* d_idx = nullptr;
  if (idx)
  {
    CU(ensure(pf->d_idx, pf->idx_cap, cnt));
    d_idx = pf->d_idx;
  }
  {
    Timed t(pf);
    if (int rc = normalize_async(pf))  // ParticleFilter.cpp:258
      return rc;
    CU(launch_chain_prefix(pf->d_p, n, 0, pf->d_prefix, nullptr, pf->d_dense, pf->chain_mode, pf->stream));
    CU(launch_chain_thresholds(sample_num, pf->d_thr, pf->d_dense, pf->chain_mode, pf->stream));
    CU(launch_uniform_search(pf->d_p, pf->d_prefix, n, pf->d_thr, sample_num, k0, k1, out, d_idx, pf->stream));
    CU(launch_count_emitted(pf->d_thr, sample_num, pf->d_prefix, n, pf->d_int, pf->stream));
    pf->launches += 4;
    pf->launches += chain_extra_launches_take();
  }
  if (!d_out)
    if (int rc = adopt_alt(pf, S))
      return rc;
  int em = 0;
  CU(cudaMemcpyAsync(&em, pf->d_int, sizeof(int), cudaMemcpyDeviceToHost, pf->stream));
  if (idx && cnt)
    CU(cudaMemcpyAsync(idx, d_idx, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  if (emitted)
    *emitted = em;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_mean(amcl3d_b200_pf_t pf, int exact, float out7[7])
{
  NvtxRange nvtx_range("getMean");
  if (!pf || !out7)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (int rc = use_device(pf->grid->device))
    return rc;
  {
    Timed t(pf);
    CU(launch_mean(pf->d_p, pf->n, exact, pf->d_scalars + 8, pf->d_scratch, pf->stream));
    pf->launches += exact ? 1 : 2;
  }
  CU(cudaMemcpyAsync(out7, pf->d_scalars + 8, 7 * sizeof(float), cudaMemcpyDeviceToHost, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_mean_device(amcl3d_b200_pf_t pf, int exact, void* d_out7)
{
  if (!pf || !d_out7)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (int rc = use_device(pf->grid->device))
    return rc;
  {
    Timed t(pf);
    CU(launch_mean(pf->d_p, pf->n, exact, reinterpret_cast<float*>(d_out7), pf->d_scratch, pf->stream));
    pf->launches += exact ? 1 : 2;
  }
  return finish_call(pf);
}

int amcl3d_b200_pf_best(amcl3d_b200_pf_t pf, float out7[7])
{
  if (!pf || !out7)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (int rc = use_device(pf->grid->device))
    return rc;
  {
    Timed t(pf);
    CU(launch_max_weight(pf->d_p, pf->n, pf->d_scalars + 2, pf->d_u64 + 1, pf->d_scratch, pf->stream));
    pf->launches += 2;
    pf->launches += chain_extra_launches_take();
  }
  unsigned long long arg = ~0ull;
  CU(cudaMemcpyAsync(&arg, pf->d_u64 + 1, sizeof(arg), cudaMemcpyDeviceToHost, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  for (int k = 0; k < 7; ++k)
    out7[k] = 0.f;  // best_p starts as the zero particle (ParticleFilter.cpp:220-221)
  if (arg != ~0ull && arg < pf->n)
    CU(cudaMemcpy(out7, pf->d_p + 7 * arg, 7 * sizeof(float), cudaMemcpyDeviceToHost));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_predict(amcl3d_b200_pf_t pf, const double delta[6], const float* noise6)
{
  NvtxRange nvtx_range("PFMove");
  if (!pf || !delta || (pf->n && !noise6))
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (int rc = use_device(pf->grid->device))
    return rc;
  CU(ensure(pf->d_noise, pf->noise_cap, pf->n * 6));
  CU(cudaMemcpyAsync(pf->d_delta, delta, 6 * sizeof(double), cudaMemcpyHostToDevice, pf->stream));
  if (pf->n)
    CU(cudaMemcpyAsync(pf->d_noise, noise6, sizeof(float) * 6 * pf->n, cudaMemcpyHostToDevice, pf->stream));
  {
    Timed t(pf);
    CU(launch_predict(pf->d_p, pf->n, pf->d_delta, pf->d_noise, pf->stream));
    pf->launches += pf->n ? 1 : 0;
  }
  CU(cudaStreamSynchronize(pf->stream));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_prefix(amcl3d_b200_pf_t pf, void** d_prefix, float* host_out, uint64_t cap)
{
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  if (host_out && cap < pf->n)
    return fail(AMCL3D_B200_ERR_ARG, "output too small");
  if (int rc = use_device(pf->grid->device))
    return rc;
  CU(ensure(pf->d_prefix, pf->prefix_cap, pf->n));
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(pf->n)));
  {
    Timed t(pf);
    CU(launch_chain_prefix(pf->d_p, pf->n, 0, pf->d_prefix, nullptr, pf->d_dense, pf->chain_mode, pf->stream));
    pf->launches += 1;
    pf->launches += chain_extra_launches_take();
  }
  if (host_out && pf->n)
    CU(cudaMemcpyAsync(host_out, pf->d_prefix, sizeof(float) * pf->n, cudaMemcpyDeviceToHost, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  if (d_prefix)
    *d_prefix = pf->d_prefix;
  return AMCL3D_B200_OK;
}

/* ---- callers of the path: MonteCarloLocalization (amcl3d/src/amcl3d.cpp) ------------------------------------ */

// amcl3d.cpp:242-288
int amcl3d_b200_pf_compute_sample_num(amcl3d_b200_pf_t pf, int cnt_per_grid, int min_resamp, int max_resamp, int* sample_num)
{
  if (!pf || !sample_num)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  if (int rc = use_device(pf->grid->device))
    return rc;
  // computeBoundary (amcl3d.cpp:222-241): all zeros for an empty set, else min / max over the set
  float lo[3] = { 0.f, 0.f, 0.f }, hi[3] = { 0.f, 0.f, 0.f };
  uint32_t* d_keys = reinterpret_cast<uint32_t*>(pf->d_u64 + 2);  // 6 x u32 inside the u64 scratch
  if (pf->n)
  {
    Timed t(pf);
    CU(launch_particle_bounds(pf->d_p, pf->n, d_keys, pf->stream));
    pf->launches += 1;
    pf->launches += chain_extra_launches_take();
  }
  if (pf->n)
  {
    uint32_t keys[6];
    CU(cudaMemcpyAsync(keys, d_keys, sizeof(keys), cudaMemcpyDeviceToHost, pf->stream));
    CU(cudaStreamSynchronize(pf->stream));
    for (int a = 0; a < 6; ++a)
    {
      const uint32_t k = keys[a];
      const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
      float f;
      memcpy(&f, &b, 4);
      (a < 3 ? lo[a] : hi[a - 3]) = f;
    }
  }
  float xyz_step = 0.2;
  float theta_step = 8.0;
  int xwidth = std::ceil((hi[0] - lo[0]) / xyz_step) + 1;
  int ywidth = std::ceil((hi[1] - lo[1]) / xyz_step) + 1;
  int zwidth = std::ceil((hi[2] - lo[2]) / xyz_step) + 1;
  int theta = std::ceil((M_PI * 2 / theta_step * 180)) + 1;
  if (xwidth * ywidth * zwidth == 0)
  {
    *sample_num = min_resamp;
    return AMCL3D_B200_OK;
  }
  else if (xwidth * ywidth * zwidth > 1e3)
  {
    *sample_num = max_resamp;
    return AMCL3D_B200_OK;
  }
  const size_t cells = (size_t)xwidth * ywidth * zwidth * theta;
  const size_t words = (cells + 31) / 32;
  uint32_t* d_bits = nullptr;
  CU(cudaMalloc(&d_bits, words * sizeof(uint32_t)));
  int cnt = 0;
  cudaError_t e = launch_sample_bins(pf->d_p, pf->n, lo[0], lo[1], lo[2], xwidth, ywidth, zwidth, theta, d_bits, words, pf->d_int + 1,
                                     pf->stream);
  pf->launches += 1;
  pf->launches += chain_extra_launches_take();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(&cnt, pf->d_int + 1, sizeof(int), cudaMemcpyDeviceToHost, pf->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(pf->stream);
  cudaFree(d_bits);
  if (e != cudaSuccess)
    return fail_cuda(e, "computeSampleNum");
  int s = cnt * cnt_per_grid;
  if (s < min_resamp)
    s = min_resamp;
  if (s > max_resamp)
    s = max_resamp;
  *sample_num = s;
  return AMCL3D_B200_OK;
}

// amcl3d.cpp:295-321: mean (sequential f32, like getMean), nearest keypose to (mean.x, mean.y, 0), z += its z - mean.z
int amcl3d_b200_pf_update_sample_height(amcl3d_b200_pf_t pf, const float* keyposes_xyz, uint64_t nkey, float* delta_h)
{
  if (!pf || !keyposes_xyz || nkey == 0)
    return fail(AMCL3D_B200_ERR_ARG, "no keyposes");
  float mean7[7];
  int rc = amcl3d_b200_pf_mean(pf, /*exact=*/1, mean7);
  if (rc)
    return rc;
  // the kd-tree query of the reference is an exact 1-NN (FLANN L2_Simple<float>, f32 accumulation in x,y,z order)
  const float q[3] = { mean7[0], mean7[1], 0.f };
  float best = INFINITY;
  uint64_t bi = 0;
  for (uint64_t k = 0; k < nkey; ++k)
  {
    float result = 0.f, diff;
    diff = q[0] - keyposes_xyz[3 * k];
    result += diff * diff;
    diff = q[1] - keyposes_xyz[3 * k + 1];
    result += diff * diff;
    diff = q[2] - keyposes_xyz[3 * k + 2];
    result += diff * diff;
    if (result < best)
    {
      best = result;
      bi = k;
    }
  }
  const float dh = keyposes_xyz[3 * bi + 2] - mean7[2];
  {
    Timed t(pf);
    CU(launch_shift_z(pf->d_p, pf->n, dh, pf->stream));
    pf->launches += pf->n ? 1 : 0;
  }
  CU(cudaStreamSynchronize(pf->stream));
  if (delta_h)
    *delta_h = dh;
  return AMCL3D_B200_OK;
}

// amcl3d.cpp:192-203 (computeSampleNum is called with its default cnt_per_grid = 3, amcl3d.h)
int amcl3d_b200_pf_resample_step(amcl3d_b200_pf_t pf, int weight_multiply, int min_resamp, int max_resamp,
                                 const float* keyposes_xyz, uint64_t nkey, int* new_size)
{
  NvtxRange nvtx_range("Resample");
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  int rc = AMCL3D_B200_OK;
  if (weight_multiply)
    rc = amcl3d_b200_pf_scale_by_max(pf);
  int sample_num = 0;
  if (!rc)
    rc = amcl3d_b200_pf_compute_sample_num(pf, 3, min_resamp, max_resamp, &sample_num);
  if (!rc)
    rc = amcl3d_b200_pf_uniform_sample(pf, sample_num, 0, 0, nullptr, nullptr, nullptr);
  if (!rc && keyposes_xyz && nkey)
    rc = amcl3d_b200_pf_update_sample_height(pf, keyposes_xyz, nkey, nullptr);
  if (!rc && new_size)
    *new_size = sample_num;
  return rc;
}


/* ================================ sharded particle set over peer memory ================================ */
namespace
{
uint64_t align256(uint64_t v)
{
  return (v + 255) & ~(uint64_t)255;
}
// contiguous block partition; the first (n_total % world) ranks hold one particle more
void shard_bounds(uint64_t n_total, uint32_t world, uint32_t rank, uint64_t* lo, uint64_t* hi)
{
  const uint64_t base = n_total / world, rem = n_total % world;
  *lo = rank * base + std::min<uint64_t>(rank, rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}
int shard_check_error(amcl3d_b200_pf_t pf)
{
  uint32_t err = 0;
  CU(cudaMemcpyAsync(&err, pf->shard.d_error, sizeof(err), cudaMemcpyDeviceToHost, pf->stream));
  CU(cudaStreamSynchronize(pf->stream));
  if (err)
    return fail(AMCL3D_B200_ERR_STATE, "sharded exchange timed out waiting for rank " + std::to_string(err - 1));
  return AMCL3D_B200_OK;
}
}  // namespace

int amcl3d_b200_pf_shard_init(amcl3d_b200_pf_t pf, uint64_t n_total, int rank, int world)
{
  return amcl3d_b200_pf_shard_init_capacity(pf, n_total, n_total, rank, world);
}

int amcl3d_b200_pf_shard_init_capacity(amcl3d_b200_pf_t pf, uint64_t n_total, uint64_t n_capacity, int rank, int world)
{
  if (!pf || world < 1 || world > (int)kMaxShardWorld || rank < 0 || rank >= world || n_total == 0 || n_capacity < n_total ||
      n_capacity >= (1ull << 31))
    return fail(AMCL3D_B200_ERR_ARG, "bad shard geometry");
  if (pf->shard.on)
    return fail(AMCL3D_B200_ERR_STATE, "already sharded");
  if (int rc = use_device(pf->grid->device))
    return rc;
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  S.rank = (uint32_t)rank;
  S.world = (uint32_t)world;
  S.n_total = n_total;
  S.n_cap = n_capacity;
  S.n_max = (n_capacity + world - 1) / world;
  shard_bounds(n_total, S.world, S.rank, &S.lo, &S.hi);
  ShardPeers& P = S.peers;
  P.world = S.world;
  uint64_t off = 0;
  P.off_p = off;
  P.p_bytes = align256(S.n_max * 28);
  off += 2 * P.p_bytes;
  P.off_w = off;
  off += align256(2 * n_capacity * 4);
  P.off_wr = off;
  off += align256(2 * n_capacity * 4);
  P.off_inmap = off;
  off += align256(2 * n_capacity);
  P.off_mean = off;
  off += align256(2 * kMaxShardWorld * 8 * 4);
  P.off_flags = off;
  off += align256(kShardFlagKinds * kMaxShardWorld * 4);
  P.off_sums = off;
  off += align256(2 * 2 * kMaxShardWorld * 4);
  P.off_idx = off;
  off += align256(S.n_max * 4);
  S.region_bytes = off;
  CU(cudaMalloc(&S.region, S.region_bytes));
  CU(cudaMemset(S.region, 0, S.region_bytes));
  CU(cudaMalloc(&S.d_done, sizeof(unsigned int)));
  CU(cudaMemset(S.d_done, 0, sizeof(unsigned int)));
  CU(cudaMalloc(&S.d_error, sizeof(uint32_t)));
  CU(cudaMemset(S.d_error, 0, sizeof(uint32_t)));
  CU(cudaMalloc(&S.d_part7, 8 * sizeof(float)));
  for (uint32_t r = 0; r < kMaxShardWorld; ++r)
    P.base[r] = nullptr;
  P.base[S.rank] = S.region;
  // p_ becomes the rank's block inside the region
  if (!pf->external)
    cudaFree(pf->d_p);
  pf->d_p = reinterpret_cast<float*>(S.region + P.off_p);
  pf->external = true;
  pf->cap = S.n_max;
  pf->n = S.hi - S.lo;
  // the scratch of the replicated chains, sized for the capacity NOW: a cudaFree / cudaMalloc inside a collective call would
  // synchronise the device while another rank of this process spins in wait_kernel for a publish that is not launched yet
  CU(ensure(pf->d_prefix, pf->prefix_cap, n_capacity));
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(n_capacity)));
  S.cur = 0;
  S.on = true;
  S.attached = (world == 1);
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_shard_ipc_handle(amcl3d_b200_pf_t pf, void* handle64, size_t cap)
{
  if (!pf || !handle64 || cap < sizeof(cudaIpcMemHandle_t))
    return fail(AMCL3D_B200_ERR_ARG, "handle buffer must hold 64 bytes");
  if (!pf->shard.on)
    return fail(AMCL3D_B200_ERR_STATE, "not sharded");
  if (int rc = use_device(pf->grid->device))
    return rc;
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, pf->shard.region));
  memcpy(handle64, &h, sizeof(h));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_shard_attach_ipc(amcl3d_b200_pf_t pf, const void* handles, size_t stride)
{
  if (!pf || !handles || stride < sizeof(cudaIpcMemHandle_t))
    return fail(AMCL3D_B200_ERR_ARG, "bad handle array");
  if (!pf->shard.on)
    return fail(AMCL3D_B200_ERR_STATE, "not sharded");
  if (int rc = use_device(pf->grid->device))
    return rc;
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  if (S.attached && S.world > 1)
    return fail(AMCL3D_B200_ERR_STATE, "peers are already attached");
  for (uint32_t r = 0; r < S.world; ++r)
  {
    if (r == S.rank)
      continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const char*>(handles) + (size_t)r * stride, sizeof(h));
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    S.opened[r] = ptr;
    S.peers.base[r] = reinterpret_cast<char*>(ptr);
  }
  S.attached = true;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_shard_attach_local(amcl3d_b200_pf_t* pfs, int world)
{
  if (!pfs || world < 1 || world > (int)kMaxShardWorld)
    return fail(AMCL3D_B200_ERR_ARG, "bad handle array");
  for (int r = 0; r < world; ++r)
    if (!pfs[r] || !pfs[r]->shard.on || pfs[r]->shard.world != (uint32_t)world || pfs[r]->shard.rank != (uint32_t)r ||
        pfs[r]->shard.n_total != pfs[0]->shard.n_total || pfs[r]->shard.n_cap != pfs[0]->shard.n_cap)
      return fail(AMCL3D_B200_ERR_ARG, "handles must be shard_init'ed as ranks 0..world-1 of the same set");
  for (int a = 0; a < world; ++a)
  {
    if (int rc = use_device(pfs[a]->grid->device))
      return rc;
    for (int b = 0; b < world; ++b)
    {
      if (a == b)
        continue;
      const int da = pfs[a]->grid->device, db = pfs[b]->grid->device;
      if (da != db)
      {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, da, db));
        if (!can)
          return fail(AMCL3D_B200_ERR_CUDA, "no peer access between the devices of the shard");
        cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail_cuda(e, "cudaDeviceEnablePeerAccess");
        (void)cudaGetLastError();
      }
      pfs[a]->shard.peers.base[b] = pfs[b]->shard.region;
    }
    pfs[a]->shard.attached = true;
  }
  return AMCL3D_B200_OK;
}

namespace
{
// A collective call runs in two halves: PUBLISH (this rank's contribution stored into every rank's region, flag raised)
// and REST (everything that reads what the other ranks published).  One process per GPU (CUDA IPC): one call does both
// and a spinning wait_kernel orders them across processes (kSpin).  Ranks that are handles of ONE process: the group call
// issues the first half on every rank, orders the streams with CUDA events, then issues the second half -- no kernel ever
// spins on something the host has not launched yet (a spinning kernel deadlocks against anything that synchronises the
// device in between: a lazily loaded module at a kernel's first launch, a cudaFree, or two streams sharing a hardware queue).
enum : int { kPublish = 1, kRest = 2, kSpin = 4, kWhole = kPublish | kRest | kSpin };

// this rank's weights (+ raw range weights, isIntoMap flags) into every rank's dense copy
int shard_publish(amcl3d_b200_pf_t pf, const void* d_wr_local)
{
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  S.epoch += 1;
  CU(launch_shard_publish(pf->grid->map, S.peers, pf->d_p, reinterpret_cast<const float*>(d_wr_local), (uint32_t)pf->n,
                          (uint32_t)S.lo, (uint32_t)S.n_cap, S.epoch & 1u, S.epoch, S.rank, S.d_done, pf->stream));
  pf->launches += 1;
  return AMCL3D_B200_OK;
}
int shard_spin(amcl3d_b200_pf_t pf, uint32_t kind, uint32_t epoch)
{
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  CU(launch_shard_wait(reinterpret_cast<const uint32_t*>(S.region + S.peers.off_flags), kind, S.world, epoch, S.d_error, pf->stream));
  pf->launches += 1;
  return AMCL3D_B200_OK;
}
int shard_ready(amcl3d_b200_pf_t pf)
{
  if (!pf)
    return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  if (!S.on || !S.attached)
    return fail(AMCL3D_B200_ERR_STATE, "shard not attached to its peers");
  if (pf->n != S.hi - S.lo)
    return fail(AMCL3D_B200_ERR_STATE, "p_ must hold this rank's block of the partition");
  return use_device(pf->grid->device);
}

int shard_normalize_impl(amcl3d_b200_pf_t pf, int phases)
{
  if (int rc = shard_ready(pf))
    return rc;
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  const uint64_t N = S.n_total;
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(S.n_cap)));  // sized at shard_init: no-op
  Timed t(pf);
  if (phases & kPublish)
    if (int rc = shard_publish(pf, nullptr))
      return rc;
  if (!(phases & kRest))
    return AMCL3D_B200_OK;
  if (phases & kSpin)
    if (int rc = shard_spin(pf, 0, S.epoch))
      return rc;
  const float* w = reinterpret_cast<const float*>(S.region + S.peers.off_w) + (size_t)(S.epoch & 1u) * S.n_cap;
  // ParticleFilter.cpp:171-178 over ALL particles (same chain on every rank), :180-195 on the rank's own block
  CU(launch_chain_dense(w, N, 1, nullptr, pf->d_scalars + 0, pf->d_dense, pf->chain_mode, pf->stream));
  CU(launch_normalize(pf->grid->map, pf->d_p, pf->n, pf->d_scalars + 0, pf->stream));
  pf->launches += 2;
  pf->launches += chain_extra_launches_take();
  return AMCL3D_B200_OK;
}

int shard_resample_impl(amcl3d_b200_pf_t pf, float u01, int normalize, const void* d_wr_local, float alpha, bool want_idx, int phases)
{
  if (int rc = shard_ready(pf))
    return rc;
  if (!normalize && d_wr_local)
    return fail(AMCL3D_B200_ERR_ARG, "the range fusion is part of the normalising call");
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  const uint64_t N = S.n_total, n = pf->n;
  CU(ensure(pf->d_prefix, pf->prefix_cap, S.n_cap));  // sized at shard_init: no-ops
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(S.n_cap)));
  int32_t* d_idx = nullptr;
  if (want_idx)
  {
    CU(ensure(pf->d_idx, pf->idx_cap, n));
    d_idx = pf->d_idx;
  }
  Timed t(pf);
  if (phases & kPublish)
    if (int rc = shard_publish(pf, d_wr_local))
      return rc;
  if (!(phases & kRest))
    return AMCL3D_B200_OK;
  if (phases & kSpin)
    if (int rc = shard_spin(pf, 0, S.epoch))
      return rc;
  const uint32_t parity = S.epoch & 1u;
  float* out = reinterpret_cast<float*>(S.region + S.peers.off_p + (size_t)(S.cur ^ 1u) * S.peers.p_bytes);
  float* w = reinterpret_cast<float*>(S.region + S.peers.off_w) + (size_t)parity * S.n_cap;
  float* wr = reinterpret_cast<float*>(S.region + S.peers.off_wr) + (size_t)parity * S.n_cap;
  const uint8_t* inmap = reinterpret_cast<const uint8_t*>(S.region + S.peers.off_inmap) + (size_t)parity * S.n_cap;
  // the reference's sequential chains over ALL particles, replicated on every rank: same bits as one GPU
  if (d_wr_local)
  {
    CU(launch_chain_dense(w, N, 0, nullptr, pf->d_scalars + 0, pf->d_dense, pf->chain_mode, pf->stream));
    CU(launch_chain_dense(wr, N, 0, nullptr, pf->d_scalars + 1, pf->d_dense, pf->chain_mode, pf->stream));
    CU(launch_dense_fuse(w, wr, N, pf->d_scalars + 0, pf->d_scalars + 1, alpha, pf->stream));
    pf->launches += 3;
  }
  if (normalize)
  {
    CU(launch_chain_dense(w, N, 1, nullptr, pf->d_scalars + 0, pf->d_dense, pf->chain_mode, pf->stream));  // ParticleFilter.cpp:171-178
    CU(launch_dense_normalize(w, inmap, N, pf->d_scalars + 0, pf->stream));                                // :180-195
    pf->launches += 2;
  }
  CU(launch_chain_dense(w, N, 0, pf->d_prefix, nullptr, pf->d_dense, pf->chain_mode, pf->stream));  // :235-247
  // this rank's output slice; the selected particles are read from their owners over NVLink
  CU(launch_shard_search(S.peers, pf->d_prefix, N, u01, S.lo, S.hi, S.cur, out, d_idx, pf->stream));
  pf->launches += 2;
  pf->launches += chain_extra_launches_take();
  S.cur ^= 1u;
  pf->d_p = out;
  return AMCL3D_B200_OK;
}

int shard_mean_impl(amcl3d_b200_pf_t pf, float* dst, int phases)
{
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  if (!S.on || !S.attached)
    return fail(AMCL3D_B200_ERR_STATE, "shard not attached to its peers");
  if (int rc = use_device(pf->grid->device))
    return rc;
  Timed t(pf);
  if (phases & kPublish)
  {
    S.mean_epoch += 1;
    CU(launch_mean(pf->d_p, pf->n, 0, S.d_part7, pf->d_scratch, pf->stream));
    CU(launch_shard_publish_mean(S.peers, S.d_part7, S.mean_epoch & 1u, S.mean_epoch, S.rank, pf->stream));
    pf->launches += 3;
  }
  if (!(phases & kRest))
    return AMCL3D_B200_OK;
  if (phases & kSpin)
    if (int rc = shard_spin(pf, 1, S.mean_epoch))
      return rc;
  CU(launch_shard_combine_mean(reinterpret_cast<const float*>(S.region + S.peers.off_mean) + (size_t)(S.mean_epoch & 1u) * kMaxShardWorld * 8,
                               S.world, dst, pf->stream));
  pf->launches += 1;
  return AMCL3D_B200_OK;
}

// the group form for ranks that are handles of one process: first half on every rank, events, second half on every rank
int local_collective(amcl3d_b200_pf_t* pfs, int world, const std::function<int(int, int)>& half)
{
  if (!pfs || world < 1 || world > (int)kMaxShardWorld)
    return fail(AMCL3D_B200_ERR_ARG, "bad handle array");
  for (int r = 0; r < world; ++r)
    if (!pfs[r])
      return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
  for (int r = 0; r < world; ++r)
  {
    if (int rc = half(r, kPublish))
      return rc;
    if (int rc = use_device(pfs[r]->grid->device))
      return rc;
    CU(cudaEventRecord(pfs[r]->ev_sync, pfs[r]->stream));
  }
  for (int r = 0; r < world; ++r)
  {
    if (int rc = use_device(pfs[r]->grid->device))
      return rc;
    for (int q = 0; q < world; ++q)
      if (q != r)
        CU(cudaStreamWaitEvent(pfs[r]->stream, pfs[q]->ev_sync, 0));
    if (int rc = half(r, kRest))
      return rc;
  }
  return AMCL3D_B200_OK;
}
}  // namespace

int amcl3d_b200_pf_shard_normalize(amcl3d_b200_pf_t pf, float* wtp)
{
  NvtxRange nvtx_range("particleNormalize (sharded)");
  if (int rc = shard_normalize_impl(pf, kWhole))
    return rc;
  if (!wtp)
  {
    if (pf->async_mode)
      return AMCL3D_B200_OK;
    return shard_check_error(pf);
  }
  CU(cudaMemcpyAsync(wtp, pf->d_scalars, sizeof(float), cudaMemcpyDeviceToHost, pf->stream));
  return shard_check_error(pf);
}

int amcl3d_b200_pf_shard_resample(amcl3d_b200_pf_t pf, float u01, int normalize, const void* d_wr_local, float alpha, int32_t* idx)
{
  NvtxRange nvtx_range("Resample (sharded)");
  if (int rc = shard_resample_impl(pf, u01, normalize, d_wr_local, alpha, idx != nullptr, kWhole))
    return rc;
  if (idx && pf->n)
  {
    CU(cudaMemcpyAsync(idx, pf->d_idx, sizeof(int32_t) * pf->n, cudaMemcpyDeviceToHost, pf->stream));
    return shard_check_error(pf);
  }
  if (pf->async_mode)
    return AMCL3D_B200_OK;
  return shard_check_error(pf);
}

// fast (non-exact) mode, see shard.cu: two scalar exchanges + source-side placement, no rank reads another rank's weights
int amcl3d_b200_pf_shard_resample_fast(amcl3d_b200_pf_t pf, float u01, int32_t* idx)
{
  NvtxRange nvtx_range("Resample (sharded, fast)");
  if (int rc = shard_ready(pf))
    return rc;
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  const uint64_t N = S.n_total, n = pf->n;
  if (N < S.world)
    return fail(AMCL3D_B200_ERR_ARG, "fewer particles than ranks");
  CU(ensure(pf->d_prefix, pf->prefix_cap, S.n_cap));  // sized at shard_init: no-ops
  CU(ensure(pf->d_dense, pf->dense_cap, chain_scratch_floats(S.n_cap)));
  S.fast_epoch += 1;
  const uint32_t e = S.fast_epoch, parity = e & 1u;
  const uint32_t* flags = reinterpret_cast<const uint32_t*>(S.region + S.peers.off_flags);
  const float* sums = reinterpret_cast<const float*>(S.region + S.peers.off_sums);
  float* out = reinterpret_cast<float*>(S.region + S.peers.off_p + (size_t)(S.cur ^ 1u) * S.peers.p_bytes);
  float* sc = pf->d_scalars;  // [0] local sum -> [4],[5] global sum / before; [1] local total -> [6],[7] total / base
  {
    Timed t(pf);
    // 1. weight sum: local sequential chain, combined over the ranks in rank order
    CU(launch_chain_prefix(pf->d_p, n, 1, nullptr, sc + 0, pf->d_dense, pf->chain_mode, pf->stream));
    CU(launch_shard_publish_scalar(S.peers, sc + 0, 0, parity, e, S.rank, pf->stream));
    CU(launch_shard_wait(flags, 2, S.world, e, S.d_error, pf->stream));
    CU(launch_shard_combine_scalar(sums + ((size_t)0 * 2 + parity) * kMaxShardWorld, S.world, S.rank, sc + 4, pf->stream));
    CU(launch_normalize(pf->grid->map, pf->d_p, n, sc + 4, pf->stream));  // ParticleFilter.cpp:180-195 with the global sum
    // 2. local inclusive prefix of the normalised weights; its total places the rank's stretch in the global prefix
    CU(launch_chain_prefix(pf->d_p, n, 0, pf->d_prefix, sc + 1, pf->d_dense, pf->chain_mode, pf->stream));
    CU(launch_shard_publish_scalar(S.peers, sc + 1, 1, parity, e, S.rank, pf->stream));
    CU(launch_shard_wait(flags, 3, S.world, e, S.d_error, pf->stream));
    CU(launch_shard_combine_scalar(sums + ((size_t)1 * 2 + parity) * kMaxShardWorld, S.world, S.rank, sc + 6, pf->stream));
    // 3. place this rank's offspring into the output blocks of their owners; wait until every rank has placed its own
    CU(launch_shard_push(S.peers, pf->d_p, pf->d_prefix, (uint32_t)n, S.lo, N, sc + 6, u01, S.rank, S.cur ^ 1u, e, S.d_done,
                         pf->stream));
    CU(launch_shard_wait(flags, 4, S.world, e, S.d_error, pf->stream));
    pf->launches += 11;
    pf->launches += chain_extra_launches_take();
  }
  S.cur ^= 1u;
  pf->d_p = out;
  if (idx && n)
  {
    CU(cudaMemcpyAsync(idx, S.region + S.peers.off_idx, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, pf->stream));
    return shard_check_error(pf);
  }
  if (pf->async_mode)
    return AMCL3D_B200_OK;
  return shard_check_error(pf);
}

int amcl3d_b200_pf_shard_mean(amcl3d_b200_pf_t pf, float out7[7], void* d_out7)
{
  NvtxRange nvtx_range("getMean (sharded)");
  if (!pf || (!out7 && !d_out7))
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  float* dst = d_out7 ? reinterpret_cast<float*>(d_out7) : pf->d_scalars + 8;
  if (int rc = shard_mean_impl(pf, dst, kWhole))
    return rc;
  if (out7)
  {
    CU(cudaMemcpyAsync(out7, dst, 7 * sizeof(float), cudaMemcpyDeviceToHost, pf->stream));
    return shard_check_error(pf);
  }
  return finish_call(pf);
}

// ranks = handles of ONE process: both halves of the collective call are issued on every handle in stream order, with CUDA
// events between them instead of a spinning kernel (see kPublish / kRest above), then waited for
int amcl3d_b200_pf_shard_normalize_local(amcl3d_b200_pf_t* pfs, int world)
{
  NvtxRange nvtx_range("particleNormalize (sharded, one process)");
  if (int rc = local_collective(pfs, world, [&](int r, int phases) { return shard_normalize_impl(pfs[r], phases); }))
    return rc;
  for (int r = 0; r < world; ++r)
    if (!pfs[r]->async_mode)
      if (int rc = amcl3d_b200_pf_sync(pfs[r]))
        return rc;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_shard_resample_local(amcl3d_b200_pf_t* pfs, int world, float u01, int normalize, const void* const* d_wr_local,
                                        float alpha)
{
  NvtxRange nvtx_range("Resample (sharded, one process)");
  if (int rc = local_collective(pfs, world, [&](int r, int phases) {
        return shard_resample_impl(pfs[r], u01, normalize, d_wr_local ? d_wr_local[r] : nullptr, alpha, false, phases);
      }))
    return rc;
  for (int r = 0; r < world; ++r)
    if (!pfs[r]->async_mode)
      if (int rc = amcl3d_b200_pf_sync(pfs[r]))
        return rc;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_shard_mean_local(amcl3d_b200_pf_t* pfs, int world, float out7[7])
{
  NvtxRange nvtx_range("getMean (sharded, one process)");
  if (!out7)
    return fail(AMCL3D_B200_ERR_ARG, "bad argument");
  if (int rc = local_collective(pfs, world, [&](int r, int phases) { return shard_mean_impl(pfs[r], pfs[r]->d_scalars + 8, phases); }))
    return rc;
  for (int r = 0; r < world; ++r)
    if (int rc = amcl3d_b200_pf_sync(pfs[r]))
      return rc;
  if (int rc = use_device(pfs[0]->grid->device))
    return rc;
  CU(cudaMemcpy(out7, pfs[0]->d_scalars + 8, 7 * sizeof(float), cudaMemcpyDeviceToHost));
  return AMCL3D_B200_OK;
}

// getMean as ONE sequential f32 chain over the global particle order (ParticleFilter.cpp:198-216), i.e. the bits one GPU
// produces: rank r continues the six sums (and the running maximum) where rank r - 1 stopped -- its kernel waits for that
// rank's event and reads its seven floats (peer memory).  Serial over the ranks by construction: N_total x one FADD latency.
int amcl3d_b200_pf_shard_mean_exact_local(amcl3d_b200_pf_t* pfs, int world, float out7[7])
{
  NvtxRange nvtx_range("getMean (sharded, one process, exact chain)");
  if (!pfs || !out7 || world < 1 || world > (int)kMaxShardWorld)
    return fail(AMCL3D_B200_ERR_ARG, "bad argument");
  for (int r = 0; r < world; ++r)
  {
    if (!pfs[r])
      return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
    const amcl3d_b200_pf_s::Shard& S = pfs[r]->shard;
    if (!S.on || !S.attached || S.world != (uint32_t)world || S.rank != (uint32_t)r)
      return fail(AMCL3D_B200_ERR_STATE, "handles must be the attached ranks 0..world-1 of one set");
  }
  for (int r = 0; r < world; ++r)
  {
    amcl3d_b200_pf_t pf = pfs[r];
    if (int rc = use_device(pf->grid->device))
      return rc;
    if (r > 0)
      CU(cudaStreamWaitEvent(pf->stream, pfs[r - 1]->ev_sync, 0));
    {
      Timed t(pf);
      CU(launch_mean_exact_carry(pf->d_p, pf->n, r > 0 ? pfs[r - 1]->d_scalars + 16 : nullptr, pf->d_scalars + 16, pf->stream));
      pf->launches += 1;
    }
    CU(cudaEventRecord(pf->ev_sync, pf->stream));
  }
  amcl3d_b200_pf_t last = pfs[world - 1];
  CU(cudaMemcpyAsync(out7, last->d_scalars + 16, 7 * sizeof(float), cudaMemcpyDeviceToHost, last->stream));
  CU(cudaStreamSynchronize(last->stream));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_shard_resize(amcl3d_b200_pf_t pf, uint64_t n_total)
{
  if (!pf || !pf->shard.on)
    return fail(AMCL3D_B200_ERR_STATE, "not sharded");
  amcl3d_b200_pf_s::Shard& S = pf->shard;
  if (n_total == 0 || n_total > S.n_cap)
    return fail(AMCL3D_B200_ERR_ARG, "particle count outside the capacity of the exchange region");
  S.n_total = n_total;
  shard_bounds(n_total, S.world, S.rank, &S.lo, &S.hi);
  pf->n = S.hi - S.lo;
  return AMCL3D_B200_OK;
}

namespace
{
cudaError_t copy_between(void* dst, int dst_device, const void* src, int src_device, size_t bytes, cudaStream_t stream)
{
  if (dst_device == src_device)
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream);
  return cudaMemcpyPeerAsync(dst, dst_device, src, src_device, bytes, stream);
}
int local_group_ok(amcl3d_b200_pf_t* pfs, int world, amcl3d_b200_pf_t other)
{
  if (!pfs || !other || world < 1 || world > (int)kMaxShardWorld)
    return fail(AMCL3D_B200_ERR_ARG, "bad handle array");
  if (other->shard.on)
    return fail(AMCL3D_B200_ERR_ARG, "the handle outside the group must be a plain (unsharded) one");
  for (int r = 0; r < world; ++r)
  {
    if (!pfs[r])
      return fail(AMCL3D_B200_ERR_ARG, "NULL handle");
    const amcl3d_b200_pf_s::Shard& S = pfs[r]->shard;
    if (!S.on || !S.attached || S.world != (uint32_t)world || S.rank != (uint32_t)r || S.n_total != pfs[0]->shard.n_total)
      return fail(AMCL3D_B200_ERR_STATE, "handles must be the attached ranks 0..world-1 of one set");
  }
  return AMCL3D_B200_OK;
}
}  // namespace

int amcl3d_b200_pf_shard_gather_local(amcl3d_b200_pf_t* pfs, int world, amcl3d_b200_pf_t dst)
{
  if (int rc = local_group_ok(pfs, world, dst))
    return rc;
  const uint64_t N = pfs[0]->shard.n_total;
  for (int r = 0; r < world; ++r)
    if (pfs[r]->n != pfs[r]->shard.hi - pfs[r]->shard.lo)
      return fail(AMCL3D_B200_ERR_STATE, "p_ must hold this rank's block of the partition");
  const int ddev = dst->grid->device;
  if (int rc = use_device(ddev))
    return rc;
  if (dst->external)
  {
    if (N > dst->cap)
      return fail(AMCL3D_B200_ERR_ARG, "bound device buffer too small for the gathered set");
  }
  else if (N > dst->cap)
  {
    CU(cudaStreamSynchronize(dst->stream));  // the old buffer may still be in use by launched work
    CU(ensure_particles(dst->d_p, dst->cap, N));
  }
  for (int r = 0; r < world; ++r)
  {
    amcl3d_b200_pf_t s = pfs[r];
    if (int rc = use_device(s->grid->device))
      return rc;
    CU(cudaEventRecord(s->ev_sync, s->stream));
    if (int rc = use_device(ddev))
      return rc;
    CU(cudaStreamWaitEvent(dst->stream, s->ev_sync, 0));
    if (s->n)
      CU(copy_between(dst->d_p + 7 * s->shard.lo, ddev, s->d_p, s->grid->device, sizeof(float) * 7 * s->n, dst->stream));
  }
  dst->n = N;
  // the shards must not overwrite their blocks before the copies have read them
  CU(cudaEventRecord(dst->ev_sync, dst->stream));
  for (int r = 0; r < world; ++r)
  {
    if (int rc = use_device(pfs[r]->grid->device))
      return rc;
    CU(cudaStreamWaitEvent(pfs[r]->stream, dst->ev_sync, 0));
  }
  if (int rc = use_device(ddev))
    return rc;
  return finish_call(dst);
}

int amcl3d_b200_pf_shard_scatter_local(amcl3d_b200_pf_t src, amcl3d_b200_pf_t* pfs, int world)
{
  if (int rc = local_group_ok(pfs, world, src))
    return rc;
  const uint64_t N = src->n;
  if (N < (uint64_t)world || N > pfs[0]->shard.n_cap)
    return fail(AMCL3D_B200_ERR_ARG, "particle count outside [world, capacity of the exchange regions]");
  const int sdev = src->grid->device;
  if (int rc = use_device(sdev))
    return rc;
  CU(cudaEventRecord(src->ev_sync, src->stream));
  for (int r = 0; r < world; ++r)
  {
    amcl3d_b200_pf_t d = pfs[r];
    if (int rc = amcl3d_b200_pf_shard_resize(d, N))
      return rc;
    if (int rc = use_device(d->grid->device))
      return rc;
    CU(cudaStreamWaitEvent(d->stream, src->ev_sync, 0));
    if (d->n)
      CU(copy_between(d->d_p, d->grid->device, src->d_p + 7 * d->shard.lo, sdev, sizeof(float) * 7 * d->n, d->stream));
    CU(cudaEventRecord(d->ev_sync, d->stream));
  }
  // src must not change its set before the copies have read it
  if (int rc = use_device(sdev))
    return rc;
  for (int r = 0; r < world; ++r)
    CU(cudaStreamWaitEvent(src->stream, pfs[r]->ev_sync, 0));
  for (int r = 0; r < world; ++r)
    if (!pfs[r]->async_mode)
      if (int rc = amcl3d_b200_pf_sync(pfs[r]))
        return rc;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_shard_info(amcl3d_b200_pf_t pf, uint64_t* lo, uint64_t* hi, uint64_t* region_bytes)
{
  if (!pf || !pf->shard.on)
    return fail(AMCL3D_B200_ERR_STATE, "not sharded");
  if (lo)
    *lo = pf->shard.lo;
  if (hi)
    *hi = pf->shard.hi;
  if (region_bytes)
    *region_bytes = pf->shard.region_bytes;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_last_kernel_ms(amcl3d_b200_pf_t pf, float* ms)
{
  if (!pf || !ms)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  *ms = 0.f;
  if (!pf->timed)
    return AMCL3D_B200_OK;
  CU(cudaEventSynchronize(pf->ev1));
  CU(cudaEventElapsedTime(ms, pf->ev0, pf->ev1));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_last_phase_ms(amcl3d_b200_pf_t pf, float ms[3])
{
  if (!pf || !ms)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  ms[0] = ms[1] = ms[2] = 0.f;
  if (!pf->phase_timed)
    return AMCL3D_B200_OK;
  CU(cudaEventSynchronize(pf->ev_phase[3]));
  for (int k = 0; k < 3; ++k)
    CU(cudaEventElapsedTime(&ms[k], pf->ev_phase[k], pf->ev_phase[k + 1]));
  return AMCL3D_B200_OK;
}

int amcl3d_b200_pf_launch_count(amcl3d_b200_pf_t pf, uint64_t* launches)
{
  if (!pf || !launches)
    return fail(AMCL3D_B200_ERR_ARG, "NULL argument");
  *launches = pf->launches;
  return AMCL3D_B200_OK;
}

int amcl3d_b200_bench_gather(amcl3d_b200_grid_t g, uint64_t window_cells, uint32_t blocks, uint32_t iters, int coherent,
                             float* ms, uint64_t* loads)
{
  if (!g || !ms || !blocks || !iters)
    return fail(AMCL3D_B200_ERR_ARG, "bad argument");
  if (!g->map.has_grid)
    return fail(AMCL3D_B200_ERR_STATE, "no grid");
  if (int rc = use_device(g->device))
    return rc;
  if (window_cells == 0 || window_cells > g->map.n_cells)
    window_cells = g->map.n_cells;
  float* d_sink = nullptr;
  CU(cudaMalloc(&d_sink, sizeof(float)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  // one untimed warm-up launch, then the timed one
  cudaError_t e = launch_gather_bench(g->d_prob, window_cells, blocks, iters, coherent, d_sink, g->stream);
  if (e == cudaSuccess)
    e = cudaEventRecord(e0, g->stream);
  if (e == cudaSuccess)
    e = launch_gather_bench(g->d_prob, window_cells, blocks, iters, coherent, d_sink, g->stream);
  if (e == cudaSuccess)
    e = cudaEventRecord(e1, g->stream);
  if (e == cudaSuccess)
    e = cudaEventSynchronize(e1);
  if (e == cudaSuccess)
    e = cudaEventElapsedTime(ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_sink);
  if (e != cudaSuccess)
    return fail_cuda(e, "gather bench");
  if (loads)
    *loads = (uint64_t)blocks * 256ull * iters * 8ull;
  return AMCL3D_B200_OK;
}

}  // extern "C"
