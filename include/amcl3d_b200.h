/*
 * amcl3d_b200.h -- C-ABI of the B200-native amcl3d particle-weighting path.
 *
 * One shared library (libamcl3d_b200.so, hand-written sm_100a CUDA) behind plain C entry points:
 * opaque handles, plain pointers and sizes, int status codes.  There is no CPU fallback: every
 * compute entry point launches CUDA kernels and fails with AMCL3D_B200_ERR_CUDA when no device is
 * usable.  The reference has no FFI of its own (it is a C++ static library); each entry point below
 * names the reference member function it stands behind (paths relative to the reference tree), and
 * include/amcl3d/{Grid3d,ParticleFilter,PointCloudTools}.h are the source-compatible C++ classes
 * that bind them.
 *
 * Layouts
 *   particle : 7 packed floats x,y,z,roll,pitch,yaw,w        (amcl3d/src/ParticleFilter.h:32-48)
 *   cloud    : floats, `stride` floats from one point to the next, x,y,z first (pcl::PointXYZI: stride 8)
 *   grid     : float prob per voxel, x fastest, then y, then z  (amcl3d/src/PointCloudTools.h:30-52)
 *
 * Threading: a handle is not thread-safe (neither is the reference); different handles are independent.
 * All functions are synchronous with respect to the host unless stated otherwise.
 */
#ifndef AMCL3D_B200_H
#define AMCL3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define AMCL3D_B200_ABI_VERSION 2

enum
{
  AMCL3D_B200_OK = 0,
  AMCL3D_B200_ERR_ARG = 1,    /* bad argument / handle */
  AMCL3D_B200_ERR_STATE = 2,  /* map or grid not set (the reference returns 0 / false here) */
  AMCL3D_B200_ERR_CUDA = 3,   /* CUDA runtime error, see amcl3d_b200_last_error() */
  AMCL3D_B200_ERR_IO = 4,     /* file open / short read */
  AMCL3D_B200_ERR_MISMATCH = 5, /* .grid sensor_dev differs (Grid3d.cpp:422-427) */
  AMCL3D_B200_ERR_EMPTY = 6   /* cloud with <= 1 point (PointCloudTools.cpp:89-91 throws) */
};

/* how `const auto sr = sin(roll)` binds (Grid3d.cpp:240-245) */
enum
{
  AMCL3D_B200_TRIG_F64 = 0, /* default: ::sin(double), f64 products rounded once to f32 */
  AMCL3D_B200_TRIG_F32 = 1  /* sinf/cosf and f32 products (device libm; not bit-pinned) */
};

/* grid builders */
enum
{
  AMCL3D_B200_GRID_QUIRK = 0,    /* computeGrid as written: prob = c1*expf(-(d^2)^2*c2) (PointCloudTools.cpp:160-162) */
  AMCL3D_B200_GRID_GAUSSIAN = 1, /* c1*expf(-d^2*c2): labelled extension */
  AMCL3D_B200_GRID_SPLAT = 2     /* computeGrid2: 27-neighbour max splat (PointCloudTools.cpp:176-254) */
};

typedef struct amcl3d_b200_grid_s* amcl3d_b200_grid_t;
typedef struct amcl3d_b200_pf_s* amcl3d_b200_pf_t;

/* rosinrange_msg/msg/RangePose.msg:1-6 -- the fields the range weight consumes */
typedef struct
{
  double position[3];
  double range;
} amcl3d_b200_beacon_t;

typedef struct
{
  double min[3], max[3], resol, sensor_dev; /* PointCloudInfo + Grid3dInfo::sensor_dev */
  uint32_t size[3];                         /* Grid3dInfo::size_{x,y,z} */
  uint64_t n_cells;                         /* size_x*size_y*size_z (64-bit; the reference wraps at 2^32) */
  int has_map;                              /* pc_info_ set */
  int has_grid;                             /* grid_info_ set */
  int device;
  int rep;                                  /* what the weighting kernel reads: 0 linear f32, 1 / 2 = lossless
                                               8- / 16-bit dictionary codes in 128-byte voxel bricks */
  uint32_t lut_n;                           /* distinct field values (dictionary size), 0 for rep 0 */
  uint64_t coded_bytes;                     /* size of the coded copy */
} amcl3d_b200_grid_info_t;

/* ---- runtime ------------------------------------------------------------------------------ */
int amcl3d_b200_abi_version(void);
/* message of the last failing call on this thread */
const char* amcl3d_b200_last_error(void);
/* any argument may be NULL */
int amcl3d_b200_device_info(int device, int* device_count, int* sm_count, int* cc_major, int* cc_minor,
                            uint64_t* l2_bytes, uint64_t* mem_bytes);

/* Page-locks / releases a host buffer that the caller keeps handing to set_particles / get_particles (the std::vector
 * behind amcl3d::ParticleFilter::p_): the copies become single DMA transfers instead of staged pageable copies. */
int amcl3d_b200_host_pin(void* ptr, size_t bytes);
int amcl3d_b200_host_unpin(void* ptr);

/* cudaLimitMaxL2FetchGranularity of the device (32, 64 or 128 bytes) -- a measurement knob.  The library never
 * changes this limit of the host application on its own (only when AMCL3D_B200_L2_FETCH is set in the
 * environment or through this call).  Measured on B200 (profiles/r2_gather_probe_ncu.txt): a scattered 1-byte
 * gather that misses L2 costs a full 128-byte line of DRAM traffic at every setting, 32, 64 and 128 alike. */
int amcl3d_b200_set_l2_fetch_granularity(int device, int bytes);
int amcl3d_b200_get_l2_fetch_granularity(int device, int* bytes);

/* ---- Grid3d (amcl3d/src/Grid3d.{h,cpp}, PointCloudTools.{h,cpp}) ---------------------------- */
int amcl3d_b200_grid_create(int device, amcl3d_b200_grid_t* out);
int amcl3d_b200_grid_destroy(amcl3d_b200_grid_t g);
int amcl3d_b200_grid_get_info(amcl3d_b200_grid_t g, amcl3d_b200_grid_info_t* info);

/* computePointCloud (PointCloudTools.cpp:84-109): min/max of the map cloud by a device reduction,
 * installs them with `resol` as the handle's PointCloudInfo and sizes the grid (PointCloudTools.cpp:
 * 120-130).  n <= 1 -> AMCL3D_B200_ERR_EMPTY.  min/max (optional) receive the bounds. */
int amcl3d_b200_grid_compute_bounds(amcl3d_b200_grid_t g, const float* xyz, uint32_t stride, uint64_t n,
                                    double resol, double min[3], double max[3]);
/* installs PointCloudInfo directly */
int amcl3d_b200_grid_set_map(amcl3d_b200_grid_t g, const double min[3], const double max[3], double resol);

/* computeGrid / computeGrid2 (PointCloudTools.cpp:111-174 / 176-254) on the device.  QUIRK/GAUSSIAN:
 * exact Euclidean distance transform on the lattice of pitch resol/sub (sub = 1 or 2) followed by the fused
 * Gaussian writer.  Map points are SNAPPED to that lattice first: the result is bit-exact with the reference's
 * kd-tree build when the points already lie on it (maps made of voxel centres / corners), and for arbitrary clouds the
 * nearest-point distance of a voxel differs from the reference's by at most sqrt(3)/2 * resol/sub (0.43 resol at
 * sub = 2), the field accordingly (tests/test_gpu_grid.py::test_off_lattice_map_points_stay_within_the_stated_snap_bound).  SPLAT takes arbitrary points.  keep_edt != 0 keeps the int32 squared distances
 * (units (resol/sub)^2) for amcl3d_b200_grid_download_edt and runs the exact lower-envelope passes; without it
 * the y / z passes only search the window beyond which the probability is exactly 0.0f (min(d2, cap)) -- the
 * float grid is the same, bit for bit (environment AMCL3D_B200_EDT_FAST=0 forces the exact passes). */
int amcl3d_b200_grid_build(amcl3d_b200_grid_t g, const float* xyz, uint32_t stride, uint64_t n,
                           double sensor_dev, int mode, int sub, int keep_edt);
/* the same build from map points already resident in DEVICE memory (no host round trip; used in place) */
int amcl3d_b200_grid_build_device(amcl3d_b200_grid_t g, const void* d_xyz, uint32_t stride, uint64_t n,
                                  double sensor_dev, int mode, int sub, int keep_edt);
/* device time in ms of the kernels of the last grid_build / grid_build_device on this handle (CUDA events on the
 * handle's stream, after the H2D copy of the points) */
int amcl3d_b200_grid_last_build_ms(amcl3d_b200_grid_t g, float* ms);
int amcl3d_b200_grid_download_edt(amcl3d_b200_grid_t g, int32_t* d2, uint64_t cap);
/* per grid handle: 1 (default) capped passes when the EDT is not kept, 0 the exact envelope passes always (the float
 * grid is the same bit for bit; A/B switch).  Environment AMCL3D_B200_EDT_FAST=0 sets the default for new handles. */
int amcl3d_b200_grid_set_edt_fast(amcl3d_b200_grid_t g, int enable);

/* Replica of a built / loaded grid on another device of the same box: the grid is replicated, the particles are
 * sharded.  Peer copy of the coded cells (what update() reads); the linear float cells travel only with
 * with_float != 0 or when the grid has no codes (download / save of a codes-only replica fail with ERR_STATE). */
int amcl3d_b200_grid_clone(amcl3d_b200_grid_t src, int device, int with_float, amcl3d_b200_grid_t* out);

/* loadGrid's payload (Grid3d.cpp:404-442): cells + sizes + sensor_dev from host memory */
int amcl3d_b200_grid_upload(amcl3d_b200_grid_t g, const float* prob, const uint32_t size[3], double sensor_dev);
int amcl3d_b200_grid_download(amcl3d_b200_grid_t g, float* prob, uint64_t cap);
/* saveGrid / loadGrid file format (Grid3d.cpp:388-395): u32 sx, sy, sz, f64 sensor_dev, cells */
int amcl3d_b200_grid_save_file(amcl3d_b200_grid_t g, const char* path);
int amcl3d_b200_grid_load_file(amcl3d_b200_grid_t g, const char* path, double sensor_dev);

/* Grid3d::isIntoMap (Grid3d.cpp:365-372); *inside = 0 when no map is set */
int amcl3d_b200_grid_is_into_map(amcl3d_b200_grid_t g, float x, float y, float z, int* inside);
/* Grid3d::computeCloudWeight 7-arg (Grid3d.cpp:234-301) for one pose; *weight = 0 when the grid or
 * the map is missing (Grid3d.cpp:237-238) */
int amcl3d_b200_grid_cloud_weight(amcl3d_b200_grid_t g, const float* cloud, uint32_t stride, uint32_t npts,
                                  float tx, float ty, float tz, float roll, float pitch, float yaw,
                                  int trig_mode, float* weight);
/* enable (default) / disable the lossless dictionary-coded copy that grid_build makes for the weighting
 * kernel (an A/B switch: results are bit-identical either way).  Environment AMCL3D_B200_CODES=0 sets
 * the default for new handles. */
int amcl3d_b200_grid_use_codes(amcl3d_b200_grid_t g, int enable);
/* raw device pointer of the cells (for callers that keep data resident, e.g. NCCL broadcast) */
int amcl3d_b200_grid_device_ptr(amcl3d_b200_grid_t g, void** d_prob);

/* ---- ParticleFilter (amcl3d/src/ParticleFilter.{h,cpp}) -------------------------------------- */
int amcl3d_b200_pf_create(amcl3d_b200_grid_t g, amcl3d_b200_pf_t* out);
int amcl3d_b200_pf_destroy(amcl3d_b200_pf_t pf);
int amcl3d_b200_pf_set_trig_mode(amcl3d_b200_pf_t pf, int trig_mode);
/* How this handle evaluates the sequential f32 recurrences of particleNormalize / resample / uniformSample:
 * 0 (default) the exact binade scan spread over many CTAs, 1 the plain serial chain kernel, 2 the exact binade scan
 * on one CTA.  All give the reference's bits; the switch exists for A/B timing and cross-checking. */
int amcl3d_b200_pf_set_chain_mode(amcl3d_b200_pf_t pf, int mode);
/* Stream order instead of call order.  The reference API is synchronous and so is every entry point by default.
 * With async enabled, the calls that hand nothing back to the host -- update (in_bounds == NULL), update_split,
 * fuse_range, normalize (wtp == NULL), scale_by_max, resample (idx == NULL) -- only enqueue their kernels on the
 * handle's stream and return; anything that returns a value (mean, best, get_particles, a non-NULL out pointer)
 * still waits for it, which also orders everything enqueued before.  amcl3d_b200_pf_sync waits explicitly.
 * Lets a caller that keeps the particle set resident run frames back to back without a host round trip per stage. */
int amcl3d_b200_pf_set_async(amcl3d_b200_pf_t pf, int enable);
int amcl3d_b200_pf_sync(amcl3d_b200_pf_t pf);
/* enable (default) / disable the two-phase update: phase 1 sweeps the cloud in Morton order (so that all
 * particles work in the same neighbourhood of the map at the same time) and parks 1-byte codes, phase 2 adds
 * them in the original cloud order -- bit-identical weights.  Used for u8-coded (device-built) grids and clouds of at
 * least 256 points, from ONE particle on: it parallelises over (particle tile, cloud chunk) pairs, which is what a
 * small set needs too (600 x 2 000: 0.34 -> 0.07 ms; 100 k x 10 k: 27.9 -> 4.1 ms against the one-launch path on the
 * float grid).  Takes effect at the next update (the Morton order is built lazily).  Environment
 * AMCL3D_B200_TWO_PHASE=0 sets the default for new handles. */
int amcl3d_b200_pf_set_two_phase(amcl3d_b200_pf_t pf, int enable);
/* run this handle's kernels on a caller-owned cudaStream_t (NULL = the handle's own non-blocking stream; pass
 * cudaStreamLegacy / cudaStreamPerThread explicitly to order with the default stream) */
int amcl3d_b200_pf_set_stream(amcl3d_b200_pf_t pf, void* cuda_stream);

/* p_ mirror: host <-> device copies of the particle set */
int amcl3d_b200_pf_set_particles(amcl3d_b200_pf_t pf, const float* aos7, uint64_t n);
int amcl3d_b200_pf_get_particles(amcl3d_b200_pf_t pf, float* aos7, uint64_t cap, uint64_t* n);
int amcl3d_b200_pf_size(amcl3d_b200_pf_t pf, uint64_t* n);
/* adopt a caller-owned DEVICE buffer of `capacity` particles as p_ (e.g. a torch tensor that also
 * takes part in NCCL collectives); n of them are live.  The handle never frees it. */
int amcl3d_b200_pf_bind_device(amcl3d_b200_pf_t pf, void* d_aos7, uint64_t n, uint64_t capacity);
int amcl3d_b200_pf_device_ptr(amcl3d_b200_pf_t pf, void** d_aos7, uint64_t* n);

/* the sensor cloud: copied H2D and packed (x,y,z,0) per point for the kernels */
int amcl3d_b200_pf_set_cloud(amcl3d_b200_pf_t pf, const float* cloud, uint32_t stride, uint32_t npts);
/* same from a DEVICE buffer (no host round trip) */
int amcl3d_b200_pf_set_cloud_device(amcl3d_b200_pf_t pf, const void* d_cloud, uint32_t stride, uint32_t npts);
/* The input downsample in front of update(): ds_.setLeafSize(voxel_size x3) + ds_.filter(), amcl3d.cpp:37,51-52
 * (pcl::VoxelGrid<PointXYZI> centroid filter with default settings, PCL filters/impl/voxel_grid.hpp) -- the raw
 * cloud goes H2D once, is reduced to one centroid per occupied leaf on the device and becomes the cloud of the
 * following update() without a host round trip.  intensity_offset: float index of the intensity inside a point
 * (4 for pcl::PointXYZI's 8-float layout, 3 for packed xyzi, < 0: none).  skip_nonfinite = !cloud.is_dense.
 * Cell set, order and counts are exact; centroids are f32 sums in input order / (float)count (PCL's order inside
 * a cell is that of an unstable sort, so a PCL run can differ in the last ulp).  If PCL's "leaf size too small"
 * guard trips the input passes through unchanged, as in PCL.  npts_out: points now set. */
int amcl3d_b200_pf_set_cloud_downsampled(amcl3d_b200_pf_t pf, const float* cloud, uint32_t stride, int intensity_offset,
                                         uint32_t npts, float leaf, int skip_nonfinite, uint32_t* npts_out);
/* the cloud currently set, 4 floats per point (x y z intensity; intensity 0 unless it came from the downsample) */
int amcl3d_b200_pf_get_cloud(amcl3d_b200_pf_t pf, float* xyzi, uint32_t cap, uint32_t* npts);

/* ParticleFilter::update(cloud) (ParticleFilter.cpp:114-139) on the cloud set above: w = 0 outside the
 * map, else computeCloudWeight.  With nb > 0 the range-only beacon weight (row R of SURVEY 8(a);
 * not present in the reference tree) is evaluated in the same kernel and fused as
 * w = alpha*wp/sum(wp) + (1-alpha)*wr/sum(wr).  in_bounds (optional) receives the number of
 * particle x point evaluations that loaded a voxel. */
int amcl3d_b200_pf_update(amcl3d_b200_pf_t pf, const amcl3d_b200_beacon_t* beacons, uint32_t nb, float alpha,
                          float sigma, uint64_t* in_bounds);
/* The same update with the fusion left to the caller: p_[i].w receives the raw cloud weight and d_wr_out (DEVICE,
 * n floats) the raw range weight.  Multi-GPU callers all-gather both, then call amcl3d_b200_pf_fuse_range on the
 * gathered set so that the two sums of the fusion are global. */
int amcl3d_b200_pf_update_split(amcl3d_b200_pf_t pf, const amcl3d_b200_beacon_t* beacons, uint32_t nb, float sigma,
                                void* d_wr_out, uint64_t* in_bounds);
/* w = alpha*w/sum(w) + (1-alpha)*wr/sum(wr) over the handle's particle set, d_wr = DEVICE array of n range weights */
int amcl3d_b200_pf_fuse_range(amcl3d_b200_pf_t pf, const void* d_wr, float alpha);
/* convenience: set_cloud + update */
int amcl3d_b200_pf_update_cloud(amcl3d_b200_pf_t pf, const float* cloud, uint32_t stride, uint32_t npts);

/* ParticleFilter::particleNormalize (ParticleFilter.cpp:168-196); wtp (optional) = the f32 sum */
int amcl3d_b200_pf_normalize(amcl3d_b200_pf_t pf, float* wtp);
/* MonteCarloLocalization::addParticleWeight (amcl3d.cpp:205-217): w *= 1/max(w) */
int amcl3d_b200_pf_scale_by_max(amcl3d_b200_pf_t pf);

/* ParticleFilter::resample (ParticleFilter.cpp:230-254) with r = (1/N)*u01.  Bit-exact indices:
 * the inclusive prefix is the reference's sequential f32 chain.  Outputs [m0, m1) only are produced
 * (m1 = 0 means N): d_out (optional, DEVICE, (m1-m0) particles) receives them, otherwise p_ is replaced
 * (requires m0 = 0, m1 = N).  idx (optional, HOST, m1-m0 ints) receives the source indices.  The
 * reference's read past the end (ParticleFilter.cpp:244-248) is clamped to N-1. */
int amcl3d_b200_pf_resample(amcl3d_b200_pf_t pf, float u01, uint64_t m0, uint64_t m1, void* d_out, int32_t* idx);

/* ParticleFilter::uniformSample (ParticleFilter.cpp:256-280): normalise, then S outputs at thresholds
 * 0.5/S + k/S (sequential f32 chain), first index with prefix > threshold.  Outputs [k0, k1) (k1 = 0
 * means S) go to d_out (DEVICE) or replace p_ (then size becomes S, requires k0 = 0, k1 = S); the tail
 * that the reference leaves as zero particles is zero here too, idx = -1 there.  emitted (optional). */
int amcl3d_b200_pf_uniform_sample(amcl3d_b200_pf_t pf, int sample_num, uint64_t k0, uint64_t k1, void* d_out,
                                  int32_t* idx, int* emitted);

/* ParticleFilter::getMean (ParticleFilter.cpp:198-216): exact != 0 replays the sequential f32 sums,
 * otherwise a deterministic f64 tree reduction of the f32 products.  out7 = x..yaw, w = max weight */
int amcl3d_b200_pf_mean(amcl3d_b200_pf_t pf, int exact, float out7[7]);
/* getMean with the 7 floats written to DEVICE memory of the caller (e.g. the send buffer of a collective);
 * no host wait in async mode */
int amcl3d_b200_pf_mean_device(amcl3d_b200_pf_t pf, int exact, void* d_out7);
/* ParticleFilter::getBestParticle (ParticleFilter.cpp:218-228): first arg-max with w > 0 */
int amcl3d_b200_pf_best(amcl3d_b200_pf_t pf, float out7[7]);

/* ParticleFilter::predict (ParticleFilter.cpp:91-112) with caller-supplied noise: 6 floats per
 * particle in draw order x,y,z,roll,pitch,yaw (what ranGaussian returned), HOST memory */
int amcl3d_b200_pf_predict(amcl3d_b200_pf_t pf, const double delta[6], const float* noise6);

/* ---- the direct callers of the path: MonteCarloLocalization (amcl3d/src/amcl3d.cpp) ---------- */
/* computeSampleNum (amcl3d.cpp:242-288, computeBoundary :222-241): occupied 0.2 m x yaw bins of the particle
 * set times cnt_per_grid, clamped to [min_resamp, max_resamp] */
int amcl3d_b200_pf_compute_sample_num(amcl3d_b200_pf_t pf, int cnt_per_grid, int min_resamp, int max_resamp,
                                      int* sample_num);
/* updateSampleHeight + getMapHeight (amcl3d.cpp:295-321): z += z(keypose nearest to (mean.x, mean.y, 0)) - mean.z;
 * keyposes_xyz = nkey x 3 HOST floats (Grid3d::keyposes_3d); delta_h optional */
int amcl3d_b200_pf_update_sample_height(amcl3d_b200_pf_t pf, const float* keyposes_xyz, uint64_t nkey, float* delta_h);
/* PFResample (amcl3d.cpp:192-203): [addParticleWeight] -> computeSampleNum(3) -> uniformSample -> updateSampleHeight
 * (skipped when no keyposes are given), particles never leave the device; new_size optional */
int amcl3d_b200_pf_resample_step(amcl3d_b200_pf_t pf, int weight_multiply, int min_resamp, int max_resamp,
                                 const float* keyposes_xyz, uint64_t nkey, int* new_size);

/* ---- particles sharded over the GPUs of one NVLink box ------------------------------------------
 * The grid is replicated (one grid handle per device); rank r of `world` owns the contiguous block
 * [lo, hi) of the global particle order (the first n_total % world ranks hold one particle more).  update() is local.
 * The exchange step of the global particleNormalize + resample (ParticleFilter.cpp:168-196, 230-254) runs in our own
 * kernels over peer memory: every rank stores its 4-byte weights into every peer's dense copy over NVLink (no 28-byte
 * records travel), replays the reference's sequential f32 chains over the full copy -- so weights, prefix and indices
 * are bit-identical to one GPU and to the CPU -- and loads the particles its output slice selects from their owners.
 * Ranks may be processes (one per GPU: exchange the 64-byte handles, e.g. with torch.distributed / MPI, then
 * attach_ipc) or handles of one process (attach_local, peer access).  The shard_* calls are collective: every rank
 * makes the same sequence of them.  A peer that never arrives is reported after ~4 s (ERR_STATE), not waited for. */
/* turns the handle into rank `rank` of `world`: p_ becomes this rank's block (size hi - lo) inside a freshly allocated
 * exchange region; fill it with set_particles / get it with get_particles as usual */
int amcl3d_b200_pf_shard_init(amcl3d_b200_pf_t pf, uint64_t n_total, int rank, int world);
int amcl3d_b200_pf_shard_info(amcl3d_b200_pf_t pf, uint64_t* lo, uint64_t* hi, uint64_t* region_bytes);
/* the same with room to grow: the region is sized for n_capacity particles in total (n_total <= n_capacity), so that steps
 * which change the particle count every frame (uniformSample, PFResample: amcl3d.cpp:192-203, ParticleFilter.cpp:256-280)
 * do not remake the region.  shard_resize (collective, every rank with the same value, no exchange in flight) re-draws the
 * block partition for a new total <= n_capacity; the contents of p_ are unspecified until it is filled again. */
int amcl3d_b200_pf_shard_init_capacity(amcl3d_b200_pf_t pf, uint64_t n_total, uint64_t n_capacity, int rank, int world);
int amcl3d_b200_pf_shard_resize(amcl3d_b200_pf_t pf, uint64_t n_total);
/* CUDA IPC handle (64 bytes) of this rank's region / open the regions of all ranks (`world` handles, `stride` bytes apart,
 * indexed by rank; the own entry is ignored) */
int amcl3d_b200_pf_shard_ipc_handle(amcl3d_b200_pf_t pf, void* handle64, size_t cap);
int amcl3d_b200_pf_shard_attach_ipc(amcl3d_b200_pf_t pf, const void* handles, size_t stride);
/* all ranks are handles of this process (pfs[r] = rank r): enables peer access between their devices */
int amcl3d_b200_pf_shard_attach_local(amcl3d_b200_pf_t* pfs, int world);
/* global particleNormalize (ParticleFilter.cpp:168-196): the weight sum runs over all ranks' particles, the rank's own
 * block is normalised in place (p_ keeps its particles).  wtp (optional, HOST) receives the global sum. */
int amcl3d_b200_pf_shard_normalize(amcl3d_b200_pf_t pf, float* wtp);
/* global resample(r = u01 / N) (ParticleFilter.cpp:230-254); p_ becomes this rank's slice [lo, hi) of the resampled set.
 * normalize != 0: particleNormalize first, inside the same exchange (the fused call of a frame); 0: the weights are
 * taken as they are (the caller ran shard_normalize, like `particleNormalize(); resample();` of the reference).
 * d_wr_local (optional, DEVICE, hi - lo floats, needs normalize != 0): raw range-beacon weights of update_split, fused
 * with alpha over the global sums before normalising (row R).  idx (optional, HOST, hi - lo ints): global source indices. */
int amcl3d_b200_pf_shard_resample(amcl3d_b200_pf_t pf, float u01, int normalize, const void* d_wr_local, float alpha, int32_t* idx);
/* Fast (NON-exact) variant of the fused normalise + resample, north_star's "allreduce of the weight sum + allgather of
 * prefix offsets": two scalar exchanges, every rank scans only its own block, and the source side stores each selected
 * particle straight into the output block of the rank that owns the slot -- no rank reads another rank's weights.  The
 * per-rank sums are combined in rank order instead of one sequential chain over all particles, so a threshold within an
 * ulp of a prefix value may select the neighbouring particle: indices are not guaranteed identical to one GPU
 * (tests/test_gpu_sharded.py reports the mismatch count).  No range-beacon fusion in this mode. */
int amcl3d_b200_pf_shard_resample_fast(amcl3d_b200_pf_t pf, float u01, int32_t* idx);
/* getMean over all ranks (partial sums exchanged through peer memory, combined in rank order in f64); out7 (HOST) and /
 * or d_out7 (DEVICE), at least one */
int amcl3d_b200_pf_shard_mean(amcl3d_b200_pf_t pf, float out7[7], void* d_out7);

/* the collective calls for ranks that are handles of ONE process (attach_local; ranks may share a device): the publishing
 * half of the exchange is issued on every handle, the streams are ordered with CUDA events, then the reading half is
 * issued -- no kernel spins on work of this process that is not launched yet (the per-handle calls above, whose wait_kernel
 * spins on the peers' flags, are for one process per rank; called rank after rank from one thread they would time out).
 * d_wr_local: NULL or `world` DEVICE pointers (rank r's raw range weights). */
int amcl3d_b200_pf_shard_normalize_local(amcl3d_b200_pf_t* pfs, int world);
int amcl3d_b200_pf_shard_resample_local(amcl3d_b200_pf_t* pfs, int world, float u01, int normalize, const void* const* d_wr_local,
                                        float alpha);
int amcl3d_b200_pf_shard_mean_local(amcl3d_b200_pf_t* pfs, int world, float out7[7]);
/* the same as ONE sequential f32 chain over the global order -- bit-identical to getMean of one GPU (ParticleFilter.cpp:198-216):
 * rank r continues the sums where rank r - 1 stopped (its seven floats read from peer memory).  Serial over the ranks. */
int amcl3d_b200_pf_shard_mean_exact_local(amcl3d_b200_pf_t* pfs, int world, float out7[7]);
/* Steps of the reference that have no sharded form (uniformSample, addParticleWeight, computeSampleNum, updateSampleHeight,
 * getBestParticle: their exact f32 chains run over the whole set anyway) on a set that lives in shards, without the host:
 * gather_local copies the blocks of all ranks into p_ of the plain handle `dst` (any device of the box) device to device
 * over NVLink, in stream order (dst's stream waits for the shards' streams and the other way round); scatter_local is
 * the way back: it resizes the shards to the count `src` holds now (<= their capacity, >= world) and copies every rank
 * its block of the new partition. */
int amcl3d_b200_pf_shard_gather_local(amcl3d_b200_pf_t* pfs, int world, amcl3d_b200_pf_t dst);
int amcl3d_b200_pf_shard_scatter_local(amcl3d_b200_pf_t src, amcl3d_b200_pf_t* pfs, int world);

/* sequential f32 inclusive prefix of the current weights (ParticleFilter.cpp:259-264), for tests and
 * for multi-GPU callers: DEVICE pointer valid until the next call on the handle */
int amcl3d_b200_pf_prefix(amcl3d_b200_pf_t pf, void** d_prefix, float* host_out, uint64_t cap);

/* device time in ms of the kernels of the last update / normalize / resample / uniform_sample /
 * mean call on this handle, measured with CUDA events on the handle's stream */
int amcl3d_b200_pf_last_kernel_ms(amcl3d_b200_pf_t pf, float* ms);
/* the large-set update of the last amcl3d_b200_pf_update* call split by kernel: ms[0] pose_kernel, ms[1] sweep_kernel
 * (the gather), ms[2] ordered_sum_kernel; zeros when the single-kernel path ran (last slice only when sliced) */
int amcl3d_b200_pf_last_phase_ms(amcl3d_b200_pf_t pf, float ms[3]);
/* kernels launched by this handle since creation */
int amcl3d_b200_pf_launch_count(amcl3d_b200_pf_t pf, uint64_t* launches);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* random-sector gather micro-benchmark used to state the gather roofline: every thread
 * of `blocks x 256` does `iters` x 8 independent 4-byte loads at hashed indices inside
 * [0, window_cells) of the grid's cells; returns device ms and the number of loads issued.
 * via_tex != 0 is reserved. */
int amcl3d_b200_bench_gather(amcl3d_b200_grid_t g, uint64_t window_cells, uint32_t blocks, uint32_t iters,
                             int coherent, float* ms, uint64_t* loads);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
