"""Python mirror of the reference's Grid3d / ParticleFilter interface over the C-ABI.

Names, argument meaning and error behaviour follow amcl3d/src/Grid3d.h and ParticleFilter.h
(open -> bool, computeCloudWeight -> 0 before a map is open, isIntoMap -> False without a map ...),
so the parity tests read like the reference's own tests.  Every method is one C-ABI call; arrays are
numpy float32 on the host or raw device pointers (ints) for resident data.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .capi import GRID_GAUSSIAN, GRID_QUIRK, GRID_SPLAT, TRIG_F32, TRIG_F64, check  # noqa: F401


def device_info(device: int = 0) -> dict:
    L = capi.load()
    cnt, sm, maj, mnr = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    l2, mem = C.c_uint64(), C.c_uint64()
    check(L.amcl3d_b200_device_info(device, C.byref(cnt), C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(l2), C.byref(mem)))
    return dict(device_count=cnt.value, sm_count=sm.value, cc=(maj.value, mnr.value), l2_bytes=l2.value, mem_bytes=mem.value)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Grid3d:
    """amcl3d::Grid3d (amcl3d/src/Grid3d.h:33-147)."""

    def __init__(self, device: int = 0):
        self._L = capi.load()
        self._h = C.c_void_p()
        check(self._L.amcl3d_b200_grid_create(device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.amcl3d_b200_grid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- map / grid set-up ------------------------------------------------------------------
    def info(self) -> dict:
        gi = capi.GridInfo()
        check(self._L.amcl3d_b200_grid_get_info(self._h, C.byref(gi)))
        return dict(min=np.array(list(gi.min)), max=np.array(list(gi.max)), resol=gi.resol, sensor_dev=gi.sensor_dev,
                    size=tuple(int(v) for v in gi.size), n_cells=int(gi.n_cells), has_map=bool(gi.has_map),
                    has_grid=bool(gi.has_grid), device=gi.device, rep=gi.rep, lut_n=int(gi.lut_n),
                    coded_bytes=int(gi.coded_bytes))

    def set_edt_fast(self, enable: bool):
        """A/B switch of the capped build passes (same float grid bit for bit either way)."""
        check(self._L.amcl3d_b200_grid_set_edt_fast(self._h, int(bool(enable))))

    def useCodes(self, enable: bool):
        """A/B switch for the lossless dictionary-coded copy (bit-identical results either way)."""
        check(self._L.amcl3d_b200_grid_use_codes(self._h, int(bool(enable))))

    def computePointCloud(self, xyz, resolution):
        """computePointCloud (PointCloudTools.cpp:84-109): bounds of the map cloud; raises on <= 1 point."""
        xyz = _f32(xyz)
        mn, mx = (C.c_double * 3)(), (C.c_double * 3)()
        rc = self._L.amcl3d_b200_grid_compute_bounds(self._h, xyz.ctypes.data, xyz.shape[1], xyz.shape[0], float(resolution), mn, mx)
        if rc == capi.ERR_EMPTY:
            raise RuntimeError("OcTree is empty")
        check(rc)
        return np.array(list(mn)), np.array(list(mx))

    def setMap(self, mn, mx, resol):
        a, b = (C.c_double * 3)(*[float(v) for v in mn]), (C.c_double * 3)(*[float(v) for v in mx])
        check(self._L.amcl3d_b200_grid_set_map(self._h, a, b, float(resol)))

    def computeGrid(self, xyz, sensor_dev, mode=GRID_QUIRK, sub=2, keep_edt=False):
        """computeGrid / computeGrid2 (PointCloudTools.cpp:111-254) on the device."""
        xyz = _f32(xyz)
        check(self._L.amcl3d_b200_grid_build(self._h, xyz.ctypes.data, xyz.shape[1], xyz.shape[0], float(sensor_dev), mode, sub, int(keep_edt)))

    def computeGridDevice(self, d_xyz: int, stride: int, n: int, sensor_dev, mode=GRID_QUIRK, sub=2, keep_edt=False):
        """computeGrid from map points resident in device memory (raw pointer)."""
        check(self._L.amcl3d_b200_grid_build_device(self._h, C.c_void_p(d_xyz), stride, n, float(sensor_dev), mode, sub, int(keep_edt)))

    def last_build_ms(self) -> float:
        ms = C.c_float(0)
        check(self._L.amcl3d_b200_grid_last_build_ms(self._h, C.byref(ms)))
        return ms.value

    def edt(self) -> np.ndarray:
        n = self.info()["n_cells"]
        out = np.empty(n, np.int32)
        check(self._L.amcl3d_b200_grid_download_edt(self._h, out.ctypes.data, n))
        return out

    def uploadGrid(self, prob, size, sensor_dev):
        prob = _f32(prob).reshape(-1)
        assert prob.size == int(size[0]) * int(size[1]) * int(size[2])
        check(self._L.amcl3d_b200_grid_upload(self._h, prob.ctypes.data, (C.c_uint32 * 3)(*[int(v) for v in size]), float(sensor_dev)))

    def downloadGrid(self) -> np.ndarray:
        n = self.info()["n_cells"]
        out = np.empty(n, np.float32)
        check(self._L.amcl3d_b200_grid_download(self._h, out.ctypes.data, n))
        return out

    def saveGrid(self, path) -> bool:
        return self._L.amcl3d_b200_grid_save_file(self._h, os.fspath(path).encode()) == capi.OK

    def loadGrid(self, path, sensor_dev) -> bool:
        return self._L.amcl3d_b200_grid_load_file(self._h, os.fspath(path).encode(), float(sensor_dev)) == capi.OK

    def open_points(self, xyz, sensor_dev, resolution, grid_path=None, mode=GRID_SPLAT, sub=2) -> bool:
        """The body of Grid3d::open after the PCDs are loaded (Grid3d.cpp:155-197): bounds, cached
        .grid if present and matching, else build (the fork builds with computeGrid2) and save."""
        try:
            self.computePointCloud(xyz, resolution)
        except Exception:
            return False
        if grid_path is not None and os.path.exists(grid_path) and self.loadGrid(grid_path, sensor_dev):
            return True
        self.computeGrid(xyz, sensor_dev, mode=mode, sub=sub)
        if grid_path is not None:
            self.saveGrid(grid_path)
        return True

    # ---- queries ----------------------------------------------------------------------------
    def isIntoMap(self, x, y, z) -> bool:
        out = C.c_int(0)
        check(self._L.amcl3d_b200_grid_is_into_map(self._h, float(np.float32(x)), float(np.float32(y)), float(np.float32(z)), C.byref(out)))
        return bool(out.value)

    def computeCloudWeight(self, cloud, tx, ty, tz, roll, pitch, yaw, trig_mode=TRIG_F64) -> np.float32:
        cloud = _f32(cloud)
        stride = cloud.shape[1] if cloud.ndim == 2 else 3
        npts = cloud.shape[0] if cloud.ndim == 2 else 0
        w = C.c_float(0)
        pose = [float(np.float32(v)) for v in (tx, ty, tz, roll, pitch, yaw)]
        check(self._L.amcl3d_b200_grid_cloud_weight(self._h, cloud.ctypes.data if npts else None, stride, npts, *pose, trig_mode, C.byref(w)))
        return np.float32(w.value)

    def clone(self, device: int, with_float=False) -> "Grid3d":
        """Replica of this grid on another device (peer copy of the coded cells; the float cells only on request)."""
        r = Grid3d.__new__(Grid3d)
        r._L = self._L
        r._h = C.c_void_p()
        check(self._L.amcl3d_b200_grid_clone(self._h, int(device), int(bool(with_float)), C.byref(r._h)))
        return r

    def device_ptr(self) -> int:
        p = C.c_void_p()
        check(self._L.amcl3d_b200_grid_device_ptr(self._h, C.byref(p)))
        return p.value or 0

    def bench_gather(self, window_cells=0, blocks=148 * 8, iters=64, coherent=False):
        ms, loads = C.c_float(0), C.c_uint64(0)
        check(self._L.amcl3d_b200_bench_gather(self._h, int(window_cells), blocks, iters, int(coherent) if not isinstance(coherent, bool) else int(coherent), C.byref(ms), C.byref(loads)))
        return ms.value, loads.value


class ParticleFilter:
    """amcl3d::ParticleFilter (amcl3d/src/ParticleFilter.h:104-252) bound to a Grid3d."""

    def __init__(self, grid3d: Grid3d):
        self._L = capi.load()
        self.grid3d_ = grid3d
        self._h = C.c_void_p()
        check(self._L.amcl3d_b200_pf_create(grid3d._h, C.byref(self._h)))
        self._keep = None  # keeps a bound torch tensor alive

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.amcl3d_b200_pf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- p_ -----------------------------------------------------------------------------------
    @property
    def p_(self) -> np.ndarray:
        n = self.size()
        out = np.empty((n, 7), np.float32)
        got = C.c_uint64(0)
        check(self._L.amcl3d_b200_pf_get_particles(self._h, out.ctypes.data, n, C.byref(got)))
        return out

    def read_particles(self, out: np.ndarray) -> int:
        """p_ into a caller-owned (n, 7) float32 array -- e.g. a view of pinned memory, which makes the device->host
        copy a single DMA instead of a staged pageable copy.  Returns the number of particles written."""
        assert out.dtype == np.float32 and out.ndim == 2 and out.shape[1] == 7 and out.flags["C_CONTIGUOUS"]
        got = C.c_uint64(0)
        check(self._L.amcl3d_b200_pf_get_particles(self._h, out.ctypes.data, out.shape[0], C.byref(got)))
        return int(got.value)

    @p_.setter
    def p_(self, aos7):
        aos7 = _f32(aos7)
        assert aos7.ndim == 2 and aos7.shape[1] == 7
        check(self._L.amcl3d_b200_pf_set_particles(self._h, aos7.ctypes.data, aos7.shape[0]))

    def size(self) -> int:
        n = C.c_uint64(0)
        check(self._L.amcl3d_b200_pf_size(self._h, C.byref(n)))
        return int(n.value)

    def bind_device(self, ptr: int, n: int, capacity: int, keep=None):
        check(self._L.amcl3d_b200_pf_bind_device(self._h, C.c_void_p(ptr), n, capacity))
        self._keep = keep

    def device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64(0)
        check(self._L.amcl3d_b200_pf_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value or 0, int(n.value)

    def set_stream(self, cuda_stream: int | None):
        """Run this handle's kernels on a caller-owned stream.  None = the handle's own stream; 0 (torch's default
        stream is the legacy NULL stream) is passed as cudaStreamLegacy so that work really is ordered with it."""
        if cuda_stream is None:
            handle = 0
        elif cuda_stream == 0:
            handle = 1  # cudaStreamLegacy
        else:
            handle = cuda_stream
        check(self._L.amcl3d_b200_pf_set_stream(self._h, C.c_void_p(handle)))

    def set_two_phase(self, enable: bool):
        """A/B switch of the Morton-sweep + ordered-sum update (bit-identical results either way)."""
        check(self._L.amcl3d_b200_pf_set_two_phase(self._h, int(bool(enable))))

    def set_trig_mode(self, mode):
        check(self._L.amcl3d_b200_pf_set_trig_mode(self._h, mode))

    def set_chain_mode(self, mode: int):
        """0: exact parallel binade scan (default), 1: serial chain kernel, 2: one-CTA scan.  Same bits either way."""
        check(self._L.amcl3d_b200_pf_set_chain_mode(self._h, int(mode)))

    # ---- the path -----------------------------------------------------------------------------
    def set_cloud(self, cloud):
        cloud = _f32(cloud)
        npts = cloud.shape[0]
        check(self._L.amcl3d_b200_pf_set_cloud(self._h, cloud.ctypes.data if npts else None, cloud.shape[1] if cloud.ndim == 2 else 3, npts))

    def set_cloud_downsampled(self, cloud, leaf, intensity_offset=None, skip_nonfinite=False) -> int:
        """ds_.filter() of amcl3d.cpp:51-52 (pcl::VoxelGrid centroid filter, leaf = voxel_size) on the device; the
        result becomes the cloud of the next update().  Returns the number of points now set."""
        cloud = _f32(cloud)
        npts, stride = cloud.shape[0], cloud.shape[1]
        if intensity_offset is None:
            intensity_offset = 4 if stride >= 8 else (3 if stride >= 4 else -1)
        out = C.c_uint32(0)
        check(self._L.amcl3d_b200_pf_set_cloud_downsampled(self._h, cloud.ctypes.data if npts else None, stride,
                                                           int(intensity_offset), npts, float(leaf), int(skip_nonfinite),
                                                           C.byref(out)))
        return int(out.value)

    def get_cloud(self) -> np.ndarray:
        """The cloud currently set, (npts, 4) = x y z intensity."""
        n = C.c_uint32(0)
        check(self._L.amcl3d_b200_pf_get_cloud(self._h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), np.float32)
        if n.value:
            check(self._L.amcl3d_b200_pf_get_cloud(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def set_cloud_device(self, ptr: int, stride: int, npts: int):
        check(self._L.amcl3d_b200_pf_set_cloud_device(self._h, C.c_void_p(ptr), stride, npts))

    def update(self, cloud=None, beacons=None, alpha=1.0, sigma=1.0, count=False):
        """ParticleFilter::update(cloud) (ParticleFilter.cpp:114-139).  cloud=None reuses the resident one."""
        if cloud is not None:
            self.set_cloud(cloud)
        nb = 0
        arr = None
        if beacons is not None:
            pos, rng = beacons
            nb = len(rng)
            arr = (capi.Beacon * nb)()
            for k in range(nb):
                for a in range(3):
                    arr[k].position[a] = float(pos[k][a])
                arr[k].range = float(rng[k])
        c = C.c_uint64(0)
        check(self._L.amcl3d_b200_pf_update(self._h, arr, nb, float(alpha), float(sigma), C.byref(c) if count else None))
        return int(c.value) if count else None

    @staticmethod
    def _beacon_array(beacons):
        pos, rng = beacons
        arr = (capi.Beacon * len(rng))()
        for k in range(len(rng)):
            for a in range(3):
                arr[k].position[a] = float(pos[k][a])
            arr[k].range = float(rng[k])
        return arr, len(rng)

    def update_split(self, beacons, sigma, d_wr_out: int):
        """update() with the range fusion left to the caller: raw cloud weight in p_, raw range weight in d_wr_out."""
        arr, nb = self._beacon_array(beacons)
        check(self._L.amcl3d_b200_pf_update_split(self._h, arr, nb, float(sigma), C.c_void_p(d_wr_out), None))

    def fuse_range(self, d_wr: int, alpha):
        check(self._L.amcl3d_b200_pf_fuse_range(self._h, C.c_void_p(d_wr), float(alpha)))

    def particleNormalize(self, fetch=True):
        """ParticleFilter::particleNormalize; returns wtp (fetch=False: nothing is read back, no host wait in async mode)."""
        if not fetch:
            check(self._L.amcl3d_b200_pf_normalize(self._h, None))
            return None
        wtp = C.c_float(0)
        check(self._L.amcl3d_b200_pf_normalize(self._h, C.byref(wtp)))
        return np.float32(wtp.value)

    def set_async(self, enable: bool):
        """Stream order instead of call order for the calls that return nothing to the host (see the C header)."""
        check(self._L.amcl3d_b200_pf_set_async(self._h, int(bool(enable))))

    def sync(self):
        check(self._L.amcl3d_b200_pf_sync(self._h))

    def addParticleWeight(self):
        check(self._L.amcl3d_b200_pf_scale_by_max(self._h))

    def resample(self, u01, m0=0, m1=0, d_out=0, want_idx=False):
        """ParticleFilter::resample (ParticleFilter.cpp:230-254) with r = u01 / N."""
        n = self.size()
        cnt = (m1 or n) - m0
        idx = np.empty(cnt, np.int32) if want_idx else None
        check(self._L.amcl3d_b200_pf_resample(self._h, float(np.float32(u01)), m0, m1, C.c_void_p(d_out or 0), idx.ctypes.data if want_idx else None))
        return idx

    def uniformSample(self, sample_num, k0=0, k1=0, d_out=0, want_idx=False):
        """ParticleFilter::uniformSample (ParticleFilter.cpp:256-280).  Returns (idx | None, emitted)."""
        cnt = (k1 or sample_num) - k0
        idx = np.empty(cnt, np.int32) if want_idx else None
        em = C.c_int(0)
        check(self._L.amcl3d_b200_pf_uniform_sample(self._h, int(sample_num), k0, k1, C.c_void_p(d_out or 0), idx.ctypes.data if want_idx else None, C.byref(em)))
        return idx, em.value

    def getMean(self, exact=False) -> np.ndarray:
        out = np.zeros(7, np.float32)
        check(self._L.amcl3d_b200_pf_mean(self._h, int(exact), out.ctypes.data))
        return out

    def mean_device(self, d_out7: int, exact=False):
        """getMean into 7 floats of caller-owned device memory (no host wait in async mode)."""
        check(self._L.amcl3d_b200_pf_mean_device(self._h, int(exact), C.c_void_p(d_out7)))

    def getBestParticle(self) -> np.ndarray:
        out = np.zeros(7, np.float32)
        check(self._L.amcl3d_b200_pf_best(self._h, out.ctypes.data))
        return out

    def predict(self, delta6, noise6):
        d = np.ascontiguousarray(delta6, np.float64)
        nz = _f32(noise6)
        check(self._L.amcl3d_b200_pf_predict(self._h, d.ctypes.data_as(C.POINTER(C.c_double)), nz.ctypes.data))

    # ---- MonteCarloLocalization steps around the path (amcl3d.cpp) ------------------------------------------
    def computeSampleNum(self, cnt_per_grid=3, min_resamp=150, max_resamp=800) -> int:
        out = C.c_int(0)
        check(self._L.amcl3d_b200_pf_compute_sample_num(self._h, cnt_per_grid, min_resamp, max_resamp, C.byref(out)))
        return out.value

    def updateSampleHeight(self, keyposes) -> np.float32:
        kp = _f32(keyposes)
        dh = C.c_float(0)
        check(self._L.amcl3d_b200_pf_update_sample_height(self._h, kp.ctypes.data, kp.shape[0], C.byref(dh)))
        return np.float32(dh.value)

    def PFResample(self, weight_multiply, min_resamp, max_resamp, keyposes) -> int:
        kp = _f32(keyposes)
        out = C.c_int(0)
        check(self._L.amcl3d_b200_pf_resample_step(self._h, int(weight_multiply), min_resamp, max_resamp, kp.ctypes.data, kp.shape[0], C.byref(out)))
        return out.value

    def prefix(self) -> np.ndarray:
        n = self.size()
        out = np.empty(n, np.float32)
        check(self._L.amcl3d_b200_pf_prefix(self._h, None, out.ctypes.data, n))
        return out

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        check(self._L.amcl3d_b200_pf_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def last_phase_ms(self):
        """(pose, sweep, ordered sum) ms of the last large-set update; zeros when the single-kernel path ran."""
        ms = (C.c_float * 3)()
        check(self._L.amcl3d_b200_pf_last_phase_ms(self._h, ms))
        return tuple(float(v) for v in ms)

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        check(self._L.amcl3d_b200_pf_launch_count(self._h, C.byref(n)))
        return int(n.value)

    # ---- particles sharded over the GPUs of one box, exchange over peer memory (csrc/shard.cu) ----------------
    def shard_init(self, n_total: int, rank: int, world: int):
        """This handle becomes rank `rank` of `world`; p_ is its contiguous block [lo, hi) of the global order."""
        check(self._L.amcl3d_b200_pf_shard_init(self._h, int(n_total), int(rank), int(world)))
        return self.shard_info()[:2]

    def shard_init_capacity(self, n_total: int, n_capacity: int, rank: int, world: int):
        """shard_init with an exchange region sized for n_capacity particles in total: the set may be resized up to that."""
        check(self._L.amcl3d_b200_pf_shard_init_capacity(self._h, int(n_total), int(n_capacity), int(rank), int(world)))
        return self.shard_info()[:2]

    def shard_resize(self, n_total: int):
        """Collective: new total particle count (<= capacity), the block partition is re-drawn; p_ is to be filled again."""
        check(self._L.amcl3d_b200_pf_shard_resize(self._h, int(n_total)))
        return self.shard_info()[:2]

    @staticmethod
    def shard_gather_local(pfs, dst):
        """The blocks of all ranks (handles of this process) into p_ of the plain handle dst, device to device."""
        arr = (C.c_void_p * len(pfs))(*[p._h.value for p in pfs])
        check(capi.load().amcl3d_b200_pf_shard_gather_local(arr, len(pfs), dst._h))

    @staticmethod
    def shard_scatter_local(src, pfs):
        """p_ of the plain handle src over the ranks (resized to its count), device to device."""
        arr = (C.c_void_p * len(pfs))(*[p._h.value for p in pfs])
        check(capi.load().amcl3d_b200_pf_shard_scatter_local(src._h, arr, len(pfs)))

    def shard_info(self):
        lo, hi, nb = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(self._L.amcl3d_b200_pf_shard_info(self._h, C.byref(lo), C.byref(hi), C.byref(nb)))
        return int(lo.value), int(hi.value), int(nb.value)

    def shard_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(self._L.amcl3d_b200_pf_shard_ipc_handle(self._h, buf, 64))
        return buf.raw

    def shard_attach_ipc(self, handles):
        """handles: the 64-byte handles of all ranks, indexed by rank (one process per GPU)."""
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) % 64 == 0
        check(self._L.amcl3d_b200_pf_shard_attach_ipc(self._h, blob, 64))

    @staticmethod
    def shard_attach_local(pfs):
        """All ranks are handles of this process (pfs[r] = rank r), one per device."""
        arr = (C.c_void_p * len(pfs))(*[p._h.value for p in pfs])
        check(capi.load().amcl3d_b200_pf_shard_attach_local(arr, len(pfs)))

    @staticmethod
    def shard_resample_local(pfs, u01, normalize=True, d_wr_local=None, alpha=1.0):
        """The collective resample for ranks that are handles of one process: launched on all, then waited for."""
        arr = (C.c_void_p * len(pfs))(*[p._h.value for p in pfs])
        wr = (C.c_void_p * len(pfs))(*[int(v) for v in d_wr_local]) if d_wr_local is not None else None
        check(capi.load().amcl3d_b200_pf_shard_resample_local(arr, len(pfs), float(np.float32(u01)), int(bool(normalize)), wr,
                                                              float(alpha)))

    @staticmethod
    def shard_normalize_local(pfs):
        arr = (C.c_void_p * len(pfs))(*[p._h.value for p in pfs])
        check(capi.load().amcl3d_b200_pf_shard_normalize_local(arr, len(pfs)))

    def shard_normalize(self, fetch=True):
        """Global particleNormalize: the sum runs over all ranks, this rank's block is normalised in place."""
        if not fetch:
            check(self._L.amcl3d_b200_pf_shard_normalize(self._h, None))
            return None
        wtp = C.c_float(0)
        check(self._L.amcl3d_b200_pf_shard_normalize(self._h, C.byref(wtp)))
        return np.float32(wtp.value)

    @staticmethod
    def shard_mean_exact_local(pfs) -> np.ndarray:
        """getMean as one sequential f32 chain carried from rank to rank: the bits one GPU produces."""
        arr = (C.c_void_p * len(pfs))(*[p._h.value for p in pfs])
        out = np.zeros(7, np.float32)
        check(capi.load().amcl3d_b200_pf_shard_mean_exact_local(arr, len(pfs), out.ctypes.data))
        return out

    @staticmethod
    def shard_mean_local(pfs) -> np.ndarray:
        arr = (C.c_void_p * len(pfs))(*[p._h.value for p in pfs])
        out = np.zeros(7, np.float32)
        check(capi.load().amcl3d_b200_pf_shard_mean_local(arr, len(pfs), out.ctypes.data))
        return out

    def shard_resample(self, u01, d_wr_local: int = 0, alpha=1.0, want_idx=False, normalize=True):
        """Global [particleNormalize +] resample over all ranks; p_ becomes this rank's slice of the resampled set."""
        idx = np.empty(self.size(), np.int32) if want_idx else None
        check(self._L.amcl3d_b200_pf_shard_resample(self._h, float(np.float32(u01)), int(bool(normalize)), C.c_void_p(d_wr_local or 0),
                                                    float(alpha), idx.ctypes.data if want_idx else None))
        return idx

    def shard_resample_fast(self, u01, want_idx=False):
        """The NON-exact sharded normalise + resample (scalar exchanges + source-side placement); see the C header."""
        idx = np.empty(self.size(), np.int32) if want_idx else None
        check(self._L.amcl3d_b200_pf_shard_resample_fast(self._h, float(np.float32(u01)), idx.ctypes.data if want_idx else None))
        return idx

    def shard_mean(self, d_out7: int = 0, fetch=True):
        """getMean over all ranks; fetch=False leaves the 7 floats in d_out7 (device) without a host wait in async mode."""
        out = np.zeros(7, np.float32) if fetch else None
        check(self._L.amcl3d_b200_pf_shard_mean(self._h, out.ctypes.data if fetch else None, C.c_void_p(d_out7 or 0)))
        return out
