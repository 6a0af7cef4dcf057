"""ctypes loaders for the CPU oracle (oracle/amcl3d_oracle.c) and the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, nowhere else.  Nothing in amcl3d_test_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "_build" / "libamcl3d_oracle.so"
REF_SO = HERE / "_ref" / "libamcl3d_ref.so"
REF_SRC = Path("/root/reference/amcl3d/src")

TRIG_F64, TRIG_F32 = 0, 1
GRID_QUIRK, GRID_GAUSSIAN, GRID_SPLAT = 0, 1, 2
EDT_INF = np.int32(2**31 - 1)

particle_dtype = np.dtype([(k, "<f4") for k in ("x", "y", "z", "roll", "pitch", "yaw", "w")])


class OrcMap(C.Structure):
    _fields_ = [
        ("min", C.c_double * 3),
        ("max", C.c_double * 3),
        ("resol", C.c_double),
        ("sensor_dev", C.c_double),
        ("size", C.c_uint32 * 3),
        ("step_y", C.c_uint32),
        ("step_z", C.c_uint32),
        ("n_cells", C.c_uint64),
        ("prob", C.POINTER(C.c_float)),
        ("has_pc", C.c_int),
    ]


class OrcBeacon(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("range", C.c_double)]


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    targets = ["oracle"]
    if ref and REF_SRC.exists():
        targets.append("ref")
    subprocess.run(["make", "-C", str(HERE), *targets], check=True, stdout=subprocess.DEVNULL)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _as_f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        assert a.ndim == 2 and a.shape[1] >= cols, a.shape
    return a


class Map:
    """Geometry + cells of one likelihood field, kept alive for the C side."""

    def __init__(self, mn, mx, resol, sensor_dev, prob=None, size=None, has_pc=True):
        self.c = OrcMap()
        for a in range(3):
            self.c.min[a] = float(mn[a])
            self.c.max[a] = float(mx[a])
        self.c.resol = float(resol)
        self.c.sensor_dev = float(sensor_dev)
        self.c.has_pc = 1 if has_pc else 0
        lib().orc_grid_geometry(C.byref(self.c))
        if size is not None:  # geometry read from a .grid header instead of recomputed
            for a in range(3):
                self.c.size[a] = int(size[a])
            self.c.step_y = np.uint32(size[0])
            self.c.step_z = np.uint32((int(size[0]) * int(size[1])) & 0xFFFFFFFF)
            self.c.n_cells = (int(size[0]) * int(size[1]) * int(size[2])) & 0xFFFFFFFF
        self.prob = None
        if prob is not None:
            self.set_prob(prob)

    def set_prob(self, prob):
        self.prob = np.ascontiguousarray(prob, dtype=np.float32).reshape(-1)
        assert self.prob.size == self.n_cells, (self.prob.size, self.n_cells)
        self.c.prob = _fp(self.prob)

    @property
    def size(self):
        return tuple(int(v) for v in self.c.size)

    @property
    def n_cells(self):
        return int(self.c.n_cells)

    @property
    def mn(self):
        return np.array(list(self.c.min))

    @property
    def mx(self):
        return np.array(list(self.c.max))

    @property
    def resol(self):
        return float(self.c.resol)

    @property
    def sensor_dev(self):
        return float(self.c.sensor_dev)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not ORACLE_SO.exists():
        build(ref=False)
    L = C.CDLL(str(ORACLE_SO))
    P = C.POINTER
    mp = P(OrcMap)
    f32p, i32p, f64p, vp = P(C.c_float), P(C.c_int32), P(C.c_double), C.c_void_p
    sig = {
        "orc_num_threads": (C.c_int, []),
        "orc_compute_bounds": (C.c_int, [f32p, C.c_uint32, C.c_uint64, f64p, f64p]),
        "orc_grid_geometry": (None, [mp]),
        "orc_is_into_map": (C.c_int, [mp, C.c_float, C.c_float, C.c_float]),
        "orc_cloud_weight": (C.c_float, [mp, f32p, C.c_uint32, C.c_uint64] + [C.c_float] * 6 + [C.c_int, P(C.c_int)]),
        "orc_update": (None, [mp, vp, C.c_uint64, f32p, C.c_uint32, C.c_uint64, C.c_int, C.c_int]),
        "orc_count_in_bounds": (C.c_uint64, [mp, vp, C.c_uint64, f32p, C.c_uint32, C.c_uint64, C.c_int, C.c_int]),
        "orc_normalize": (C.c_float, [mp, vp, C.c_uint64]),
        "orc_scale_by_max": (None, [vp, C.c_uint64]),
        "orc_resample": (None, [vp, C.c_uint64, C.c_float, vp, i32p]),
        "orc_uniform_sample": (C.c_int, [mp, vp, C.c_uint64, C.c_int, vp, i32p]),
        "orc_mean": (None, [vp, C.c_uint64, vp]),
        "orc_best": (None, [vp, C.c_uint64, vp]),
        "orc_prefix": (None, [vp, C.c_uint64, f32p]),
        "orc_compute_sample_num": (C.c_int, [vp, C.c_uint64, C.c_int, C.c_int, C.c_int]),
        "orc_update_sample_height": (C.c_float, [vp, C.c_uint64, f32p, C.c_uint64]),
        "orc_voxel_filter": (C.c_int64, [f32p, C.c_uint32, C.c_int, C.c_uint64, C.c_float, C.c_int, f32p]),
        "orc_pf_resample_step": (C.c_int, [mp, vp, C.c_uint64, C.c_int, C.c_int, C.c_int, f32p, C.c_uint64, vp]),
        "orc_reloc_pose": (None, [vp, C.c_uint64, f32p, f32p, f32p, vp]),
        "orc_predict": (None, [vp, C.c_uint64, f64p, f32p]),
        "orc_compute_grid_nn": (C.c_int, [mp, f32p, C.c_uint32, C.c_uint64, C.c_int, f32p, C.c_int]),
        "orc_compute_grid2": (C.c_int, [mp, f32p, C.c_uint32, C.c_uint64, f32p]),
        "orc_edt_exact": (C.c_int, [mp, f32p, C.c_uint32, C.c_uint64, C.c_int, i32p, C.c_int]),
        "orc_edt_brute": (C.c_int, [mp, f32p, C.c_uint32, C.c_uint64, C.c_int, i32p]),
        "orc_grid_from_edt": (None, [mp, i32p, C.c_int, C.c_int, f32p]),
        "orc_range_weight": (C.c_float, [vp, P(OrcBeacon), C.c_uint32, C.c_float]),
        "orc_fuse_range": (None, [vp, C.c_uint64, f32p, C.c_float]),
        "orc_save_grid": (C.c_int, [C.c_char_p, mp]),
        "orc_load_grid": (C.c_int, [C.c_char_p, C.c_double, P(C.c_uint32), f64p, f32p, C.c_uint64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def num_threads() -> int:
    return int(lib().orc_num_threads())


def particles(a) -> np.ndarray:
    """(n,7) float32 C-contiguous copy."""
    a = np.array(a, dtype=np.float32, order="C", copy=True)
    assert a.ndim == 2 and a.shape[1] == 7
    return a


def compute_bounds(xyz):
    xyz = _as_f32(xyz, 3)
    mn = (C.c_double * 3)()
    mx = (C.c_double * 3)()
    rc = lib().orc_compute_bounds(_fp(xyz), xyz.shape[1], xyz.shape[0], mn, mx)
    if rc != 0:
        raise RuntimeError("OcTree is empty")
    return np.array(list(mn)), np.array(list(mx))


def is_into_map(m: Map, x, y, z) -> bool:
    return bool(lib().orc_is_into_map(C.byref(m.c), x, y, z))


def cloud_weight(m: Map, cloud, pose, trig=TRIG_F64, return_n=False):
    cloud = _as_f32(cloud, 3)
    n_in = C.c_int(0)
    pose = [float(np.float32(v)) for v in pose]
    w = lib().orc_cloud_weight(C.byref(m.c), _fp(cloud), cloud.shape[1], cloud.shape[0], *pose, trig, C.byref(n_in))
    return (np.float32(w), n_in.value) if return_n else np.float32(w)


def update(m: Map, p, cloud, trig=TRIG_F64, threads=1):
    """In place on p (n,7) float32."""
    cloud = _as_f32(cloud, 3)
    assert p.dtype == np.float32 and p.flags.c_contiguous
    lib().orc_update(C.byref(m.c), p.ctypes.data, p.shape[0], _fp(cloud), cloud.shape[1], cloud.shape[0], trig, threads)
    return p


def count_in_bounds(m: Map, p, cloud, trig=TRIG_F64, threads=0) -> int:
    cloud = _as_f32(cloud, 3)
    threads = threads or num_threads()
    return int(
        lib().orc_count_in_bounds(C.byref(m.c), p.ctypes.data, p.shape[0], _fp(cloud), cloud.shape[1], cloud.shape[0], trig, threads)
    )


def normalize(m: Map, p) -> np.float32:
    return np.float32(lib().orc_normalize(C.byref(m.c), p.ctypes.data, p.shape[0]))


def scale_by_max(p):
    lib().orc_scale_by_max(p.ctypes.data, p.shape[0])
    return p


def resample(p, u01):
    n = p.shape[0]
    out = np.zeros((n, 7), np.float32)
    idx = np.zeros(n, np.int32)
    lib().orc_resample(p.ctypes.data, n, float(np.float32(u01)), out.ctypes.data, idx.ctypes.data_as(C.POINTER(C.c_int32)))
    return out, idx


def uniform_sample(m: Map, p, S):
    """Normalises p in place (like the reference), returns (out (S,7), idx (S,), emitted)."""
    out = np.zeros((max(S, 0), 7), np.float32)
    idx = np.zeros(max(S, 0), np.int32)
    k = lib().orc_uniform_sample(C.byref(m.c), p.ctypes.data, p.shape[0], S, out.ctypes.data, idx.ctypes.data_as(C.POINTER(C.c_int32)))
    return out, idx, int(k)


def mean(p):
    out = np.zeros(7, np.float32)
    lib().orc_mean(p.ctypes.data, p.shape[0], out.ctypes.data)
    return out


def best(p):
    out = np.zeros(7, np.float32)
    lib().orc_best(p.ctypes.data, p.shape[0], out.ctypes.data)
    return out


def prefix(p):
    out = np.zeros(p.shape[0], np.float32)
    lib().orc_prefix(p.ctypes.data, p.shape[0], _fp(out))
    return out


def compute_sample_num(p, cnt_per_grid=3, min_resamp=150, max_resamp=800) -> int:
    return int(lib().orc_compute_sample_num(p.ctypes.data, p.shape[0], cnt_per_grid, min_resamp, max_resamp))


def update_sample_height(p, keyposes) -> np.float32:
    kp = _as_f32(keyposes, 3)
    return np.float32(lib().orc_update_sample_height(p.ctypes.data, p.shape[0], _fp(kp), kp.shape[0]))


def pf_resample_step(m: Map, p, weight_multiply, min_resamp, max_resamp, keyposes):
    kp = _as_f32(keyposes, 3)
    out = np.zeros((max(max_resamp, 1), 7), np.float32)
    n = lib().orc_pf_resample_step(C.byref(m.c), p.ctypes.data, p.shape[0], int(weight_multiply), min_resamp, max_resamp, _fp(kp), kp.shape[0], out.ctypes.data)
    return out[:n].copy()


def voxel_filter(pts, leaf, ioff=3, skip_nonfinite=False):
    """pcl::VoxelGrid centroid filter (amcl3d.cpp:37,51-52); pts (n, stride>=3) f32 -> (cells, 4) or None when
    PCL's leaf-too-small guard passes the input through."""
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros((max(pts.shape[0], 1), 4), np.float32)
    k = lib().orc_voxel_filter(_fp(pts), pts.shape[1], int(ioff), pts.shape[0], float(leaf), int(skip_nonfinite), _fp(out))
    return None if k < 0 else out[:k].copy()


def reloc_pose(n, init, dev, noise):
    p = np.zeros((n, 7), np.float32)
    init = np.ascontiguousarray(init, np.float32)
    dev = np.ascontiguousarray(dev, np.float32)
    noise = np.ascontiguousarray(noise, np.float32)
    m = np.zeros(7, np.float32)
    lib().orc_reloc_pose(p.ctypes.data, n, _fp(init), _fp(dev), _fp(noise), m.ctypes.data)
    return p, m


def predict(p, delta, noise):
    delta = np.ascontiguousarray(delta, np.float64)
    noise = np.ascontiguousarray(noise, np.float32)
    lib().orc_predict(p.ctypes.data, p.shape[0], delta.ctypes.data_as(C.POINTER(C.c_double)), _fp(noise))
    return p


def compute_grid_nn(m: Map, xyz, mode=GRID_QUIRK, threads=0):
    xyz = _as_f32(xyz, 3)
    prob = np.zeros(m.n_cells, np.float32)
    rc = lib().orc_compute_grid_nn(C.byref(m.c), _fp(xyz), xyz.shape[1], xyz.shape[0], mode, _fp(prob), threads or num_threads())
    assert rc == 0
    return prob


def compute_grid2(m: Map, xyz):
    xyz = _as_f32(xyz, 3)
    prob = np.zeros(m.n_cells, np.float32)
    rc = lib().orc_compute_grid2(C.byref(m.c), _fp(xyz), xyz.shape[1], xyz.shape[0], _fp(prob))
    assert rc == 0
    return prob


def edt_exact(m: Map, xyz, sub=2, threads=0):
    xyz = _as_f32(xyz, 3)
    sx, sy, sz = m.size
    d2 = np.zeros(sx * sy * sz, np.int32)
    rc = lib().orc_edt_exact(C.byref(m.c), _fp(xyz), xyz.shape[1], xyz.shape[0], sub, d2.ctypes.data_as(C.POINTER(C.c_int32)), threads or num_threads())
    assert rc == 0
    return d2


def edt_brute(m: Map, xyz, sub=2):
    xyz = _as_f32(xyz, 3)
    sx, sy, sz = m.size
    d2 = np.zeros(sx * sy * sz, np.int32)
    rc = lib().orc_edt_brute(C.byref(m.c), _fp(xyz), xyz.shape[1], xyz.shape[0], sub, d2.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return d2


def grid_from_edt(m: Map, d2, sub=2, mode=GRID_QUIRK):
    d2 = np.ascontiguousarray(d2, np.int32)
    prob = np.zeros(d2.size, np.float32)
    lib().orc_grid_from_edt(C.byref(m.c), d2.ctypes.data_as(C.POINTER(C.c_int32)), sub, mode, _fp(prob))
    return prob


def make_beacons(pos, rng):
    arr = (OrcBeacon * len(rng))()
    for i in range(len(rng)):
        for a in range(3):
            arr[i].position[a] = float(pos[i][a])
        arr[i].range = float(rng[i])
    return arr


def range_weights(p, pos, rng, sigma):
    b = make_beacons(pos, rng)
    out = np.zeros(p.shape[0], np.float32)
    for i in range(p.shape[0]):
        out[i] = lib().orc_range_weight(p[i].ctypes.data, b, len(rng), float(np.float32(sigma)))
    return out


def fuse_range(p, wr, alpha):
    wr = np.ascontiguousarray(wr, np.float32)
    lib().orc_fuse_range(p.ctypes.data, p.shape[0], _fp(wr), float(np.float32(alpha)))
    return p


def save_grid(path, m: Map) -> int:
    return int(lib().orc_save_grid(str(path).encode(), C.byref(m.c)))


def load_grid(path, sensor_dev):
    size = (C.c_uint32 * 3)()
    dev = C.c_double(0)
    rc = lib().orc_load_grid(str(path).encode(), float(sensor_dev), size, C.byref(dev), None, 0)
    if rc != 0:
        return rc, None, None, dev.value
    cells = (int(size[0]) * int(size[1]) * int(size[2])) & 0xFFFFFFFF
    prob = np.zeros(cells, np.float32)
    rc = lib().orc_load_grid(str(path).encode(), float(sensor_dev), size, C.byref(dev), _fp(prob), cells)
    return rc, tuple(int(v) for v in size), prob, dev.value


# --------------------------------------------------------------------------------------------------
# compiled reference (oracle/_ref)
# --------------------------------------------------------------------------------------------------
_ref = None


def ref_available() -> bool:
    return REF_SO.exists()


def ref():
    global _ref
    if _ref is not None:
        return _ref
    if not REF_SO.exists():
        if REF_SRC.exists():
            build(ref=True)
        else:
            raise FileNotFoundError("oracle/_ref/libamcl3d_ref.so missing and /root/reference absent")
    L = C.CDLL(str(REF_SO))
    P = C.POINTER
    f32p, f64p, u32p, vp = P(C.c_float), P(C.c_double), P(C.c_uint32), C.c_void_p
    sig = {
        "ref_grid_create": (vp, []),
        "ref_grid_destroy": (None, [vp]),
        "ref_grid_install": (None, [vp, f64p, f64p, C.c_double, f32p, u32p, C.c_double]),
        "ref_grid_cloud_weight": (C.c_float, [vp, f32p, C.c_uint32, C.c_uint64] + [C.c_float] * 6),
        "ref_grid_is_into_map": (C.c_int, [vp, C.c_float, C.c_float, C.c_float]),
        "ref_grid_open": (C.c_int, [vp, C.c_char_p, C.c_double]),
        "ref_grid_save": (C.c_int, [vp, C.c_char_p]),
        "ref_grid_load": (C.c_int, [vp, C.c_char_p, C.c_double]),
        "ref_grid_get": (C.c_int, [vp, f64p, f64p, f64p, u32p, f64p, f32p, C.c_uint64]),
        "ref_compute_grid": (C.c_int64, [f32p, C.c_uint32, C.c_uint64, C.c_double, C.c_double, C.c_int, C.c_int, f64p, f64p, u32p, f32p, C.c_uint64]),
        "ref_pf_create": (vp, [vp]),
        "ref_pf_destroy": (None, [vp]),
        "ref_pf_seed": (None, [vp, C.c_uint32]),
        "ref_pf_set": (None, [vp, f32p, C.c_uint64]),
        "ref_pf_size": (C.c_uint64, [vp]),
        "ref_pf_get": (None, [vp, f32p]),
        "ref_pf_update": (None, [vp, f32p, C.c_uint32, C.c_uint64]),
        "ref_pf_normalize": (None, [vp]),
        "ref_pf_resample": (C.c_float, [vp]),
        "ref_pf_uniform_sample": (C.c_int, [vp, C.c_int]),
        "ref_pf_mean": (None, [vp, f32p]),
        "ref_pf_best": (None, [vp, f32p]),
        "ref_pf_reloc_pose": (None, [vp, C.c_uint32, C.c_int, f32p, f32p, f32p]),
        "ref_pf_predict": (None, [vp, C.c_uint32, f64p, f64p, f32p]),
        "ref_trig_width": (C.c_int, []),
        "ref_mcl_create": (vp, [vp, f32p, C.c_uint64]),
        "ref_mcl_destroy": (None, [vp]),
        "ref_mcl_set": (None, [vp, f32p, C.c_uint64]),
        "ref_mcl_size": (C.c_uint64, [vp]),
        "ref_mcl_get": (None, [vp, f32p]),
        "ref_mcl_set_particle_num": (None, [vp, C.c_int, C.c_int]),
        "ref_mcl_add_particle_weight": (None, [vp]),
        "ref_mcl_compute_sample_num": (C.c_int, [vp, C.c_int]),
        "ref_mcl_update_sample_height": (None, [vp]),
        "ref_mcl_pf_resample": (None, [vp, C.c_int]),
        "ref_mcl_downsample": (C.c_uint64, [vp, f32p, C.c_uint64, C.c_double, C.c_int, f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _ref = L
    return L


class RefGrid:
    """amcl3d::Grid3d of the compiled reference with pc_info_/grid_info_ installed directly."""

    def __init__(self, m: Map | None = None):
        self.h = ref().ref_grid_create()
        if m is not None:
            self.install(m)

    def install(self, m: Map):
        mn = (C.c_double * 3)(*m.mn)
        mx = (C.c_double * 3)(*m.mx)
        size = (C.c_uint32 * 3)(*m.size)
        ref().ref_grid_install(self.h, mn, mx, m.resol, _fp(m.prob) if m.prob is not None else None, size, m.sensor_dev)

    def cloud_weight(self, cloud, pose):
        cloud = _as_f32(cloud, 3)
        pose = [float(np.float32(v)) for v in pose]
        return np.float32(ref().ref_grid_cloud_weight(self.h, _fp(cloud), cloud.shape[1], cloud.shape[0], *pose))

    def is_into_map(self, x, y, z):
        return bool(ref().ref_grid_is_into_map(self.h, x, y, z))

    def save(self, path):
        return bool(ref().ref_grid_save(self.h, str(path).encode()))

    def load(self, path, sensor_dev):
        return bool(ref().ref_grid_load(self.h, str(path).encode(), float(sensor_dev)))

    def open(self, map_dir, sensor_dev):
        return bool(ref().ref_grid_open(self.h, str(map_dir).encode(), float(sensor_dev)))

    def get(self, with_cells=True):
        mn = (C.c_double * 3)()
        mx = (C.c_double * 3)()
        resol = C.c_double(0)
        dev = C.c_double(0)
        size = (C.c_uint32 * 3)()
        rc = ref().ref_grid_get(self.h, mn, mx, C.byref(resol), size, C.byref(dev), None, 0)
        prob = None
        if rc == 1 and with_cells:
            cells = (int(size[0]) * int(size[1]) * int(size[2])) & 0xFFFFFFFF
            prob = np.zeros(cells, np.float32)
            ref().ref_grid_get(self.h, mn, mx, C.byref(resol), size, C.byref(dev), _fp(prob), cells)
        return dict(min=np.array(list(mn)), max=np.array(list(mx)), resol=resol.value, size=tuple(int(v) for v in size), sensor_dev=dev.value, prob=prob, has_grid=rc == 1)


class RefPF:
    """amcl3d::ParticleFilter of the compiled reference bound to a RefGrid."""

    def __init__(self, grid: RefGrid | None):
        self.grid = grid
        self.h = ref().ref_pf_create(grid.h if grid is not None else None)

    def set(self, p):
        p = np.ascontiguousarray(p, np.float32)
        ref().ref_pf_set(self.h, _fp(p), p.shape[0])

    def get(self):
        n = int(ref().ref_pf_size(self.h))
        out = np.zeros((n, 7), np.float32)
        ref().ref_pf_get(self.h, _fp(out))
        return out

    def seed(self, s):
        ref().ref_pf_seed(self.h, int(s))

    def update(self, cloud):
        cloud = _as_f32(cloud, 3)
        ref().ref_pf_update(self.h, _fp(cloud), cloud.shape[1], cloud.shape[0])

    def normalize(self):
        ref().ref_pf_normalize(self.h)

    def resample(self):
        """Returns the u01 the reference drew."""
        return np.float32(ref().ref_pf_resample(self.h))

    def uniform_sample(self, S):
        return int(ref().ref_pf_uniform_sample(self.h, int(S)))

    def mean(self):
        out = np.zeros(7, np.float32)
        ref().ref_pf_mean(self.h, _fp(out))
        return out

    def best(self):
        out = np.zeros(7, np.float32)
        ref().ref_pf_best(self.h, _fp(out))
        return out

    def reloc_pose(self, seed, n, init, dev):
        init = np.ascontiguousarray(init, np.float32)
        dev = np.ascontiguousarray(dev, np.float32)
        noise = np.zeros(max(6 * (abs(n) - 1), 1), np.float32)
        ref().ref_pf_reloc_pose(self.h, int(seed), int(n), _fp(init), _fp(dev), _fp(noise))
        return noise[: 6 * (abs(n) - 1)]

    def predict(self, seed, delta, noise_dev):
        delta = np.ascontiguousarray(delta, np.float64)
        nd = np.ascontiguousarray(noise_dev, np.float64)
        n = int(ref().ref_pf_size(self.h))
        noise = np.zeros(max(6 * n, 1), np.float32)
        ref().ref_pf_predict(self.h, int(seed), delta.ctypes.data_as(C.POINTER(C.c_double)), nd.ctypes.data_as(C.POINTER(C.c_double)), _fp(noise))
        return noise[: 6 * n]


class RefMCL:
    """amcl3d::MonteCarloLocalization of the compiled reference (amcl3d.cpp) bound to a RefGrid + keyposes."""

    def __init__(self, grid: RefGrid, keyposes):
        kp = _as_f32(keyposes, 3)
        self.grid = grid
        self.h = ref().ref_mcl_create(grid.h, _fp(kp), kp.shape[0])

    def set(self, p):
        p = np.ascontiguousarray(p, np.float32)
        ref().ref_mcl_set(self.h, _fp(p), p.shape[0])

    def get(self):
        n = int(ref().ref_mcl_size(self.h))
        out = np.zeros((n, 7), np.float32)
        ref().ref_mcl_get(self.h, _fp(out))
        return out

    def set_particle_num(self, mx, mn):
        ref().ref_mcl_set_particle_num(self.h, int(mx), int(mn))

    def add_particle_weight(self):
        ref().ref_mcl_add_particle_weight(self.h)

    def compute_sample_num(self, cnt_per_grid=3):
        return int(ref().ref_mcl_compute_sample_num(self.h, int(cnt_per_grid)))

    def update_sample_height(self):
        ref().ref_mcl_update_sample_height(self.h)

    def pf_resample(self, weight_multiply):
        ref().ref_mcl_pf_resample(self.h, int(weight_multiply))

    def downsample(self, xyzi, voxel_size, is_dense=True):
        xyzi = _as_f32(xyzi, 4)
        out = np.zeros((max(xyzi.shape[0], 1), 4), np.float32)
        k = ref().ref_mcl_downsample(self.h, _fp(xyzi), xyzi.shape[0], float(voxel_size), int(is_dense), _fp(out))
        return out[:k].copy()


def ref_compute_grid(xyz, resol, sensor_dev, which=0, bounds=None):
    """computePointCloud + computeGrid (which=0) / computeGrid2 (which=2) of the compiled reference."""
    xyz = _as_f32(xyz, 3)
    mn = (C.c_double * 3)()
    mx = (C.c_double * 3)()
    use_bounds = 0
    if bounds is not None:
        use_bounds = 1
        for a in range(3):
            mn[a], mx[a] = float(bounds[0][a]), float(bounds[1][a])
    size = (C.c_uint32 * 3)()
    # first call sizes the output (cheap for splat; for computeGrid we size from the geometry ourselves)
    m = None
    if bounds is None:
        bmn, bmx = compute_bounds(xyz)
    else:
        bmn, bmx = bounds
    m = Map(bmn, bmx, resol, sensor_dev)
    prob = np.zeros(m.n_cells, np.float32)
    cells = ref().ref_compute_grid(_fp(xyz), xyz.shape[1], xyz.shape[0], float(resol), float(sensor_dev), which, use_bounds, mn, mx, size, _fp(prob), prob.size)
    if cells < 0:
        raise RuntimeError(f"reference grid build failed ({cells})")
    assert cells == prob.size, (cells, prob.size)
    return dict(min=np.array(list(mn)), max=np.array(list(mx)), size=tuple(int(v) for v in size), prob=prob)
