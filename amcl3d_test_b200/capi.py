"""ctypes binding of the C-ABI (include/amcl3d_b200.h).  Loading never falls back to anything:
if libamcl3d_b200.so is missing it is built with nvcc; if that fails the import fails."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import build as _build

OK, ERR_ARG, ERR_STATE, ERR_CUDA, ERR_IO, ERR_MISMATCH, ERR_EMPTY = range(7)
TRIG_F64, TRIG_F32 = 0, 1
GRID_QUIRK, GRID_GAUSSIAN, GRID_SPLAT = 0, 1, 2

# every symbol include/amcl3d_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "amcl3d_b200_abi_version", "amcl3d_b200_last_error", "amcl3d_b200_device_info",
    "amcl3d_b200_set_l2_fetch_granularity", "amcl3d_b200_get_l2_fetch_granularity", "amcl3d_b200_pf_set_chain_mode",
    "amcl3d_b200_grid_set_edt_fast", "amcl3d_b200_grid_build_device", "amcl3d_b200_grid_last_build_ms",
    "amcl3d_b200_pf_last_phase_ms",
    "amcl3d_b200_pf_shard_init", "amcl3d_b200_pf_shard_init_capacity", "amcl3d_b200_pf_shard_resize",
    "amcl3d_b200_pf_shard_gather_local", "amcl3d_b200_pf_shard_scatter_local", "amcl3d_b200_pf_shard_mean_exact_local", "amcl3d_b200_pf_shard_info", "amcl3d_b200_pf_shard_ipc_handle",
    "amcl3d_b200_pf_shard_attach_ipc", "amcl3d_b200_pf_shard_attach_local", "amcl3d_b200_pf_shard_resample",
    "amcl3d_b200_pf_shard_mean", "amcl3d_b200_pf_shard_resample_local", "amcl3d_b200_pf_shard_mean_local",
    "amcl3d_b200_pf_shard_normalize", "amcl3d_b200_pf_shard_normalize_local", "amcl3d_b200_grid_clone",
    "amcl3d_b200_host_pin", "amcl3d_b200_host_unpin", "amcl3d_b200_pf_shard_resample_fast",
    "amcl3d_b200_grid_create", "amcl3d_b200_grid_destroy", "amcl3d_b200_grid_get_info",
    "amcl3d_b200_grid_compute_bounds", "amcl3d_b200_grid_set_map", "amcl3d_b200_grid_build",
    "amcl3d_b200_grid_download_edt", "amcl3d_b200_grid_upload", "amcl3d_b200_grid_download",
    "amcl3d_b200_grid_save_file", "amcl3d_b200_grid_load_file", "amcl3d_b200_grid_is_into_map",
    "amcl3d_b200_grid_cloud_weight", "amcl3d_b200_grid_device_ptr", "amcl3d_b200_grid_use_codes",
    "amcl3d_b200_pf_create", "amcl3d_b200_pf_destroy", "amcl3d_b200_pf_set_trig_mode", "amcl3d_b200_pf_set_stream",
    "amcl3d_b200_pf_set_two_phase",
    "amcl3d_b200_pf_set_async", "amcl3d_b200_pf_sync",
    "amcl3d_b200_pf_set_particles", "amcl3d_b200_pf_get_particles", "amcl3d_b200_pf_size",
    "amcl3d_b200_pf_bind_device", "amcl3d_b200_pf_device_ptr", "amcl3d_b200_pf_set_cloud",
    "amcl3d_b200_pf_set_cloud_device", "amcl3d_b200_pf_set_cloud_downsampled", "amcl3d_b200_pf_get_cloud",
    "amcl3d_b200_pf_update", "amcl3d_b200_pf_update_cloud",
    "amcl3d_b200_pf_update_split", "amcl3d_b200_pf_fuse_range",
    "amcl3d_b200_pf_normalize", "amcl3d_b200_pf_scale_by_max", "amcl3d_b200_pf_resample",
    "amcl3d_b200_pf_uniform_sample", "amcl3d_b200_pf_mean", "amcl3d_b200_pf_mean_device", "amcl3d_b200_pf_best", "amcl3d_b200_pf_predict",
    "amcl3d_b200_pf_prefix", "amcl3d_b200_pf_last_kernel_ms", "amcl3d_b200_pf_launch_count",
    "amcl3d_b200_pf_compute_sample_num", "amcl3d_b200_pf_update_sample_height", "amcl3d_b200_pf_resample_step",
    "amcl3d_b200_bench_gather",
]


class Beacon(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("range", C.c_double)]


class GridInfo(C.Structure):
    _fields_ = [
        ("min", C.c_double * 3), ("max", C.c_double * 3), ("resol", C.c_double), ("sensor_dev", C.c_double),
        ("size", C.c_uint32 * 3), ("n_cells", C.c_uint64), ("has_map", C.c_int), ("has_grid", C.c_int),
        ("device", C.c_int), ("rep", C.c_int), ("lut_n", C.c_uint32), ("coded_bytes", C.c_uint64),
    ]


class Amcl3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"amcl3d_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib_path() -> Path:
    return _build.LIB


def load():
    """Loads (building first if needed) libamcl3d_b200.so and declares the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("AMCL3D_B200_LIB"):  # tuning sweeps: a variant of the library built with other kernel parameters
        path = Path(os.environ["AMCL3D_B200_LIB"])
    if not path.exists():
        _build.build_native()
    L = C.CDLL(str(path))
    P = C.POINTER
    vp, f32p, f64p, u32p, i32p = C.c_void_p, P(C.c_float), P(C.c_double), P(C.c_uint32), P(C.c_int32)
    u64, u32, f32, f64, i = C.c_uint64, C.c_uint32, C.c_float, C.c_double, C.c_int
    sig = {
        "amcl3d_b200_abi_version": (i, []),
        "amcl3d_b200_last_error": (C.c_char_p, []),
        "amcl3d_b200_device_info": (i, [i, P(i), P(i), P(i), P(i), P(u64), P(u64)]),
        "amcl3d_b200_pf_set_chain_mode": (i, [vp, i]),
        "amcl3d_b200_grid_set_edt_fast": (i, [vp, i]),
        "amcl3d_b200_grid_build_device": (i, [vp, vp, u32, u64, f64, i, i, i]),
        "amcl3d_b200_grid_last_build_ms": (i, [vp, P(f32)]),
        "amcl3d_b200_pf_last_phase_ms": (i, [vp, P(f32)]),
        "amcl3d_b200_pf_shard_init": (i, [vp, u64, i, i]),
        "amcl3d_b200_pf_shard_init_capacity": (i, [vp, u64, u64, i, i]),
        "amcl3d_b200_pf_shard_resize": (i, [vp, u64]),
        "amcl3d_b200_pf_shard_gather_local": (i, [P(vp), i, vp]),
        "amcl3d_b200_pf_shard_scatter_local": (i, [vp, P(vp), i]),
        "amcl3d_b200_pf_shard_info": (i, [vp, P(u64), P(u64), P(u64)]),
        "amcl3d_b200_pf_shard_ipc_handle": (i, [vp, vp, C.c_size_t]),
        "amcl3d_b200_pf_shard_attach_ipc": (i, [vp, vp, C.c_size_t]),
        "amcl3d_b200_pf_shard_attach_local": (i, [P(vp), i]),
        "amcl3d_b200_pf_shard_resample": (i, [vp, f32, i, vp, f32, vp]),
        "amcl3d_b200_pf_shard_normalize": (i, [vp, P(f32)]),
        "amcl3d_b200_pf_shard_normalize_local": (i, [P(vp), i]),
        "amcl3d_b200_grid_clone": (i, [vp, i, i, P(vp)]),
        "amcl3d_b200_host_pin": (i, [vp, C.c_size_t]),
        "amcl3d_b200_host_unpin": (i, [vp]),
        "amcl3d_b200_pf_shard_resample_fast": (i, [vp, f32, vp]),
        "amcl3d_b200_pf_shard_mean": (i, [vp, vp, vp]),
        "amcl3d_b200_pf_shard_resample_local": (i, [P(vp), i, f32, i, P(vp), f32]),
        "amcl3d_b200_pf_shard_mean_local": (i, [P(vp), i, vp]),
        "amcl3d_b200_pf_shard_mean_exact_local": (i, [P(vp), i, vp]),
        "amcl3d_b200_set_l2_fetch_granularity": (i, [i, i]),
        "amcl3d_b200_get_l2_fetch_granularity": (i, [i, P(i)]),
        "amcl3d_b200_grid_create": (i, [i, P(vp)]),
        "amcl3d_b200_grid_destroy": (i, [vp]),
        "amcl3d_b200_grid_get_info": (i, [vp, P(GridInfo)]),
        "amcl3d_b200_grid_compute_bounds": (i, [vp, vp, u32, u64, f64, f64p, f64p]),
        "amcl3d_b200_grid_set_map": (i, [vp, f64p, f64p, f64]),
        "amcl3d_b200_grid_build": (i, [vp, vp, u32, u64, f64, i, i, i]),
        "amcl3d_b200_grid_download_edt": (i, [vp, vp, u64]),
        "amcl3d_b200_grid_upload": (i, [vp, vp, u32p, f64]),
        "amcl3d_b200_grid_download": (i, [vp, vp, u64]),
        "amcl3d_b200_grid_save_file": (i, [vp, C.c_char_p]),
        "amcl3d_b200_grid_load_file": (i, [vp, C.c_char_p, f64]),
        "amcl3d_b200_grid_is_into_map": (i, [vp, f32, f32, f32, P(i)]),
        "amcl3d_b200_grid_cloud_weight": (i, [vp, vp, u32, u32, f32, f32, f32, f32, f32, f32, i, P(f32)]),
        "amcl3d_b200_grid_device_ptr": (i, [vp, P(vp)]),
        "amcl3d_b200_grid_use_codes": (i, [vp, i]),
        "amcl3d_b200_pf_create": (i, [vp, P(vp)]),
        "amcl3d_b200_pf_destroy": (i, [vp]),
        "amcl3d_b200_pf_set_trig_mode": (i, [vp, i]),
        "amcl3d_b200_pf_set_stream": (i, [vp, vp]),
        "amcl3d_b200_pf_set_two_phase": (i, [vp, i]),
        "amcl3d_b200_pf_set_particles": (i, [vp, vp, u64]),
        "amcl3d_b200_pf_get_particles": (i, [vp, vp, u64, P(u64)]),
        "amcl3d_b200_pf_size": (i, [vp, P(u64)]),
        "amcl3d_b200_pf_bind_device": (i, [vp, vp, u64, u64]),
        "amcl3d_b200_pf_device_ptr": (i, [vp, P(vp), P(u64)]),
        "amcl3d_b200_pf_set_async": (i, [vp, i]),
        "amcl3d_b200_pf_sync": (i, [vp]),
        "amcl3d_b200_pf_set_cloud": (i, [vp, vp, u32, u32]),
        "amcl3d_b200_pf_set_cloud_device": (i, [vp, vp, u32, u32]),
        "amcl3d_b200_pf_set_cloud_downsampled": (i, [vp, vp, u32, i, u32, f32, i, P(u32)]),
        "amcl3d_b200_pf_get_cloud": (i, [vp, vp, u32, P(u32)]),
        "amcl3d_b200_pf_update": (i, [vp, P(Beacon), u32, f32, f32, P(u64)]),
        "amcl3d_b200_pf_update_cloud": (i, [vp, vp, u32, u32]),
        "amcl3d_b200_pf_update_split": (i, [vp, P(Beacon), u32, f32, vp, P(u64)]),
        "amcl3d_b200_pf_fuse_range": (i, [vp, vp, f32]),
        "amcl3d_b200_pf_normalize": (i, [vp, P(f32)]),
        "amcl3d_b200_pf_scale_by_max": (i, [vp]),
        "amcl3d_b200_pf_resample": (i, [vp, f32, u64, u64, vp, vp]),
        "amcl3d_b200_pf_uniform_sample": (i, [vp, i, u64, u64, vp, vp, P(i)]),
        "amcl3d_b200_pf_mean": (i, [vp, i, vp]),
        "amcl3d_b200_pf_mean_device": (i, [vp, i, vp]),
        "amcl3d_b200_pf_best": (i, [vp, vp]),
        "amcl3d_b200_pf_predict": (i, [vp, f64p, vp]),
        "amcl3d_b200_pf_prefix": (i, [vp, P(vp), vp, u64]),
        "amcl3d_b200_pf_compute_sample_num": (i, [vp, i, i, i, P(i)]),
        "amcl3d_b200_pf_update_sample_height": (i, [vp, vp, u64, P(f32)]),
        "amcl3d_b200_pf_resample_step": (i, [vp, i, i, i, vp, u64, P(i)]),
        "amcl3d_b200_pf_last_kernel_ms": (i, [vp, P(f32)]),
        "amcl3d_b200_pf_launch_count": (i, [vp, P(u64)]),
        "amcl3d_b200_bench_gather": (i, [vp, u64, u32, u32, i, P(f32), P(u64)]),
    }
    assert sorted(sig) == sorted(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != OK:
        raise Amcl3dError(rc, load().amcl3d_b200_last_error().decode(errors="replace"))
