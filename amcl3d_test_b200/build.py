"""Builds libamcl3d_b200.so (the C-ABI + sm_100a kernels) in-tree with nvcc.

The library is compiled for sm_100a only, with -fmad=false so that no f32/f64 multiply-add is ever
contracted (the reference is built for baseline x86-64 and rounds every operation), and
-lineinfo so ncu's source page maps to these files.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libamcl3d_b200.so"
SOURCES = ["update.cu", "pfops.cu", "gridbuild.cu", "mcl.cu", "voxelfilter.cu", "shard.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_native(verbose: bool = False, force: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    headers = [CSRC / "common.cuh", PKG.parent / "include" / "amcl3d_b200.h"]
    host_cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else None
    ccbin = ["-ccbin", host_cxx] if host_cxx else []
    objs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        objs.append(obj)
        if force or _stale(obj, [CSRC / src, *headers]):
            extra = os.environ.get("AMCL3D_NVCC_DEFS", "").split()  # e.g. "-DAMCL3D_UNROLL=16" for tuning sweeps
            cmd = [nvcc(), *ccbin, *NVCC_FLAGS, *extra, "-Xptxas", "-v", "-c", str(CSRC / src), "-o", str(obj)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            (objdir / (src + ".ptxas.txt")).write_text(r.stderr)
    if force or _stale(LIB, objs):
        cmd = [nvcc(), *ccbin, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs), "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


HOST_LIB = LIBDIR / "libamcl3d_host.so"
HOST_TEST = LIBDIR / "host_api_test"


def build_host(verbose: bool = False, force: bool = False) -> Path:
    """The reference-compatible C++ classes (include/amcl3d/*.h, host/*.cpp) over the C-ABI, plus the C++ test driver."""
    build_native()
    root = PKG.parent
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    srcs = [PKG / "host" / n for n in ("Grid3d.cpp", "ParticleFilter.cpp", "amcl3d.cpp", "map_io.cpp")]
    hdrs = list((root / "include").rglob("*.h")) + [PKG / "host" / "map_io.h"]
    common = [cxx, "-std=c++14", "-O2", "-fPIC", "-Wall", f"-I{root / 'include'}", f"-I{PKG / 'host'}"]
    if force or _stale(HOST_LIB, [*srcs, *hdrs, LIB]):
        cmd = [*common, "-shared", "-o", str(HOST_LIB), *map(str, srcs), f"-L{LIBDIR}", "-lamcl3d_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("host library build failed")
    drv = root / "tests" / "cpp" / "host_api_test.cpp"
    if drv.exists() and (force or _stale(HOST_TEST, [drv, HOST_LIB, *hdrs])):
        cmd = [*common, "-o", str(HOST_TEST), str(drv), f"-L{LIBDIR}", "-lamcl3d_host", "-lamcl3d_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("host test driver build failed")
    return HOST_LIB


if __name__ == "__main__":
    build_host(verbose="-v" in sys.argv)
    print(build_native(verbose="-v" in sys.argv, force="-f" in sys.argv))
