"""The C-ABI library loads without a GPU, exports exactly what include/amcl3d_b200.h declares, and its
compute entry points fail loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "amcl3d_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amcl3d_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from amcl3d_test_b200 import capi

    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from amcl3d_test_b200 import capi

    L = capi.load()
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.lib_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (amcl3d_b200_\w+)", out))
    assert set(header_symbols()) <= exported
    assert L.amcl3d_b200_abi_version() == 2
    # nothing but the C-ABI leaks out of the library
    leaked = [l for l in out.splitlines() if " T " in l and "amcl3d_b200_" not in l]
    assert not leaked, leaked[:5]


def test_is_sm100a_only():
    from amcl3d_test_b200 import capi

    out = subprocess.run(["cuobjdump", "-lelf", str(capi.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+\w*)", out))
    assert archs == {"100a"}, archs


def _has_cuda():
    from amcl3d_test_b200 import capi

    L = capi.load()
    n = C.c_int(0)
    rc = L.amcl3d_b200_device_info(0, C.byref(n), None, None, None, None, None)
    return rc == capi.OK and n.value > 0


def test_fails_loudly_without_a_device():
    from amcl3d_test_b200 import capi

    if _has_cuda():
        pytest.skip("a CUDA device is present")
    L = capi.load()
    h = C.c_void_p()
    rc = L.amcl3d_b200_grid_create(0, C.byref(h))
    assert rc == capi.ERR_CUDA and not h.value
    assert L.amcl3d_b200_last_error()
    from amcl3d_test_b200 import Grid3d
    from amcl3d_test_b200.capi import Amcl3dError

    with pytest.raises(Amcl3dError):
        Grid3d(0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under amcl3d_test_b200/ or include/ may reference it."""
    for path in list((ROOT / "amcl3d_test_b200").rglob("*.py")) + list((ROOT / "amcl3d_test_b200").rglob("*.cu")) + \
            list((ROOT / "amcl3d_test_b200").rglob("*.cpp")) + list((ROOT / "include").rglob("*.h")):
        text = path.read_text()
        assert "oracle" not in text.replace("oracle/ is", ""), path
