"""CPU: the C-ABI libraries load and export every symbol include/*.h declares; host-side argument
validation works without a GPU (no compute calls here)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = {"gd_raster.h": "libgd_raster.so", "gd_unet.h": "libgd_unet.so"}


def declared_functions(header):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gd_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.parametrize("header", sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))))
def test_exports(header):
    libname = LIBS[os.path.basename(header)]
    path = os.path.join(ROOT, "garmentdreamer_b200", "lib", libname)
    assert os.path.exists(path), f"{path} not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(path)
    fns = declared_functions(header)
    assert fns, "no functions parsed from header"
    for fn in fns:
        assert hasattr(lib, fn), f"{libname} does not export {fn}"


def test_raster_host_side_validation():
    from garmentdreamer_b200 import _lib
    lib = _lib.raster_lib()
    assert b"sm_100a" in lib.gd_raster_version()
    g, b, i = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.gd_raster_state_bytes(100000, 512, 512, 4, 1 << 21, ctypes.byref(g), ctypes.byref(b), ctypes.byref(i)) == 0
    # 112 bytes per instance in the binning arena; 48-byte records per (view, Gaussian)
    assert b.value >= 112 * (1 << 21) and g.value >= 4 * 100000 * 48
    assert lib.gd_raster_state_bytes(10, 512, 512, 0, 16, None, None, None) == -1  # B out of range
    assert b"B must be" in lib.gd_last_error()
    a = _lib.GdFwdArgs()
    a.P, a.W, a.H, a.B = 10, 64, 64, 1
    assert lib.gd_raster_forward(ctypes.byref(a), None) == -1  # neither shs nor colors
    assert b"exactly one" in lib.gd_last_error()
    assert lib.gd_launch_count() == 0
