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


def test_ctypes_mirrors_match_the_c_headers(tmp_path):
    """Every struct of include/*.h has a ctypes mirror on the Python side of the boundary: compile the headers as C (gcc) and
    compare sizeof and the offset of every field, so an argument added on one side only cannot go unnoticed."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    from garmentdreamer_b200 import _lib
    from garmentdreamer_b200.parallel import GdPeerTable
    from garmentdreamer_b200.unet_ops import GdGemmArgs
    mirrors = {"GdView": _lib.GdView, "GdCounters": _lib.GdCounters, "GdFwdArgs": _lib.GdFwdArgs, "GdBwdArgs": _lib.GdBwdArgs,
               "GdStateView": _lib.GdStateView, "GdPeerTable": GdPeerTable, "GdGemmArgs": GdGemmArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gd_raster.h"', '#include "gd_unet.h"', "int main(void) {"]
    for name, cls in mirrors.items():
        lines.append(f'  printf("{name} sizeof %zu\\n", sizeof({name}));')
        for field in cls._fields_:
            lines.append(f'  printf("{name} {field[0]} %zu\\n", offsetof({name}, {field[0]}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-x", "c", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    for line in out.strip().splitlines():
        name, what, value = line.split()
        cls = mirrors[name]
        mine = ctypes.sizeof(cls) if what == "sizeof" else getattr(cls, what).offset
        assert mine == int(value), f"{name}.{what}: ctypes {mine} != C {value}"
