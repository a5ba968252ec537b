"""ctypes loader for the C-ABI libraries declared in include/*.h (built by csrc/Makefile)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
GD_MAX_VIEWS = 32

c_float_p = ctypes.c_void_p  # device pointers travel as raw addresses


class GdView(ctypes.Structure):
    _fields_ = [
        ("viewmatrix", ctypes.c_void_p),
        ("projmatrix", ctypes.c_void_p),
        ("campos", ctypes.c_void_p),
        ("tanfovx", ctypes.c_float),
        ("tanfovy", ctypes.c_float),
    ]


class GdCounters(ctypes.Structure):
    _fields_ = [
        ("num_rendered", ctypes.c_uint32),
        ("overflow", ctypes.c_uint32),
        ("view_base", ctypes.c_uint32 * (GD_MAX_VIEWS + 1)),
        ("bwd_items", ctypes.c_uint32),
        ("bwd_next", ctypes.c_uint32),
        ("depth_max_bits", ctypes.c_uint32),
    ]


class GdFwdArgs(ctypes.Structure):
    _fields_ = [
        ("P", ctypes.c_int), ("D", ctypes.c_int), ("M", ctypes.c_int),
        ("W", ctypes.c_int), ("H", ctypes.c_int), ("B", ctypes.c_int),
        ("background", ctypes.c_void_p),
        ("means3D", ctypes.c_void_p),
        ("shs", ctypes.c_void_p),
        ("colors_precomp", ctypes.c_void_p),
        ("opacities", ctypes.c_void_p),
        ("scales", ctypes.c_void_p),
        ("scale_modifier", ctypes.c_float),
        ("rotations", ctypes.c_void_p),
        ("cov3D_precomp", ctypes.c_void_p),
        ("views", GdView * GD_MAX_VIEWS),
        ("prefiltered", ctypes.c_int),
        ("debug", ctypes.c_int),
        ("out_color", ctypes.c_void_p),
        ("out_depth", ctypes.c_void_p),
        ("out_alpha", ctypes.c_void_p),
        ("radii", ctypes.c_void_p),
        ("geom_buffer", ctypes.c_void_p), ("geom_bytes", ctypes.c_size_t),
        ("binning_buffer", ctypes.c_void_p), ("binning_bytes", ctypes.c_size_t),
        ("img_buffer", ctypes.c_void_p), ("img_bytes", ctypes.c_size_t),
        ("max_rendered", ctypes.c_uint32),
    ]


class GdBwdArgs(ctypes.Structure):
    _fields_ = [
        ("P", ctypes.c_int), ("D", ctypes.c_int), ("M", ctypes.c_int),
        ("W", ctypes.c_int), ("H", ctypes.c_int), ("B", ctypes.c_int),
        ("background", ctypes.c_void_p),
        ("means3D", ctypes.c_void_p),
        ("shs", ctypes.c_void_p),
        ("colors_precomp", ctypes.c_void_p),
        ("scales", ctypes.c_void_p),
        ("scale_modifier", ctypes.c_float),
        ("rotations", ctypes.c_void_p),
        ("cov3D_precomp", ctypes.c_void_p),
        ("views", GdView * GD_MAX_VIEWS),
        ("radii", ctypes.c_void_p),
        ("out_alpha", ctypes.c_void_p),
        ("dL_dcolor", ctypes.c_void_p),
        ("dL_ddepth", ctypes.c_void_p),
        ("dL_dalpha", ctypes.c_void_p),
        ("debug", ctypes.c_int),
        ("sum_views", ctypes.c_int),
        ("dL_dmeans2D", ctypes.c_void_p),
        ("dL_dcolors", ctypes.c_void_p),
        ("dL_dopacity", ctypes.c_void_p),
        ("dL_dmeans3D", ctypes.c_void_p),
        ("dL_dcov3D", ctypes.c_void_p),
        ("dL_dsh", ctypes.c_void_p),
        ("dL_dscales", ctypes.c_void_p),
        ("dL_drotations", ctypes.c_void_p),
        ("dL_dconic", ctypes.c_void_p),
        ("dL_ddepths", ctypes.c_void_p),
        ("geom_buffer", ctypes.c_void_p), ("geom_bytes", ctypes.c_size_t),
        ("binning_buffer", ctypes.c_void_p), ("binning_bytes", ctypes.c_size_t),
        ("img_buffer", ctypes.c_void_p), ("img_bytes", ctypes.c_size_t),
        ("max_rendered", ctypes.c_uint32),
    ]


class GdStateView(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "records", "tiles_touched", "point_offsets", "cov3D", "clamped", "counters", "point_list",
        "tile_keys", "sorted_records", "instance_slot", "instance_grad", "ranges", "n_contrib")]


_raster = None


def raster_lib():
    """Returns the loaded libgd_raster.so; raises (never falls back) if it has not been built."""
    global _raster
    if _raster is not None:
        return _raster
    path = os.environ.get("GD_RASTER_LIB") or os.path.join(LIB_DIR, "libgd_raster.so")   # override: tuning builds
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `make -C garmentdreamer_b200/csrc raster` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    lib.gd_last_error.restype = ctypes.c_char_p
    lib.gd_raster_version.restype = ctypes.c_char_p
    lib.gd_launch_count.restype = ctypes.c_uint64
    lib.gd_raster_state_bytes.restype = ctypes.c_int
    lib.gd_raster_state_bytes.argtypes = [ctypes.c_int] * 4 + [ctypes.c_uint32] + [
        ctypes.POINTER(ctypes.c_size_t)] * 3
    lib.gd_raster_state_view.restype = ctypes.c_int
    lib.gd_raster_state_view.argtypes = [ctypes.c_int] * 4 + [ctypes.c_uint32] + [
        ctypes.c_void_p] * 3 + [ctypes.POINTER(GdStateView)]
    lib.gd_raster_forward.restype = ctypes.c_int
    lib.gd_raster_forward.argtypes = [ctypes.POINTER(GdFwdArgs), ctypes.c_void_p]
    lib.gd_raster_backward.restype = ctypes.c_int
    lib.gd_raster_backward.argtypes = [ctypes.POINTER(GdBwdArgs), ctypes.c_void_p]
    lib.gd_mark_visible.restype = ctypes.c_int
    lib.gd_mark_visible.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5
    _raster = lib
    return lib


def last_error(lib):
    return lib.gd_last_error().decode("utf-8", "replace")
