"""Batched camera construction on the GPU: mirror of the reference
``Camera`` (Garment_3DGS/gaussiansplatting/scene/cameras.py:19-53) for a whole view batch.

The reference builds every camera on the CPU inside the per-iteration loop
(threestudio/systems/GaussianDreamer.py:189-191): two 4x4 LU inversions, two H2D copies and a GPU
inverse per view. ``cameras_from_c2w`` does the batch with one kernel of libgd_raster.so and
returns the ``raster.View`` list the batched rasteriser takes."""
import ctypes
import math

import torch

from . import _lib, raster

ZNEAR, ZFAR = 0.01, 100.0   # cameras.py:41-42


def fov2focal(fov, pixels):
    return pixels / (2 * math.tan(fov / 2))


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def cameras_from_c2w(c2w, fovy, height, width):
    """c2w: CUDA fp32 [B,4,4] (batch['c2w_3dgs']); fovy: B floats (radians; tensor or list).
    Returns (views, packed) with packed [B,35] = world_view_transform | full_proj_transform |
    camera_center on the device and views = [raster.View] pointing into it."""
    if not c2w.is_cuda:
        raise RuntimeError("cameras_from_c2w is CUDA-only (no CPU fallback)")
    B = c2w.shape[0]
    c2w = c2w.detach().to(torch.float32).contiguous()
    fovy = [float(f) for f in (fovy.tolist() if torch.is_tensor(fovy) else fovy)]
    fovx = [focal2fov(fov2focal(f, height), width) for f in fovy]          # cameras.py:24
    tx = (ctypes.c_float * B)(*[math.tan(f / 2) for f in fovx])            # graphics_utils.py:74-75
    ty = (ctypes.c_float * B)(*[math.tan(f / 2) for f in fovy])
    out = torch.empty((B, 35), dtype=torch.float32, device=c2w.device)
    L = _lib.raster_lib()
    L.gd_cameras_from_c2w.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                                      ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    L.gd_cameras_from_c2w.restype = ctypes.c_int
    rc = L.gd_cameras_from_c2w(B, c2w.data_ptr(), tx, ty, ZNEAR, ZFAR, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"gd_cameras_from_c2w failed ({rc}): {L.gd_last_error().decode()}")
    # GaussianRasterizationSettings takes tan(FoV * 0.5) (gaussian_renderer/__init__.py:33-34)
    views = [raster.View(out[b, 0:16], out[b, 16:32], out[b, 32:35], math.tan(fovx[b] * 0.5), math.tan(fovy[b] * 0.5))
             for b in range(B)]
    return views, out
