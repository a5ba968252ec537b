"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of oracle/raster_oracle.c (the plain-C CPU
restatement of the reference rasteriser). Never imported by the product package.

forward()/backward() mirror CudaRasterizer::Rasterizer::forward/backward
(DGR/cuda_rasterizer/rasterizer_impl.cu:197-447) for ONE view and return every intermediate the
reference keeps in GeometryState / BinningState / ImageState (rasterizer_impl.h:33-67).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgd_oracle.so")
        src = os.path.join(_HERE, "raster_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
        _LIB = ctypes.CDLL(path)
        _LIB.gdo_scan.restype = ctypes.c_uint32
        _LIB.gdo_higher_msb.restype = ctypes.c_uint32
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def forward(means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy, bg, *,
            shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
            scale_modifier=1.0, sh_degree=0):
    L = lib()
    means3D, opacities = _f(means3D), _f(opacities).reshape(-1)
    shs, colors_precomp, scales, rotations, cov3D_precomp = map(
        _f, (shs, colors_precomp, scales, rotations, cov3D_precomp))
    viewmatrix, projmatrix, campos, bg = (_f(np.asarray(x).reshape(-1)) for x in
                                          (viewmatrix, projmatrix, campos, bg))
    P = means3D.shape[0]
    M = shs.shape[1] if shs is not None else 0
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    st = {
        "P": P, "W": W, "H": H, "M": M, "D": sh_degree,
        "radii": np.zeros(P, np.int32), "means2D": np.zeros((P, 2), np.float32),
        "depths": np.zeros(P, np.float32), "cov3D": np.zeros((P, 6), np.float32),
        "rgb": np.zeros((P, 3), np.float32), "conic_opacity": np.zeros((P, 4), np.float32),
        "tiles_touched": np.zeros(P, np.uint32), "clamped": np.zeros((P, 3), np.uint8),
        "point_offsets": np.zeros(P, np.uint32),
    }
    L.gdo_preprocess(P, sh_degree, M, _p(means3D), _p(scales), ctypes.c_float(scale_modifier),
                     _p(rotations), _p(opacities), _p(shs), _p(cov3D_precomp), _p(colors_precomp),
                     _p(viewmatrix), _p(projmatrix), _p(campos), W, H, ctypes.c_float(tanfovx),
                     ctypes.c_float(tanfovy), _p(st["radii"]), _p(st["means2D"]), _p(st["depths"]),
                     _p(st["cov3D"]), _p(st["rgb"]), _p(st["conic_opacity"]),
                     _p(st["tiles_touched"]), _p(st["clamped"]))
    R = int(L.gdo_scan(P, _p(st["tiles_touched"]), _p(st["point_offsets"]))) if P else 0
    st["num_rendered"] = R
    st["keys_unsorted"] = np.zeros(R, np.uint64)
    st["values_unsorted"] = np.zeros(R, np.uint32)
    st["keys_sorted"] = np.zeros(R, np.uint64)
    st["point_list"] = np.zeros(R, np.uint32)
    st["ranges"] = np.zeros((T, 2), np.uint32)
    if R:
        L.gdo_binning(P, _p(st["means2D"]), _p(st["depths"]), _p(st["radii"]),
                      _p(st["point_offsets"]), W, H, ctypes.c_uint32(R), _p(st["keys_unsorted"]),
                      _p(st["values_unsorted"]), _p(st["keys_sorted"]), _p(st["point_list"]),
                      _p(st["ranges"]))
    feats = colors_precomp if colors_precomp is not None else st["rgb"]
    st["color"] = np.zeros((3, H, W), np.float32)
    st["depth"] = np.zeros((1, H, W), np.float32)
    st["alpha"] = np.zeros((1, H, W), np.float32)
    st["n_contrib"] = np.zeros((H, W), np.uint32)
    L.gdo_render_forward(W, H, _p(st["ranges"]), _p(st["point_list"]), _p(st["means2D"]), _p(feats),
                         _p(st["depths"]), _p(st["conic_opacity"]), _p(bg), _p(st["color"]),
                         _p(st["depth"]), _p(st["alpha"]), _p(st["n_contrib"]))
    st["_inputs"] = dict(means3D=means3D, shs=shs, colors_precomp=colors_precomp, scales=scales,
                         rotations=rotations, cov3D_precomp=cov3D_precomp, viewmatrix=viewmatrix,
                         projmatrix=projmatrix, campos=campos, bg=bg, tanfovx=tanfovx,
                         tanfovy=tanfovy, scale_modifier=scale_modifier)
    return st


def backward(st, dL_dcolor, dL_ddepth, dL_dalpha):
    L = lib()
    i = st["_inputs"]
    P, W, H, M, D = st["P"], st["W"], st["H"], st["M"], st["D"]
    dL_dcolor, dL_ddepth, dL_dalpha = _f(dL_dcolor), _f(dL_ddepth), _f(dL_dalpha)
    g = {
        "means2D": np.zeros((P, 3), np.float32), "conic": np.zeros((P, 4), np.float32),
        "opacity": np.zeros((P, 1), np.float32), "colors": np.zeros((P, 3), np.float32),
        "depths": np.zeros((P, 1), np.float32), "means3D": np.zeros((P, 3), np.float32),
        "cov3D": np.zeros((P, 6), np.float32), "sh": np.zeros((P, M, 3), np.float32),
        "scales": np.zeros((P, 3), np.float32), "rotations": np.zeros((P, 4), np.float32),
    }
    feats = i["colors_precomp"] if i["colors_precomp"] is not None else st["rgb"]
    L.gdo_render_backward(P, W, H, _p(st["ranges"]), _p(st["point_list"]), _p(i["bg"]),
                          _p(st["means2D"]), _p(st["conic_opacity"]), _p(feats), _p(st["depths"]),
                          _p(st["alpha"]), _p(st["n_contrib"]), _p(dL_dcolor), _p(dL_ddepth),
                          _p(dL_dalpha), _p(g["means2D"]), _p(g["conic"]), _p(g["opacity"]),
                          _p(g["colors"]), _p(g["depths"]))
    cov3D = i["cov3D_precomp"] if i["cov3D_precomp"] is not None else st["cov3D"]
    L.gdo_preprocess_backward(P, D, M, _p(i["means3D"]), _p(st["radii"]), _p(i["shs"]),
                              _p(st["clamped"]), _p(i["scales"]), _p(i["rotations"]),
                              ctypes.c_float(i["scale_modifier"]), _p(cov3D), _p(i["viewmatrix"]),
                              _p(i["projmatrix"]), W, H, ctypes.c_float(i["tanfovx"]),
                              ctypes.c_float(i["tanfovy"]), _p(i["campos"]), _p(g["means2D"]),
                              _p(g["conic"]), _p(g["means3D"]), _p(g["colors"]), _p(g["depths"]),
                              _p(g["cov3D"]), _p(g["sh"]), _p(g["scales"]), _p(g["rotations"]))
    return g


def mark_visible(means3D, viewmatrix):
    L = lib()
    means3D = _f(means3D)
    out = np.zeros(means3D.shape[0], np.uint8)
    L.gdo_mark_visible(means3D.shape[0], _p(means3D), _p(_f(np.asarray(viewmatrix).reshape(-1))), _p(out))
    return out.astype(bool)
