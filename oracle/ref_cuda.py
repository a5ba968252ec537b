"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/_ref/libgd_ref_raster.so, i.e. the
UNMODIFIED reference CUDA rasteriser core (compiled by oracle/Makefile from the sources under
/root/reference, plus oracle/ref_shim.cu). Needs a GPU. Used to pin the oracle and the product
kernels against the real reference and to generate tests/golden/*.npz.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libgd_ref_raster.so")


def available():
    return os.path.exists(SO) and torch.cuda.is_available()


class _State(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "depths", "clamped", "internal_radii", "means2D", "cov3D", "conic_opacity", "rgb",
        "point_offsets", "tiles_touched", "point_list_keys_unsorted", "point_list_keys",
        "point_list_unsorted", "point_list", "ranges", "n_contrib")] + [
        ("P", ctypes.c_int), ("R", ctypes.c_int), ("W", ctypes.c_int), ("H", ctypes.c_int)]


def _dp(t):
    return None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr())


class RefRasterizer:
    def __init__(self):
        self.lib = ctypes.CDLL(SO)
        self.lib.gdref_create.restype = ctypes.c_void_p
        self.lib.gdref_last_error.restype = ctypes.c_char_p
        self.lib.gdref_last_error.argtypes = [ctypes.c_void_p]
        self.lib.gdref_destroy.argtypes = [ctypes.c_void_p]
        self.h = ctypes.c_void_p(self.lib.gdref_create())
        self.rt = ctypes.CDLL("libcudart.so.12")

    def __del__(self):
        try:
            self.lib.gdref_destroy(self.h)
        except Exception:
            pass

    def _d2h(self, ptr, nbytes, dtype):
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        if nbytes:
            rc = self.rt.cudaMemcpy(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(ptr),
                                    ctypes.c_size_t(nbytes), 2)
            assert rc == 0, f"cudaMemcpy failed {rc}"
        return out

    def forward(self, means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy,
                bg, *, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, scale_modifier=1.0, sh_degree=0, debug=False):
        """All tensors cuda fp32 contiguous. Returns dict(color, depth, alpha, radii, R)."""
        dev = means3D.device
        P = means3D.shape[0]
        M = shs.shape[1] if shs is not None and shs.numel() else 0
        color = torch.zeros((3, H, W), device=dev)
        depth = torch.zeros((1, H, W), device=dev)
        alpha = torch.zeros((1, H, W), device=dev)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        R = self.lib.gdref_forward(
            self.h, P, sh_degree, M, _dp(bg), W, H, _dp(means3D), _dp(shs), _dp(colors_precomp),
            _dp(opacities), _dp(scales), ctypes.c_float(scale_modifier), _dp(rotations),
            _dp(cov3D_precomp), _dp(viewmatrix), _dp(projmatrix), _dp(campos),
            ctypes.c_float(tanfovx), ctypes.c_float(tanfovy), 0, _dp(color), _dp(depth),
            _dp(alpha), _dp(radii), int(debug))
        torch.cuda.synchronize()
        if R < 0:
            raise RuntimeError(self.lib.gdref_last_error(self.h).decode())
        self._last = dict(P=P, M=M, D=sh_degree, W=W, H=H, R=R)
        return {"color": color, "depth": depth, "alpha": alpha, "radii": radii, "R": R}

    def state(self):
        s = _State()
        self.lib.gdref_get_state(self.h, ctypes.byref(s))
        P, R, W, H = s.P, s.R, s.W, s.H
        T = ((W + 15) // 16) * ((H + 15) // 16)
        st = {
            "depths": self._d2h(s.depths, 4 * P, np.float32),
            "clamped": self._d2h(s.clamped, 3 * P, np.uint8).reshape(P, 3),
            "means2D": self._d2h(s.means2D, 8 * P, np.float32).reshape(P, 2),
            "cov3D": self._d2h(s.cov3D, 24 * P, np.float32).reshape(P, 6),
            "conic_opacity": self._d2h(s.conic_opacity, 16 * P, np.float32).reshape(P, 4),
            "rgb": self._d2h(s.rgb, 12 * P, np.float32).reshape(P, 3),
            "point_offsets": self._d2h(s.point_offsets, 4 * P, np.uint32),
            "tiles_touched": self._d2h(s.tiles_touched, 4 * P, np.uint32),
            "num_rendered": R,
        }
        if R > 0:
            st["keys_unsorted"] = self._d2h(s.point_list_keys_unsorted, 8 * R, np.uint64)
            st["keys_sorted"] = self._d2h(s.point_list_keys, 8 * R, np.uint64)
            st["values_unsorted"] = self._d2h(s.point_list_unsorted, 4 * R, np.uint32)
            st["point_list"] = self._d2h(s.point_list, 4 * R, np.uint32)
        else:
            st["keys_unsorted"] = st["keys_sorted"] = np.zeros(0, np.uint64)
            st["values_unsorted"] = st["point_list"] = np.zeros(0, np.uint32)
        st["ranges"] = self._d2h(s.ranges, 8 * T, np.uint32).reshape(T, 2)
        st["n_contrib"] = self._d2h(s.n_contrib, 4 * W * H, np.uint32).reshape(H, W)
        return st

    def backward(self, means3D, radii, alpha, viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg,
                 dL_dcolor, dL_ddepth, dL_dalpha, *, shs=None, colors_precomp=None, scales=None,
                 rotations=None, cov3D_precomp=None, scale_modifier=1.0, debug=False):
        L = self._last
        P, M, W, H, R, D = L["P"], L["M"], L["W"], L["H"], L["R"], L["D"]
        dev = means3D.device
        z = lambda *s: torch.zeros(s, device=dev)
        g = {"means2D": z(P, 3), "conic": z(P, 2, 2), "opacity": z(P, 1), "colors": z(P, 3),
             "depths": z(P, 1), "means3D": z(P, 3), "cov3D": z(P, 6), "sh": z(P, M, 3),
             "scales": z(P, 3), "rotations": z(P, 4)}
        torch.cuda.synchronize()
        rc = self.lib.gdref_backward(
            self.h, P, D, M, R, _dp(bg), W, H, _dp(means3D), _dp(shs), _dp(colors_precomp),
            _dp(alpha), _dp(scales), ctypes.c_float(scale_modifier), _dp(rotations),
            _dp(cov3D_precomp), _dp(viewmatrix), _dp(projmatrix), _dp(campos),
            ctypes.c_float(tanfovx), ctypes.c_float(tanfovy), _dp(radii), _dp(dL_dcolor),
            _dp(dL_ddepth), _dp(dL_dalpha), ctypes.c_void_p(g["means2D"].data_ptr()),
            ctypes.c_void_p(g["conic"].data_ptr()), ctypes.c_void_p(g["opacity"].data_ptr()),
            ctypes.c_void_p(g["colors"].data_ptr()), ctypes.c_void_p(g["depths"].data_ptr()),
            ctypes.c_void_p(g["means3D"].data_ptr()), ctypes.c_void_p(g["cov3D"].data_ptr()),
            ctypes.c_void_p(g["sh"].data_ptr()) if M else ctypes.c_void_p(g["colors"].data_ptr()),
            ctypes.c_void_p(g["scales"].data_ptr()), ctypes.c_void_p(g["rotations"].data_ptr()),
            int(debug))
        torch.cuda.synchronize()
        if rc < 0:
            raise RuntimeError(self.lib.gdref_last_error(self.h).decode())
        return g
