"""Host-side mirror of the parts of the reference GaussianModel that sit on the per-iteration path
(SURVEY.md s.8 row f2), on three fused CUDA kernels of libgd_raster.so (include/gd_raster.h):

  * the activations  get_xyz / get_features / get_opacity / get_scaling / get_rotation
        Garment_3DGS/gaussiansplatting/scene/gaussian_model.py:95-115
  * training_setup / update_learning_rate / optimizer.step() -- one Adam(eps=1e-15) over the
    parameter groups xyz, f_dc, opacity, scaling, rotation        :140-186
  * add_densification_stats + the max_radii2D update               :415-419,
        Garment_3DGS/threestudio/systems/GaussianDreamer.py:263-279

Names and argument meaning follow the reference; autograd is replaced by the explicit chain rule
inside gd_params_adam (the rasteriser hands back gradients w.r.t. the ACTIVATED parameters).
Densify / prune / PLY I/O are outside the hot path and not mirrored here.
"""
import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """Log-linear learning-rate decay (gaussiansplatting/utils/general_utils.py:29-62)."""
    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        return delay_rate * np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
    return helper


@dataclass
class OptimizationParams:
    """gaussiansplatting/arguments/__init__.py:70-81"""
    position_lr_init: float = 0.00005
    position_lr_final: float = 0.000025
    position_lr_delay_mult: float = 0.5
    position_lr_max_steps: int = 30_000
    feature_lr: float = 0.0125
    opacity_lr: float = 0.01
    scaling_lr: float = 0.005
    rotation_lr: float = 0.001
    percent_dense: float = 0.01


def _lib_params():
    L = _lib.raster_lib()
    if not getattr(L, "_gd_params_ready", False):
        vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.gd_params_activate.argtypes = [i, vp, vp, vp, vp, vp, vp, vp]
        L.gd_params_adam.argtypes = [i, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_float), f, f, f, i, vp]
        L.gd_densify_stats.argtypes = [i, i, vp, vp, vp, vp, vp, vp]
        for n in ("gd_params_activate", "gd_params_adam", "gd_densify_stats"):
            getattr(L, n).restype = ctypes.c_int
        L._gd_params_ready = True
    return L


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {_lib.raster_lib().gd_last_error().decode()}")


class GaussianParams:
    """Raw (pre-activation) parameters of sh_degree-0 Gaussians + Adam state, CUDA fp32."""

    def __init__(self, xyz, f_dc, opacity, scaling, rotation, spatial_lr_scale=1.0):
        dev = xyz.device
        if not xyz.is_cuda:
            raise RuntimeError("GaussianParams is CUDA-only (no CPU fallback)")
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous().clone()
        self._xyz, self._features_dc = f32(xyz), f32(f_dc).reshape(-1, 1, 3)
        self._opacity, self._scaling, self._rotation = f32(opacity).reshape(-1, 1), f32(scaling), f32(rotation)
        P = self._xyz.shape[0]
        self.spatial_lr_scale = spatial_lr_scale
        self.max_radii2D = torch.zeros(P, device=dev)
        self.xyz_gradient_accum = torch.zeros((P, 1), device=dev)
        self.denom = torch.zeros((P, 1), device=dev)
        self.exp_avg = self.exp_avg_sq = None
        self.step_count = 0
        self.lrs = None

    @property
    def P(self):
        return self._xyz.shape[0]

    # ---- gaussian_model.py:140-186 ---------------------------------------------------------
    def training_setup(self, training_args=None):
        a = training_args or OptimizationParams()
        self.percent_dense = a.percent_dense
        self.xyz_gradient_accum.zero_(); self.denom.zero_()
        self.lrs = [a.position_lr_init * self.spatial_lr_scale, a.feature_lr, a.opacity_lr, a.scaling_lr, a.rotation_lr]
        self.xyz_scheduler_args = get_expon_lr_func(lr_init=a.position_lr_init * self.spatial_lr_scale,
                                                    lr_final=a.position_lr_final * self.spatial_lr_scale,
                                                    lr_delay_mult=a.position_lr_delay_mult, max_steps=a.position_lr_max_steps)
        self.exp_avg = torch.zeros(14 * self.P, device=self._xyz.device)
        self.exp_avg_sq = torch.zeros(14 * self.P, device=self._xyz.device)
        self.step_count = 0

    def update_learning_rate(self, iteration):
        self.lrs[0] = float(self.xyz_scheduler_args(iteration))
        return self.lrs[0]

    # ---- gaussian_model.py:95-115, all five at once ------------------------------------------
    def activated(self, out=None):
        """Packed activated parameters [14P]: xyz | f_dc | sigmoid(opacity) | exp(scaling) | normalize(rotation)."""
        P = self.P
        out = torch.empty(14 * P, device=self._xyz.device) if out is None else out
        _chk(_lib_params().gd_params_activate(P, self._xyz.data_ptr(), self._features_dc.data_ptr(), self._opacity.data_ptr(),
                                             self._scaling.data_ptr(), self._rotation.data_ptr(), out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "gd_params_activate")
        return out

    @staticmethod
    def unpack(p, P):
        return (p[0:3 * P].view(P, 3), p[3 * P:6 * P].view(P, 1, 3), p[6 * P:7 * P].view(P, 1),
                p[7 * P:10 * P].view(P, 3), p[10 * P:14 * P].view(P, 4))

    get_xyz = property(lambda self: self._xyz)
    get_features = property(lambda self: self._features_dc)
    get_opacity = property(lambda self: self.unpack(self.activated(), self.P)[2])
    get_scaling = property(lambda self: self.unpack(self.activated(), self.P)[3])
    get_rotation = property(lambda self: self.unpack(self.activated(), self.P)[4])

    # ---- backward through the activations + optimizer.step() -----------------------------------
    def adam_step(self, packed_grad, beta1=0.9, beta2=0.999, eps=1e-15):
        """packed_grad [14P]: dL/d(activated parameters), e.g. the rasteriser backward output summed
        over views (and all-reduced over ranks)."""
        if self.exp_avg is None:
            raise RuntimeError("call training_setup() first")
        self.step_count += 1
        lr = (ctypes.c_float * 5)(*self.lrs)
        _chk(_lib_params().gd_params_adam(self.P, self._xyz.data_ptr(), self._features_dc.data_ptr(), self._opacity.data_ptr(),
                                         self._scaling.data_ptr(), self._rotation.data_ptr(), packed_grad.data_ptr(),
                                         self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), lr, beta1, beta2, eps,
                                         self.step_count, torch.cuda.current_stream().cuda_stream), "gd_params_adam")

    def adam_step_peers(self, px, densify=True, beta1=0.9, beta2=0.999, eps=1e-15):
        """Multi-GPU form of add_densification_stats + adam_step: consumes the reduced buffers of a
        parallel.PeerExchange whose allreduce() was launched on this stream (waits on the peers' flags on the device)."""
        if self.exp_avg is None:
            raise RuntimeError("call training_setup() first")
        self.step_count += 1
        lr = (ctypes.c_float * 5)(*self.lrs)
        _chk(px.L.gd_params_adam_peers(self.P, self._xyz.data_ptr(), self._features_dc.data_ptr(), self._opacity.data_ptr(),
                                       self._scaling.data_ptr(), self._rotation.data_ptr(), ctypes.byref(px.table), px.epoch,
                                       self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), lr, beta1, beta2, eps, self.step_count,
                                       int(densify), self.xyz_gradient_accum.data_ptr(), self.denom.data_ptr(),
                                       self.max_radii2D.data_ptr(), torch.cuda.current_stream().cuda_stream), "gd_params_adam_peers")

    # ---- gaussian_model.py:415-419 + GaussianDreamer.py:263-279 -----------------------------------
    def add_densification_stats(self, viewspace_point_grad_sum, radii):
        """viewspace_point_grad_sum [P,3]: sum over the views of dL/dmeans2D; radii i32 [B,P]."""
        radii = radii.reshape(-1, self.P).contiguous()
        _chk(_lib_params().gd_densify_stats(self.P, radii.shape[0], viewspace_point_grad_sum.contiguous().data_ptr(), radii.data_ptr(),
                                           self.xyz_gradient_accum.data_ptr(), self.denom.data_ptr(), self.max_radii2D.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "gd_densify_stats")


# ---- on-disk format at the stage boundary ----------------------------------
# last_3dgs.ply as written by GaussianModel.save_ply (gaussian_model.py:187-218) with plyfile:
# binary little-endian, one "vertex" element, float32 properties
#   x y z nx ny nz f_dc_0..2 [f_rest_*] opacity scale_0..2 rot_0..3      (raw, pre-activation values)
# Host-side I/O only (numpy); consumers: Normal_estimator_Metric3D / Garment_Deformer_NeTF.
def ply_attribute_names(n_rest=0):
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names += ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    return names


def save_ply(path, xyz, f_dc, opacity, scaling, rotation, f_rest=None):
    """Arrays / tensors of raw parameters: xyz [P,3], f_dc [P,1,3], opacity [P,1], scaling [P,3],
    rotation [P,4], optional f_rest [P,R,3] (stored channel-major like the reference's
    transpose(1,2).flatten())."""
    import os
    to_np = lambda t: (t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)).astype(np.float32)
    xyz, opacity, scaling, rotation = to_np(xyz), to_np(opacity).reshape(-1, 1), to_np(scaling), to_np(rotation)
    P = xyz.shape[0]
    dc = to_np(f_dc).reshape(P, 1, 3).transpose(0, 2, 1).reshape(P, 3)
    rest = np.zeros((P, 0), np.float32)
    if f_rest is not None and np.asarray(to_np(f_rest)).size:
        r = to_np(f_rest)
        rest = r.reshape(P, r.size // (3 * P), 3).transpose(0, 2, 1).reshape(P, -1)
    names = ply_attribute_names(rest.shape[1])
    table = np.concatenate([xyz, np.zeros_like(xyz), dc, rest, opacity, scaling, rotation], axis=1).astype("<f4")
    assert table.shape[1] == len(names)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(np.ascontiguousarray(table).tobytes())


def load_ply(path):
    """Reads a 3DGS .ply (binary little-endian float32 vertex table, any property order) and returns a
    dict of float32 numpy arrays shaped like the reference's load_ply (gaussian_model.py:225-262):
    xyz [P,3], f_dc [P,1,3], f_rest [P,R,3], opacity [P,1], scaling [P,3], rotation [P,4]."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").split("\n")
    if lines[0].strip() != "ply" or "binary_little_endian" not in lines[1]:
        raise ValueError("load_ply: only binary little-endian PLY is supported")
    P, names, in_vertex = 0, [], False
    for ln in lines:
        t = ln.split()
        if t[:2] == ["element", "vertex"]:
            P, in_vertex = int(t[2]), True
        elif t[:1] == ["element"]:
            in_vertex = False
        elif t[:1] == ["property"] and in_vertex:
            if t[1] not in ("float", "float32"):
                raise ValueError(f"load_ply: property {t[2]} is {t[1]}, expected float")
            names.append(t[2])
    table = np.frombuffer(data, dtype="<f4", count=P * len(names), offset=end).reshape(P, len(names))
    col = {n: i for i, n in enumerate(names)}
    pick = lambda ns: np.stack([table[:, col[n]] for n in ns], 1).astype(np.float32)
    by_idx = lambda prefix: sorted((n for n in names if n.startswith(prefix)), key=lambda x: int(x.split("_")[-1]))
    rest_names = by_idx("f_rest_")
    f_rest = pick(rest_names).reshape(P, 3, len(rest_names) // 3).transpose(0, 2, 1) if rest_names else np.zeros((P, 0, 3), np.float32)
    return {"xyz": pick(["x", "y", "z"]), "f_dc": pick(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(P, 1, 3),
            "f_rest": np.ascontiguousarray(f_rest), "opacity": pick(["opacity"]), "scaling": pick(by_idx("scale_")),
            "rotation": pick(by_idx("rot"))}
