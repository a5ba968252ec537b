"""Torch-facing binding of the C-ABI rasteriser (include/gd_raster.h).

Host-side mirror of the reference's torch binding
(Garment_3DGS/gaussiansplatting/submodules/diff-gaussian-rasterization/rasterize_points.cu:35-229
and diff_gaussian_rasterization/__init__.py:44-158), extended to B views per call:

* ``forward_views`` / ``backward_views``  -- raw calls, caller keeps the returned ``RasterState``;
* ``RasterizeViews``                      -- autograd.Function over B views (the batched hot path);
* ``mark_visible``.

PyTorch is used for device memory and streams only; every kernel is in libgd_raster.so.
"""
import ctypes
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import GD_MAX_VIEWS


@dataclass
class View:
    """Per-view constants (reference: GaussianRasterizationSettings, DGR __init__.py:160-172)."""
    viewmatrix: torch.Tensor  # [4,4] cuda fp32, transposed world->camera
    projmatrix: torch.Tensor  # [4,4] cuda fp32, transposed full projection
    campos: torch.Tensor      # [3]
    tanfovx: float
    tanfovy: float


@dataclass
class RasterState:
    """Caller-owned state of one forward call, needed by backward (geom/binning/img buffers)."""
    P: int
    W: int
    H: int
    B: int
    D: int
    M: int
    cap: int
    geom: torch.Tensor
    binning: torch.Tensor
    img: torch.Tensor
    views: List[View]
    num_rendered: Optional[int] = None
    view_base: Optional[List[int]] = None
    keep: list = field(default_factory=list)  # tensors that must outlive the async launches


# capacity guess per (P, W, H, B): grows from observed instance counts, never shrinks
_cap_hint = {}


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    if t is None or t.numel() == 0:
        return None
    if t.device != device:
        raise ValueError(f"tensor on {t.device}, expected {device}")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _fill_views(arr, views: Sequence[View], device, keep):
    for b, v in enumerate(views):
        vm, pm, cp = _f32c(v.viewmatrix, device), _f32c(v.projmatrix, device), _f32c(v.campos, device)
        keep.extend([vm, pm, cp])
        arr[b].viewmatrix = vm.data_ptr()
        arr[b].projmatrix = pm.data_ptr()
        arr[b].campos = cp.data_ptr()
        arr[b].tanfovx = float(v.tanfovx)
        arr[b].tanfovy = float(v.tanfovy)


def state_bytes(P, W, H, B, cap):
    lib = _lib.raster_lib()
    g, b, i = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    rc = lib.gd_raster_state_bytes(P, W, H, B, cap, ctypes.byref(g), ctypes.byref(b), ctypes.byref(i))
    if rc != 0:
        raise RuntimeError(f"gd_raster_state_bytes: {_lib.last_error(lib)}")
    return g.value, b.value, i.value


def read_counters(state: RasterState):
    """Blocking read-back of the device counters (the reference's cudaMemcpy of num_rendered)."""
    lib = _lib.raster_lib()
    sv = _lib.GdStateView()
    lib.gd_raster_state_view(state.P, state.W, state.H, state.B, state.cap, state.geom.data_ptr(),
                             state.binning.data_ptr(), state.img.data_ptr(), ctypes.byref(sv))
    off = sv.counters - state.geom.data_ptr()
    n32 = ctypes.sizeof(_lib.GdCounters) // 4
    host = state.geom[off:off + 4 * n32].view(torch.int32).cpu()
    vals = [int(x) & 0xFFFFFFFF for x in host.tolist()]
    state.num_rendered = vals[0]
    state.view_base = vals[2:2 + state.B + 1]
    return vals[0], bool(vals[1])


def forward_views(means3D, opacities, views: Sequence[View], W: int, H: int, bg, *, shs=None,
                  colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                  scale_modifier=1.0, sh_degree=0, prefiltered=False, debug=False, cap=None,
                  sync=True):
    """Rasterises B views. Returns (color[B,3,H,W], depth[B,1,H,W], alpha[B,1,H,W], radii[B,P], state).

    sync=True reads the instance count back (as the reference does) and transparently re-runs
    with a larger arena on overflow; sync=False never touches the host (caller checks later).
    """
    lib = _lib.raster_lib()
    if not means3D.is_cuda:
        raise RuntimeError("garmentdreamer_b200 rasteriser is CUDA-only (no CPU fallback)")
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:57-59
    dev = means3D.device
    B = len(views)
    if not 1 <= B <= GD_MAX_VIEWS:
        raise ValueError(f"1..{GD_MAX_VIEWS} views per call")
    P = means3D.shape[0]
    means3D = _f32c(means3D, dev) if P else means3D
    opacities = _f32c(opacities, dev)
    shs, colors_precomp = _f32c(shs, dev), _f32c(colors_precomp, dev)
    scales, rotations, cov3D_precomp = _f32c(scales, dev), _f32c(rotations, dev), _f32c(cov3D_precomp, dev)
    bg = _f32c(bg, dev)
    M = shs.shape[1] if shs is not None else 0
    key = (P, W, H, B)
    if cap is None:
        cap = _cap_hint.get(key, max(1 << 16, 8 * P * B))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        color = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        alpha = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros((B, P), dtype=torch.int32, device=dev)
        while True:
            gb, bb, ib = state_bytes(P, W, H, B, cap)
            state = RasterState(P, W, H, B, int(sh_degree), M, cap,
                                torch.empty(gb, dtype=torch.uint8, device=dev),
                                torch.empty(bb, dtype=torch.uint8, device=dev),
                                torch.empty(ib, dtype=torch.uint8, device=dev), list(views))
            a = _lib.GdFwdArgs()
            a.P, a.D, a.M, a.W, a.H, a.B = P, int(sh_degree), M, W, H, B
            a.background = _ptr(bg)
            a.means3D = _ptr(means3D)
            a.shs, a.colors_precomp = _ptr(shs), _ptr(colors_precomp)
            a.opacities = _ptr(opacities)
            a.scales, a.rotations, a.cov3D_precomp = _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp)
            a.scale_modifier = float(scale_modifier)
            _fill_views(a.views, views, dev, state.keep)
            a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
            a.out_color, a.out_depth, a.out_alpha = color.data_ptr(), depth.data_ptr(), alpha.data_ptr()
            a.radii = radii.data_ptr()
            a.geom_buffer, a.geom_bytes = state.geom.data_ptr(), gb
            a.binning_buffer, a.binning_bytes = state.binning.data_ptr(), bb
            a.img_buffer, a.img_bytes = state.img.data_ptr(), ib
            a.max_rendered = cap
            if P == 0:
                # reference: empty outputs stay zero, rendered = 0 (rasterize_points.cu:82-118)
                color.zero_(); depth.zero_(); alpha.zero_()
                state.num_rendered, state.view_base = 0, [0] * (B + 1)
                return color, depth, alpha, radii, state
            rc = lib.gd_raster_forward(ctypes.byref(a), stream)
            if rc != 0:
                raise RuntimeError(f"gd_raster_forward failed ({rc}): {_lib.last_error(lib)}")
            if not sync:
                return color, depth, alpha, radii, state
            n, overflow = read_counters(state)
            _cap_hint[key] = max(_cap_hint.get(key, 0), int(n * 1.25) + 4096)
            if not overflow:
                return color, depth, alpha, radii, state
            cap = int(n * 1.25) + 4096


def backward_views(state: RasterState, means3D, radii, out_alpha, bg, dL_dcolor, dL_ddepth,
                   dL_dalpha, *, shs=None, colors_precomp=None, scales=None, rotations=None,
                   cov3D_precomp=None, scale_modifier=1.0, sum_views=False, debug=False,
                   want_aux=False, out=None):
    """Returns a dict of gradients; leading dim B unless sum_views (reference order of
    rasterize_points.cu:207: means2D, colors, opacity, means3D, cov3D, sh, scales, rotations)."""
    lib = _lib.raster_lib()
    dev = means3D.device
    P, W, H, B, M = state.P, state.W, state.H, state.B, state.M
    lead = () if sum_views else (B,)
    f = dict(dtype=torch.float32, device=dev)
    means3D = _f32c(means3D, dev) if P else means3D
    shs, colors_precomp = _f32c(shs, dev), _f32c(colors_precomp, dev)
    scales, rotations, cov3D_precomp = _f32c(scales, dev), _f32c(rotations, dev), _f32c(cov3D_precomp, dev)
    bg = _f32c(bg, dev)
    g = {
        "means2D": torch.empty(lead + (P, 3), **f), "colors": torch.empty(lead + (P, 3), **f),
        "opacity": torch.empty(lead + (P, 1), **f), "means3D": torch.empty(lead + (P, 3), **f),
        "cov3D": torch.empty(lead + (P, 6), **f), "sh": torch.empty(lead + (P, M, 3), **f),
        "scales": torch.empty(lead + (P, 3), **f), "rotations": torch.empty(lead + (P, 4), **f),
    }
    if out:  # caller-provided destinations (e.g. slices of one flat all-reduce buffer)
        for k, v in out.items():
            assert v.is_contiguous() and v.dtype == torch.float32 and v.numel() == g[k].numel(), k
            g[k] = v
    if shs is None:
        g["sh"].zero_()
    if scales is None:
        g["scales"].zero_(); g["rotations"].zero_()
    if want_aux:
        g["conic"] = torch.empty(lead + (P, 2, 2), **f)
        g["depths"] = torch.empty(lead + (P, 1), **f)
    if P == 0:
        return g
    dL_dcolor = _f32c(dL_dcolor, dev).reshape(B, 3, H, W)
    dL_ddepth = _f32c(dL_ddepth, dev).reshape(B, 1, H, W)
    dL_dalpha = _f32c(dL_dalpha, dev).reshape(B, 1, H, W)
    out_alpha = _f32c(out_alpha, dev)
    radii = radii.contiguous()
    a = _lib.GdBwdArgs()
    a.P, a.D, a.M, a.W, a.H, a.B = P, state.D, M, W, H, B
    a.background, a.means3D = _ptr(bg), _ptr(means3D)
    a.shs, a.colors_precomp = _ptr(shs), _ptr(colors_precomp)
    a.scales, a.rotations, a.cov3D_precomp = _ptr(scales), _ptr(rotations), _ptr(cov3D_precomp)
    a.scale_modifier = float(scale_modifier)
    keep = []
    _fill_views(a.views, state.views, dev, keep)
    a.radii, a.out_alpha = radii.data_ptr(), out_alpha.data_ptr()
    a.dL_dcolor, a.dL_ddepth, a.dL_dalpha = dL_dcolor.data_ptr(), dL_ddepth.data_ptr(), dL_dalpha.data_ptr()
    a.debug, a.sum_views = int(bool(debug)), int(bool(sum_views))
    a.dL_dmeans2D, a.dL_dcolors = g["means2D"].data_ptr(), g["colors"].data_ptr()
    a.dL_dopacity, a.dL_dmeans3D = g["opacity"].data_ptr(), g["means3D"].data_ptr()
    a.dL_dcov3D = g["cov3D"].data_ptr()
    a.dL_dsh = _ptr(g["sh"]) if shs is not None else None
    a.dL_dscales = g["scales"].data_ptr() if scales is not None else None
    a.dL_drotations = g["rotations"].data_ptr() if scales is not None else None
    a.dL_dconic = g["conic"].data_ptr() if want_aux else None
    a.dL_ddepths = g["depths"].data_ptr() if want_aux else None
    a.geom_buffer, a.geom_bytes = state.geom.data_ptr(), state.geom.numel()
    a.binning_buffer, a.binning_bytes = state.binning.data_ptr(), state.binning.numel()
    a.img_buffer, a.img_bytes = state.img.data_ptr(), state.img.numel()
    a.max_rendered = state.cap
    with torch.cuda.device(dev):
        rc = lib.gd_raster_backward(ctypes.byref(a), torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"gd_raster_backward failed ({rc}): {_lib.last_error(lib)}")
    return g


def mark_visible(positions, viewmatrix, projmatrix):
    lib = _lib.raster_lib()
    if not positions.is_cuda:
        raise RuntimeError("garmentdreamer_b200 rasteriser is CUDA-only (no CPU fallback)")
    dev = positions.device
    P = positions.shape[0]
    present = torch.zeros((P,), dtype=torch.bool, device=dev)
    if P:
        pos, vm, pm = _f32c(positions, dev), _f32c(viewmatrix, dev), _f32c(projmatrix, dev)
        with torch.cuda.device(dev):
            rc = lib.gd_mark_visible(P, pos.data_ptr(), vm.data_ptr(), pm.data_ptr(),
                                     present.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"gd_mark_visible failed: {_lib.last_error(lib)}")
    return present


def inspect_state(state: RasterState):
    """Copies the internal state to host tensors, in the reference's GeometryState/BinningState/
    ImageState vocabulary (rasterizer_impl.h:33-67). Test/debug helper; synchronises."""
    lib = _lib.raster_lib()
    sv = _lib.GdStateView()
    lib.gd_raster_state_view(state.P, state.W, state.H, state.B, state.cap, state.geom.data_ptr(),
                             state.binning.data_ptr(), state.img.data_ptr(), ctypes.byref(sv))
    if state.num_rendered is None:
        read_counters(state)
    P, B, W, H, R = state.P, state.B, state.W, state.H, state.num_rendered
    T = ((W + 15) // 16) * ((H + 15) // 16)

    def grab(buf, ptr, nbytes, dtype):
        off = ptr - buf.data_ptr()
        return buf[off:off + nbytes].view(dtype).cpu()

    rec = grab(state.geom, sv.records, B * P * 48, torch.float32).view(B, P, 12)
    out = {
        "conic_opacity": rec[..., 0:4].clone(),
        "means2D": rec[..., 4:6].clone(),
        "depths": rec[..., 6].clone(),
        "rgb": torch.stack([rec[..., 7], rec[..., 8], rec[..., 9]], -1),
        "tiles_touched": grab(state.geom, sv.tiles_touched, B * P * 4, torch.int32).view(B, P),
        "point_offsets": grab(state.geom, sv.point_offsets, B * P * 4, torch.int32).view(B, P),
        "cov3D": grab(state.geom, sv.cov3D, P * 24, torch.float32).view(P, 6),
        "clamped": grab(state.geom, sv.clamped, B * P, torch.uint8).view(B, P),
        "ranges": grab(state.img, sv.ranges, B * T * 8, torch.int32).view(B, T, 2),
        "n_contrib": grab(state.img, sv.n_contrib, B * H * W * 4, torch.int32).view(B, H, W),
        "num_rendered": R,
        "view_base": list(state.view_base),
    }
    if R:
        out["point_list"] = grab(state.binning, sv.point_list, R * 4, torch.int32)
        out["tile_keys"] = grab(state.binning, sv.tile_keys, R * 8, torch.int64)
        out["instance_slot"] = grab(state.binning, sv.instance_slot, R * 4, torch.int32)
    else:
        out["point_list"] = torch.zeros(0, dtype=torch.int32)
        out["tile_keys"] = torch.zeros(0, dtype=torch.int64)
        out["instance_slot"] = torch.zeros(0, dtype=torch.int32)
    return out


class RasterizeViews(torch.autograd.Function):
    """B views in one launch set; gradients summed over views (what autograd accumulates over the
    reference's per-view loop, GaussianDreamer.py:189-219). Inputs as rasterize_gaussians."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, views, W, H, bg, scale_modifier, sh_degree, sync):
        color, depth, alpha, radii, state = forward_views(
            means3D, opacities, views, W, H, bg, shs=sh, colors_precomp=colors_precomp,
            scales=scales, rotations=rotations, cov3D_precomp=cov3Ds_precomp,
            scale_modifier=scale_modifier, sh_degree=sh_degree, sync=sync)
        ctx.state, ctx.scale_modifier = state, scale_modifier
        ctx.save_for_backward(means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii,
                              alpha, bg)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha):
        means3D, sh, colors_precomp, scales, rotations, cov3Ds, radii, alpha, bg = ctx.saved_tensors
        st = ctx.state
        z = lambda ref, g: torch.zeros_like(ref) if g is None else g
        color_shape = (st.B, 3, st.H, st.W)
        g_color = torch.zeros(color_shape, device=means3D.device) if g_color is None else g_color
        g = backward_views(st, means3D, radii, alpha, bg, g_color, z(alpha, g_depth), z(alpha, g_alpha),
                           shs=sh, colors_precomp=colors_precomp, scales=scales, rotations=rotations,
                           cov3D_precomp=cov3Ds, scale_modifier=ctx.scale_modifier, sum_views=True)
        none_if_empty = lambda t, gr: gr if (t is not None and t.numel() > 0) else None
        return (g["means3D"], g["means2D"], none_if_empty(sh, g["sh"]),
                none_if_empty(colors_precomp, g["colors"]), g["opacity"],
                none_if_empty(scales, g["scales"]), none_if_empty(rotations, g["rotations"]),
                none_if_empty(cov3Ds, g["cov3D"]), None, None, None, None, None, None, None)


def rasterize_views(means3D, means2D, opacities, views, W, H, bg, *, shs=None, colors_precomp=None,
                    scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0,
                    sh_degree=0, sync=True):
    """Batched counterpart of GaussianRasterizer.forward: returns (color[B,3,H,W], radii[B,P],
    depth[B,1,H,W], alpha[B,1,H,W])."""
    e = torch.empty(0, device=means3D.device)
    return RasterizeViews.apply(means3D, means2D, e if shs is None else shs,
                                e if colors_precomp is None else colors_precomp, opacities,
                                e if scales is None else scales, e if rotations is None else rotations,
                                e if cov3D_precomp is None else cov3D_precomp, list(views), W, H, bg,
                                scale_modifier, sh_degree, sync)
