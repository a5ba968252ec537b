"""Drop-in replacement for the reference Python package ``diff_gaussian_rasterization``.

Same names, argument order, return order, dtypes, shapes and error behaviour as
Garment_3DGS/gaussiansplatting/submodules/diff-gaussian-rasterization/
diff_gaussian_rasterization/__init__.py (rasterize_gaussians :21-42, _RasterizeGaussians :44-158,
GaussianRasterizationSettings :160-172, GaussianRasterizer :174-223), so that
Garment_3DGS/gaussiansplatting/gaussian_renderer/__init__.py:14 imports it unchanged.
The native module ``_C`` is a shim over the B200 C-ABI library (include/gd_raster.h) with the
reference's pybind signatures (ext.cpp:15-18). CUDA only: no CPU fallback.

Extension (not in the reference): ``rasterize_views`` renders B views in one launch set.
"""
from collections import OrderedDict
from typing import NamedTuple

import torch
import torch.nn as nn

from garmentdreamer_b200 import raster as _raster
from garmentdreamer_b200.raster import View, rasterize_views  # noqa: F401  (batched extension)


class _NativeShim:
    """``_C``-shaped facade: rasterize_gaussians / rasterize_gaussians_backward / mark_visible."""

    # Raw `_C` callers (the reference's pybind usage) look their state up by the geomBuffer they got back; the
    # autograd Function below takes the entry out right after the forward and keeps it on its ctx, so forwards
    # under no_grad (validation, the 120 test views, markVisible loops) hold nothing and a second backward
    # (retain_graph=True) still finds its state.
    _MAX_LIVE = 8

    def __init__(self):
        self._states = OrderedDict()  # geomBuffer.data_ptr() -> RasterState

    def take_state(self, geomBuffer):
        return self._states.pop(geomBuffer.data_ptr(), None)

    def put_state(self, state):
        self._states[state.geom.data_ptr()] = state
        while len(self._states) > self._MAX_LIVE:
            self._states.popitem(last=False)

    def rasterize_gaussians(self, bg, means3D, colors, opacity, scales, rotations, scale_modifier,
                            cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                            image_width, sh, degree, campos, prefiltered, debug):
        view = View(viewmatrix, projmatrix, campos, tan_fovx, tan_fovy)
        color, depth, alpha, radii, state = _raster.forward_views(
            means3D, opacity, [view], int(image_width), int(image_height), bg, shs=sh,
            colors_precomp=colors, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp,
            scale_modifier=scale_modifier, sh_degree=degree, prefiltered=prefiltered, debug=debug,
            sync=True)
        self.put_state(state)
        return (state.num_rendered, color[0], depth[0], alpha[0], radii[0], state.geom,
                state.binning, state.img)

    def rasterize_gaussians_backward(self, bg, means3D, radii, colors, scales, rotations,
                                     scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                                     tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth,
                                     dL_dout_alpha, sh, degree, campos, geomBuffer, R,
                                     binningBuffer, imageBuffer, alphas, debug):
        state = self._states.get(geomBuffer.data_ptr())
        if state is None or state.binning.data_ptr() != binningBuffer.data_ptr():
            raise RuntimeError("rasterize_gaussians_backward: unknown state buffers (they must be "
                               "the ones returned by rasterize_gaussians)")
        g = _raster.backward_views(
            state, means3D, radii.view(1, -1), alphas.view(1, 1, state.H, state.W), bg,
            dL_dout_color, dL_dout_depth, dL_dout_alpha, shs=sh, colors_precomp=colors,
            scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp,
            scale_modifier=scale_modifier, sum_views=False, debug=debug)
        return (g["means2D"][0], g["colors"][0], g["opacity"][0], g["means3D"][0], g["cov3D"][0],
                g["sh"][0], g["scales"][0], g["rotations"][0])

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        return _raster.mark_visible(means3D, viewmatrix, projmatrix)


_C = _NativeShim()


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(
    means3D,
    means2D,
    sh,
    colors_precomp,
    opacities,
    scales,
    rotations,
    cov3Ds_precomp,
    raster_settings,
):
    return _RasterizeGaussians.apply(
        means3D,
        means2D,
        sh,
        colors_precomp,
        opacities,
        scales,
        rotations,
        cov3Ds_precomp,
        raster_settings,
    )


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered,
                rs.debug)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy them before they can be corrupted
            try:
                out = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            out = _C.rasterize_gaussians(*args)
        num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer = out
        state = _C.take_state(geomBuffer)          # owned by this graph node from here on (nothing global stays alive)
        ctx.gd_state = state if any(ctx.needs_input_grad) else None
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                              geomBuffer, binningBuffer, imgBuffer, alpha)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        num_rendered = ctx.num_rendered
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
         binningBuffer, imgBuffer, alpha) = ctx.saved_tensors
        if grad_color is None:
            grad_color = torch.zeros((3,) + tuple(alpha.shape[1:]), device=alpha.device)
        if grad_depth is None:
            grad_depth = torch.zeros_like(alpha)
        if grad_alpha is None:
            grad_alpha = torch.zeros_like(alpha)
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_color,
                grad_depth, grad_alpha, sh, rs.sh_degree, rs.campos, geomBuffer, num_rendered,
                binningBuffer, imgBuffer, alpha, rs.debug)
        if ctx.gd_state is None:
            raise RuntimeError("rasterize_gaussians: backward called on a forward that needed no gradient")
        _C.put_state(ctx.gd_state)                 # visible to the _C-shaped entry point for the duration of the call
        try:
            return _RasterizeGaussians._backward_impl(ctx, args, colors_precomp, scales, rotations, cov3Ds_precomp, sh)
        finally:
            _C.take_state(geomBuffer)

    @staticmethod
    def _backward_impl(ctx, args, colors_precomp, scales, rotations, cov3Ds_precomp, sh):
        rs = ctx.raster_settings
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _C.rasterize_gaussians_backward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            out = _C.rasterize_gaussians_backward(*args)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp,
         grad_sh, grad_scales, grad_rotations) = out
        # the reference returns a gradient for every tensor input, including absent (empty) ones;
        # autograd needs None (or a matching empty tensor) for the empty CPU placeholders
        fix = lambda t, g: g if t.numel() > 0 else None
        grads = (
            grad_means3D,
            grad_means2D,
            fix(sh, grad_sh),
            fix(colors_precomp, grad_colors_precomp),
            grad_opacities,
            fix(scales, grad_scales),
            fix(rotations, grad_rotations),
            fix(cov3Ds_precomp, grad_cov3Ds_precomp),
            None,
        )
        return grads


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        # absent inputs travel as empty CPU tensors; the native side reads them as "not given"
        shs, colors_precomp, scales, rotations, cov3D_precomp = (
            torch.Tensor([]) if t is None else t
            for t in (shs, colors_precomp, scales, rotations, cov3D_precomp))

        # Invoke C++/CUDA rasterization routine
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales,
                                   rotations, cov3D_precomp, raster_settings)
