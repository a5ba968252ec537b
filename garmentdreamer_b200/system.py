"""The reference's optimisation iteration on the B200 kernels, without Lightning.

Mirror of threestudio ``GaussianDreamer`` for the per-iteration path
(Garment_3DGS/threestudio/systems/GaussianDreamer.py):

    forward(batch)                :179-219  cameras -> per-view render -> comp_rgb / depth / opacity
    training_step(batch)          :229-262  guidance (SDS) + sparsity loss on depth-normalised opacity
    on_before_optimizer_step()    :266-283  viewspace-gradient / radii statistics for densification
    optimizer.step()                        Adam(eps=1e-15) over the Gaussian parameter groups
                                            (GS/scene/gaussian_model.py:140-186)

Here one ``training_step`` is ONE explicit chain of kernels (no autograd tape):

    gd_cameras_from_c2w  ->  gd_params_activate  ->  gd_raster_forward (B views, one launch set)
      ->  VAE encode -> compute_grad_sds (UNet batch 2B) -> VAE input-gradient backward  = dL/dcolour
      ->  gd_sparsity_grad / gd_sparsity_finish                                        = dL/ddepth
      ->  gd_raster_backward (view-summed, straight into the packed [17P] buffer: 14P parameter
          gradients + 3P viewspace gradients)
      ->  (views sharded over ranks) all-reduce SUM of that buffer, MAX of the radii
      ->  gd_densify_stats  ->  gd_params_adam

Densify / prune every 100 steps (:281-283) changes P and is outside the hot path;
``resize()`` re-allocates the per-P buffers so a caller can do it between steps.
Multi-GPU: views are sharded over ranks, Gaussians / networks / Adam state replicated; the loss
is the reference's loss over the GLOBAL batch (1 / (B * world) on the SDS term, the mean of the
sparsity term over all views, depths.max() over all views).
"""
import contextlib
import ctypes
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib, parallel, raster
from .cameras import cameras_from_c2w
from .gaussians import GaussianParams, _chk, _lib_params


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


class SparsityLoss:
    """loss_sparsity = lambda * mean(sqrt(opacity^2 + 0.01)), opacity = depths / (depths.max() + 1e-5)
    (GaussianDreamer.py:215,253-255) and its gradient w.r.t. the depth images; libgd_raster.so kernels."""

    def __init__(self, device):
        self.device = device
        self.stats = torch.zeros(3, device=device)
        self.loss = torch.zeros(1, device=device)
        self.dmax = torch.zeros(1, device=device)
        self._scratch = None
        L = _lib.raster_lib()
        vp, ll, f = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_float
        L.gd_sparsity_grad.argtypes = [ll, ll, vp, vp, f, vp, vp, vp, vp]
        L.gd_sparsity_finish.argtypes = [ll, ll, vp, vp, f, vp, vp, vp, vp]
        L.gd_sparsity_grad.restype = L.gd_sparsity_finish.restype = ctypes.c_int
        self.L = L

    def depth_max_of(self, state: raster.RasterState) -> torch.Tensor:
        """Device scalar: max over the B depth images, left in GdCounters by the forward compositor."""
        sv = _lib.GdStateView()
        self.L.gd_raster_state_view(state.P, state.W, state.H, state.B, state.cap, state.geom.data_ptr(),
                                    state.binning.data_ptr(), state.img.data_ptr(), ctypes.byref(sv))
        off = sv.counters - state.geom.data_ptr() + _lib.GdCounters.depth_max_bits.offset
        self.dmax.copy_(state.geom[off:off + 4].view(torch.float32))
        return self.dmax

    def grad(self, depth: torch.Tensor, dmax: torch.Tensor, lam: float, n_total: int, group=None):
        """depth [B,1,H,W] of this rank, dmax device scalar (already the global max). Returns dL/ddepth
        (same shape); self.loss holds the loss value (device scalar)."""
        n = depth.numel()
        nblk = (n + 1023) // 1024
        if self._scratch is None or self._scratch.numel() < 3 * nblk:
            self._scratch = torch.empty(3 * nblk, device=depth.device)
        out = torch.empty_like(depth)
        st = torch.cuda.current_stream().cuda_stream
        _chk(self.L.gd_sparsity_grad(n, n_total, depth.data_ptr(), dmax.data_ptr(), float(lam), out.data_ptr(),
                                     self._scratch.data_ptr(), self.stats.data_ptr(), st), "gd_sparsity_grad")
        if _world(group) > 1:
            dist.all_reduce(self.stats, op=dist.ReduceOp.SUM, group=group)
        _chk(self.L.gd_sparsity_finish(n, n_total, depth.data_ptr(), dmax.data_ptr(), float(lam), self.stats.data_ptr(),
                                       out.data_ptr(), self.loss.data_ptr(), st), "gd_sparsity_finish")
        return out


class GaussianDreamerB200:
    def __init__(self, gaussian: GaussianParams, guidance=None, *, lambda_sds=1.0, lambda_sparsity=1.0,
                 background=(1.0, 1.0, 1.0), group=None, adam_betas=(0.9, 0.999), adam_eps=1e-15):
        self.gaussian = gaussian
        self.guidance = guidance          # object with image_grad(color, elevation, azimuth, distances, scale) or None
        self.lambda_sds, self.lambda_sparsity = float(lambda_sds), float(lambda_sparsity)
        self.group = group
        self.dev = gaussian._xyz.device
        self.background_tensor = torch.tensor(background, dtype=torch.float32, device=self.dev)
        self.sparsity = SparsityLoss(self.dev)
        self.adam_betas, self.adam_eps = adam_betas, adam_eps
        self.true_global_step = 0
        self.timers = None                # set to {} to collect CUDA events per phase
        self.resize()

    def resize(self):
        """(Re)allocate everything sized by P -- call after densify / prune changed the Gaussians."""
        P = self.gaussian.P
        self.packed = torch.empty(14 * P, device=self.dev)           # activated parameters
        self.peers = None
        if _world(self.group) > 1 and parallel.peer_exchange_enabled():
            # gradients and radii are written straight into a symmetric allocation every peer can read (NVLink);
            # the reduction is fused with the optimiser step (parallel.PeerExchange). GD_PEER_REDUCE=0: NCCL all-reduce.
            try:
                self.peers = parallel.PeerExchange(P, self.dev, self.group)
            except Exception as e:   # noqa: BLE001 -- no symmetric memory on this box: the NCCL path below
                import warnings
                warnings.warn(f"peer exchange unavailable ({e}); falling back to NCCL all-reduce")
        if self.peers is not None:
            self.grad, self.radii_max = self.peers.grad, self.peers.radii
        else:
            self.grad = torch.empty(17 * P, device=self.dev)         # 14P parameter gradients | 3P viewspace gradients
            self.radii_max = torch.zeros(P, dtype=torch.int32, device=self.dev)
        self._P = P
        self._arena_key = None     # (P, W, H, B) whose instance-arena capacity has been measured
        self._watch = None         # (pinned copy of the device counters, event, key) of the previous step

    def _mark(self, name):
        if self.timers is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.timers.setdefault(name, []).append(e)

    # ---- GaussianDreamer.forward (:179-219) ---------------------------------------------------------
    def forward(self, batch: Dict, renderbackground: Optional[torch.Tensor] = None):
        if self.gaussian.P != self._P:
            self.resize()
        bg = self.background_tensor if renderbackground is None else renderbackground
        P = self._P
        H, W = int(batch["height"]), int(batch["width"])
        c2w = batch["c2w_3dgs"]
        if not c2w.is_cuda:
            c2w = c2w.to(self.dev, non_blocking=True)
        views, cam_packed = cameras_from_c2w(c2w, batch["fovy"], H, W)
        self.gaussian.activated(out=self.packed)
        xyz, shs, op, sc, rot = GaussianParams.unpack(self.packed, P)
        # Instance arena: the first step of a shape reads the count back (like the reference's blocking cudaMemcpy,
        # rasterizer_impl.cu:282) and sizes the arena with headroom; later steps never touch the host. Their device
        # counters are copied to pinned memory and inspected one step LATE (no stall): on overflow (the Gaussians grew)
        # that step rendered blank views and contributed zero gradients -- grow and re-measure.
        key = (P, W, H, len(views))
        if self._watch is not None and self._watch[1].query():
            host, _, wkey = self._watch
            self._watch = None
            if int(host[1]) != 0:
                raster._cap_hint[wkey] = int((int(host[0]) & 0xFFFFFFFF) * 1.5) + 4096
                self._arena_key = None
        first = self._arena_key != key
        color, depth, alpha, radii, st = raster.forward_views(xyz, op, views, W, H, bg, shs=shs, scales=sc, rotations=rot,
                                                              sync=first)
        if first:
            raster._cap_hint[key] = int(raster._cap_hint[key] * 1.2)
            self._arena_key = key
        elif self._watch is None:
            sv = _lib.GdStateView()
            _lib.raster_lib().gd_raster_state_view(st.P, st.W, st.H, st.B, st.cap, st.geom.data_ptr(), st.binning.data_ptr(),
                                                   st.img.data_ptr(), ctypes.byref(sv))
            off = sv.counters - st.geom.data_ptr()
            host = getattr(self, "_watch_host", None)
            if host is None:
                host = self._watch_host = torch.zeros(2, dtype=torch.int32).pin_memory()
            host.copy_(st.geom[off:off + 8].view(torch.int32), non_blocking=True)
            e = torch.cuda.Event()
            e.record()
            self._watch = (host, e, key)
        self._fw = (views, cam_packed, st, radii, alpha, bg)
        return {"render": color, "comp_rgb": color.permute(0, 2, 3, 1), "depth": depth.permute(0, 2, 3, 1),
                "alphas": alpha.permute(0, 2, 3, 1), "radii": radii, "state": st, "depth_3dgs": depth}

    # ---- training_step + on_before_optimizer_step + optimizer.step ------------------------------------
    def training_step(self, batch: Dict, batch_idx: int = 0):
        g = self.gaussian
        world = _world(self.group)
        self._mark("start")
        g.update_learning_rate(self.true_global_step)
        if self.true_global_step > 500 and self.guidance is not None and hasattr(self.guidance, "set_min_max_steps"):
            self.guidance.set_min_max_steps(min_step_percent=0.02, max_step_percent=0.55)   # :233-234
        out = self.forward(batch)
        self._mark("raster_fwd")
        views, _, st, radii, alpha, bg = self._fw
        color, depth = out["render"], out["depth_3dgs"]
        B, _, H, W = color.shape
        P = self._P
        # loss_sparsity on opacity = depths / (depths.max() + 1e-5): max over the WHOLE batch (:215).
        # Multi-rank: the two scalar exchanges it needs (MAX of the depth maxima, SUM of three partial sums) are only
        # consumed by the raster BACKWARD, a whole guidance call later -- so they run on a side stream under the VAE / UNet
        # instead of stalling every rank on the slowest rasteriser forward twice per step (measured at N = 8: 0.96 ms of a
        # 28.3 ms step). The main stream joins the side stream right before the backward.
        side = None
        if world > 1:
            side = getattr(self, "_side", None)
            if side is None:
                side = self._side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
            dmax = self.sparsity.depth_max_of(st)
            if world > 1:
                dist.all_reduce(dmax, op=dist.ReduceOp.MAX, group=self.group)
            dL_ddepth = self.sparsity.grad(depth, dmax, self.lambda_sparsity, depth.numel() * world, self.group)
            if side is not None:
                depth.record_stream(side)
        self._mark("sparsity")
        # loss_sds: 0.5 * mse(latents, target, 'sum') / batch_size over the global batch (:424-427)
        if self.guidance is not None:
            dL_dcolor = self.guidance.image_grad(color, batch["elevation"], batch["azimuth"], batch["camera_distances"],
                                                 scale=self.lambda_sds / (B * world))
        else:
            dL_dcolor = batch["dL_dcolor"]
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            dL_ddepth.record_stream(torch.cuda.current_stream())   # allocated under the side stream, consumed here
        self._mark("guidance")
        zeros = getattr(self, "_zeros", None)
        if zeros is None or zeros.shape != alpha.shape:
            zeros = self._zeros = torch.zeros_like(alpha)          # the loss does not touch the alpha image
        xyz, shs, op, sc, rot = GaussianParams.unpack(self.packed, P)
        o3, osh, oop, osc, orot = GaussianParams.unpack(self.grad[:14 * P], P)
        raster.backward_views(st, xyz, radii, alpha, bg, dL_dcolor, dL_ddepth, zeros, shs=shs, scales=sc, rotations=rot,
                              sum_views=True, out={"means3D": o3, "sh": osh, "opacity": oop, "scales": osc, "rotations": orot,
                                                   "means2D": self.grad[14 * P:].view(P, 3)})
        self._mark("raster_bwd")
        L = _lib_params()
        L.gd_radii_max.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.gd_radii_max.restype = ctypes.c_int
        stream = torch.cuda.current_stream().cuda_stream
        _chk(L.gd_radii_max(P, B, radii.data_ptr(), self.radii_max.data_ptr(), stream), "gd_radii_max")
        if self.peers is not None:
            # the only data-path exchange: per-Gaussian gradients (SUM) and radii (MAX) through peer memory, then
            # on_before_optimizer_step (:266-279) + optimizer.step() in the kernel that waits for the last slice
            self.peers.allreduce()
            self._mark("allreduce")
            g.adam_step_peers(self.peers, densify=self.true_global_step < 900, beta1=self.adam_betas[0], beta2=self.adam_betas[1],
                              eps=self.adam_eps)
            self._mark("adam")
            self.true_global_step += 1
            return {"loss_sparsity": self.sparsity.loss, "state": st}
        if world > 1:   # NCCL fallback: per-Gaussian gradients (SUM) and radii (MAX)
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.radii_max, op=dist.ReduceOp.MAX, group=self.group)
        self._mark("allreduce")
        # on_before_optimizer_step (:266-279)
        if self.true_global_step < 900:
            g.add_densification_stats(self.grad[14 * P:].view(P, 3), self.radii_max.view(1, P))
        g.adam_step(self.grad[:14 * P], beta1=self.adam_betas[0], beta2=self.adam_betas[1], eps=self.adam_eps)
        self._mark("adam")
        self.true_global_step += 1
        return {"loss_sparsity": self.sparsity.loss, "state": st}

    def phase_ms(self):
        """Mean ms per phase from the collected events (call after a synchronize)."""
        names = ["raster_fwd", "sparsity", "guidance", "raster_bwd", "allreduce", "adam"]
        prev, out = "start", {}
        for nme in names:
            ev0, ev1 = self.timers.get(prev, []), self.timers.get(nme, [])
            k = min(len(ev0), len(ev1))
            out[nme] = float(sum(a.elapsed_time(b) for a, b in zip(ev0[:k], ev1[:k])) / max(1, k))
            prev = nme
        return out
