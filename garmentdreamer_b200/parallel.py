"""View-sharded data parallelism for the SDS step (SURVEY.md s.8 row e).

The reference is single-GPU (generate_3dgs.py:40,58). Here each rank renders its slice of the
view batch against replicated Gaussians; the per-Gaussian gradients (what autograd sums over the
reference's per-view loop, GaussianDreamer.py:189-219) are summed over ranks with ONE all-reduce
of the packed [14*P] buffer. No other data-path collective exists. Works with any
torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
import ctypes
import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

PACK = (("means3D", 3), ("sh", 3), ("opacity", 1), ("scales", 3), ("rotations", 4))  # 14 floats / Gaussian


def shard_views(n_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the global view batch owned by `rank` (n_views % world == 0)."""
    if n_views % world:
        raise ValueError(f"{n_views} views do not shard evenly over {world} ranks")
    per = n_views // world
    return rank * per, (rank + 1) * per


def pack_layout(P: int):
    """(name, offset, numel, shape) of each gradient inside the flat struct-of-arrays buffer."""
    out, off = [], 0
    for name, c in PACK:
        shape = (P, 1, 3) if name == "sh" else (P, c)
        out.append((name, off, P * c, shape))
        off += P * c
    return out, off


def unpack(flat: torch.Tensor, P: int):
    lay, total = pack_layout(P)
    assert flat.numel() == total
    return {name: flat[off:off + n].view(shape) for name, off, n, shape in lay}


def allreduce_gradients(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the packed gradient over ranks (in place). A no-op for world size 1."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def allreduce_depth_max(depth_max: torch.Tensor, group=None) -> torch.Tensor:
    """`opacity = depths / (depths.max() + 1e-5)` takes the max over the WHOLE view batch
    (GaussianDreamer.py:215): ranks exchange one scalar before backward."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(depth_max, op=dist.ReduceOp.MAX, group=group)
    return depth_max


def allreduce_densification_stats(viewspace_grad: torch.Tensor, radii: torch.Tensor, group=None):
    """SUM of viewspace-point gradients, MAX of radii over ranks (GaussianDreamer.py:195-198,273-279)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(viewspace_grad, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(radii, op=dist.ReduceOp.MAX, group=group)
    return viewspace_grad, radii


# ---- gradient exchange over peer memory, fused with the optimiser step (include/gd_raster.h, GdPeerTable) -------------
GD_MAX_PEERS = 8


class GdPeerTable(ctypes.Structure):
    _fields_ = [("world", ctypes.c_int), ("rank", ctypes.c_int),
                ("grad", ctypes.c_void_p * GD_MAX_PEERS), ("radii", ctypes.c_void_p * GD_MAX_PEERS),
                ("red_grad", ctypes.c_void_p * GD_MAX_PEERS), ("red_radii", ctypes.c_void_p * GD_MAX_PEERS),
                ("flags", ctypes.c_void_p * GD_MAX_PEERS), ("mc_grad", ctypes.c_void_p), ("mc_red_grad", ctypes.c_void_p)]


def peer_exchange_enabled() -> bool:
    return os.environ.get("GD_PEER_REDUCE", "1") != "0"


class PeerExchange:
    """One symmetric allocation per rank (torch.distributed._symmetric_memory: cuMem + fabric / fd handles, mapped into
    every peer over NVLink) holding this rank's raster-backward outputs and the reduced copies:

        grad [17P pad 4] f32 | radii [P pad 4] i32 | red_grad | red_radii | flags [2][8] u32

    `grad` / `radii` are the buffers gd_raster_backward / gd_radii_max write; `allreduce()` launches gd_peer_allreduce
    (reduce-scatter + all-gather through peer loads / stores, or NVLS multimem when the allocation has a multicast
    address), after which GaussianParams.adam_step_peers() consumes `red_grad` / `red_radii`. No NCCL on this path."""

    def __init__(self, P: int, device, group=None, use_multicast=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.P, self.device = P, torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > GD_MAX_PEERS:
            raise RuntimeError(f"peer exchange supports up to {GD_MAX_PEERS} ranks of one NVLink domain")
        g = (17 * P + 3) // 4 * 4
        r = (P + 3) // 4 * 4
        self._off = {"grad": 0, "radii": 4 * g, "red_grad": 4 * (g + r), "red_radii": 4 * (2 * g + r), "flags": 4 * (2 * g + 2 * r)}
        nbytes = self._off["flags"] + 4 * 2 * GD_MAX_PEERS
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        ptrs = list(self.hdl.buffer_ptrs)
        mc = 0
        if use_multicast is None:
            use_multicast = os.environ.get("GD_PEER_MULTICAST", "1") != "0"
        if use_multicast and self.world > 1:
            try:
                mc = int(self.hdl.multicast_ptr or 0)   # 0 / None when the fabric has no NVLS multicast
            except Exception:   # noqa: BLE001
                mc = 0
        self.multicast = mc != 0
        t = GdPeerTable()
        t.world, t.rank = self.world, self.rank
        for w in range(self.world):
            for name in ("grad", "radii", "red_grad", "red_radii", "flags"):
                getattr(t, name)[w] = ptrs[w] + self._off[name]
        t.mc_grad = (mc + self._off["grad"]) if mc else None
        t.mc_red_grad = (mc + self._off["red_grad"]) if mc else None
        self.table = t
        f32 = lambda name, n: self.buf[self._off[name]:self._off[name] + 4 * n].view(torch.float32)
        i32 = lambda name, n: self.buf[self._off[name]:self._off[name] + 4 * n].view(torch.int32)
        self.grad, self.red_grad = f32("grad", 17 * P), f32("red_grad", 17 * P)
        self.radii, self.red_radii = i32("radii", P), i32("red_radii", P)
        self.counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = 0
        L = _lib.raster_lib()
        vp, i, u, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_float
        L.gd_peer_allreduce.argtypes = [i, ctypes.POINTER(GdPeerTable), u, vp, vp]
        L.gd_peer_allreduce.restype = ctypes.c_int
        L.gd_params_adam_peers.argtypes = [i, vp, vp, vp, vp, vp, ctypes.POINTER(GdPeerTable), u, vp, vp, ctypes.POINTER(f), f, f, f, i, i,
                                           vp, vp, vp, vp]
        L.gd_params_adam_peers.restype = ctypes.c_int
        self.L = L
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)   # every rank's flags are zeroed before anyone signals

    def allreduce(self):
        """SUM of `grad`, MAX of `radii` over the ranks into every rank's red_grad / red_radii (asynchronous)."""
        self.epoch += 1
        rc = self.L.gd_peer_allreduce(self.P, ctypes.byref(self.table), self.epoch, self.counter.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"gd_peer_allreduce failed ({rc}): {self.L.gd_last_error().decode()}")
