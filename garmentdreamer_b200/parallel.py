"""View-sharded data parallelism for the SDS step (SURVEY.md s.8 row e).

The reference is single-GPU (generate_3dgs.py:40,58). Here each rank renders its slice of the
view batch against replicated Gaussians; the per-Gaussian gradients (what autograd sums over the
reference's per-view loop, GaussianDreamer.py:189-219) are summed over ranks with ONE all-reduce
of the packed [14*P] buffer. No other data-path collective exists. Works with any
torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
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
