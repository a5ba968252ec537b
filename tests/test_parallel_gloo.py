"""CPU, world_size 2 over gloo: view sharding + packed-gradient all-reduce reproduce the
single-process sum over the whole view batch (the multi-GPU oracle of SURVEY.md s.8 row e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from garmentdreamer_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _per_view_grad(view, P):
    g = torch.Generator().manual_seed(1000 + view)
    return torch.randn(14 * P, generator=g, dtype=torch.float64).float()


def _worker(rank, world, port, P, n_views, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_views(n_views, rank, world)
    flat = torch.zeros(14 * P)
    for v in range(lo, hi):
        flat += _per_view_grad(v, P)
    parallel.allreduce_gradients(flat)
    dmax = parallel.allreduce_depth_max(torch.tensor([float(rank + 1)]))
    vs, rad = parallel.allreduce_densification_stats(torch.full((P, 3), float(rank + 1)), torch.arange(P, dtype=torch.int32) * (rank + 1))
    if rank == 0:
        torch.save({"flat": flat, "dmax": dmax, "vs": vs, "rad": rad}, out)
    dist.destroy_process_group()


def test_view_sharding_allreduce_matches_single_process(tmp_path):
    P, n_views, world = 257, 8, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), P, n_views, out), nprocs=world, join=True)
    got = torch.load(out)
    ref = sum(_per_view_grad(v, P) for v in range(n_views))
    assert torch.allclose(got["flat"], ref, atol=1e-5)
    assert float(got["dmax"]) == 2.0 and float(got["vs"][0, 0]) == 3.0
    assert torch.equal(got["rad"], torch.arange(P, dtype=torch.int32) * 2)
    parts = parallel.unpack(got["flat"], P)
    assert parts["means3D"].shape == (P, 3) and parts["sh"].shape == (P, 1, 3) and parts["rotations"].shape == (P, 4)


def test_shard_views_partition():
    for world in (1, 2, 4, 8):
        spans = [parallel.shard_views(32, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 32
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    import pytest
    with pytest.raises(ValueError):
        parallel.shard_views(6, 0, 4)
