"""torchrun --nproc-per-node N tools/peer_exchange_check.py: the peer-memory gradient exchange (parallel.PeerExchange:
gd_peer_allreduce + gd_params_adam_peers) against NCCL all-reduce + gd_densify_stats + gd_params_adam on the same inputs,
and its latency next to NCCL's. Run under `gpurun --gpus N`; output kept under profiles/."""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garmentdreamer_b200 import parallel
from garmentdreamer_b200.gaussians import GaussianParams
from garmentdreamer_b200.synthetic import garment, raw_params

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
for P in (100000, 500000, 7001):
    raw = {k: v.to(dev) for k, v in raw_params(garment(P, 0)).items()}
    mk = lambda: GaussianParams(raw["xyz"], raw["f_dc"], raw["opacity"], raw["scaling"], raw["rotation"], spatial_lr_scale=4.0)
    for mc in ((True, False) if world > 1 else (False,)):
        ga, gb = mk(), mk()
        ga.training_setup(); gb.training_setup()
        px = parallel.PeerExchange(P, dev, use_multicast=mc)
        if mc and not px.multicast:
            out[f"P{P}_multicast"] = "unsupported"
            continue
        tag = f"P{P}_{'nvls' if px.multicast else 'p2p'}"
        worst = 0.0
        for step in range(3):
            g = torch.Generator(device=dev).manual_seed(100 * step + rank)
            grad = torch.randn(17 * P, generator=g, device=dev) * 1e-3
            radii = torch.randint(0, 40, (P,), generator=g, device=dev, dtype=torch.int32)
            px.grad.copy_(grad); px.radii.copy_(radii)
            ref_g, ref_r = grad.clone(), radii.clone()
            dist.all_reduce(ref_g); dist.all_reduce(ref_r, op=dist.ReduceOp.MAX)
            ga.update_learning_rate(step); gb.update_learning_rate(step)
            ga.add_densification_stats(ref_g[14 * P:].view(P, 3), ref_r.view(1, P))
            ga.adam_step(ref_g[:14 * P])
            px.allreduce()
            gb.adam_step_peers(px)
            torch.cuda.synchronize()
            assert torch.equal(px.red_radii, ref_r), "radii MAX differs from NCCL"
            err = float((px.red_grad - ref_g).abs().max() / ref_g.abs().max())
            worst = max(worst, err)
            assert err < 1e-6, f"gradient SUM differs from NCCL by {err}"
            for a, b in ((ga._xyz, gb._xyz), (ga._rotation, gb._rotation), (ga.exp_avg, gb.exp_avg), (ga.exp_avg_sq, gb.exp_avg_sq),
                         (ga.xyz_gradient_accum, gb.xyz_gradient_accum), (ga.denom, gb.denom), (ga.max_radii2D, gb.max_radii2D)):
                assert torch.allclose(a, b, rtol=1e-4, atol=1e-7), "optimiser state differs from the NCCL path"
        # replicas must be bit-identical across ranks
        chk = torch.stack([gb._xyz.double().sum(), gb.exp_avg.double().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "replicas diverged"
        # latency: exchange + optimiser, both paths, after a barrier
        ev = lambda: torch.cuda.Event(enable_timing=True)
        def timeit(fn, n=20):
            for _ in range(3): fn()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = ev(), ev(); e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e3
        def nccl_path():
            dist.all_reduce(ref_g); dist.all_reduce(ref_r, op=dist.ReduceOp.MAX)   # (values grow: timing only)
            ga.add_densification_stats(ref_g[14 * P:].view(P, 3), ref_r.view(1, P)); ga.adam_step(ref_g[:14 * P])
        def peer_path():
            px.allreduce(); gb.adam_step_peers(px)
        out[tag] = {"max_rel_err_vs_nccl": worst, "peer_us": timeit(peer_path), "nccl_us": timeit(nccl_path)}
        del px
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
