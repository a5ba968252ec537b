#!/usr/bin/env python
"""Contract benchmark: SDS iterations/s on the synthetic garment (BASELINE.json configs[1]:
100k Gaussians, 4 camera views 512^2 per GPU), one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference path on the host cores (oracle port)
    python bench.py --phase raster --P 500000 --views 8 --res 1024 --strong   # c5 rasteriser sweep

One "step" = ONE optimisation iteration of the reference (threestudio GaussianDreamer.training_step +
on_before_optimizer_step + optimizer.step, TS/systems/GaussianDreamer.py:229-283) over B views:
  cameras from c2w -> parameter activations -> rasterise forward (B views) -> VAE encode ->
  compute_grad_sds (UNet batch 2B) -> VAE input-gradient backward -> sparsity loss on the
  depth-normalised opacity -> rasterise backward -> (N > 1) NCCL all-reduce of the packed
  per-Gaussian gradients [P,17] and of the radii -> densification statistics -> Adam.
Views are sharded over ranks (weak scaling: B views per GPU); Gaussians, networks and the
optimiser state are replicated. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--P", type=int, default=100000)
    ap.add_argument("--views", type=int, default=4, help="views per GPU (with --strong: views of the whole job)")
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--phase", default="auto", choices=["auto", "raster", "sds"])
    ap.add_argument("--strong", action="store_true", help="fixed total view count, sharded over the ranks (c5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-vae", action="store_true", help="replace the VAE encoder by the linear stand-in")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 7 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 7 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4)
                          if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def algorithmic_bytes_bwd(P, R, N, T):
    """44*R + 28*N + 8*T + 187*P per view (R, N, T summed over views here)."""
    return 44 * R + 28 * N + 8 * T + 187 * P


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def unet_available():
    return os.path.exists(os.path.join(ROOT, "garmentdreamer_b200", "lib", "libgd_unet.so"))


def workload_config(a, world, with_unet):
    """`config` of the JSON line -- identical for our arm and the reference arm."""
    per_gpu = a.views // world if a.strong else a.views
    if with_unet and not a.no_vae:
        w = (f"c2: {a.P} Gaussians, {per_gpu}x{a.res}^2 views per GPU, full reference iteration: cameras + activations + raster fwd "
             f"+ VAE encode + SD-2.1 UNet SDS grad (batch {2 * per_gpu}) + VAE input-gradient bwd + sparsity loss + raster bwd "
             f"+ densification stats + Adam (random-init fp16 networks)")
    elif with_unet:
        w = (f"c2: {a.P} Gaussians, {per_gpu}x{a.res}^2 views per GPU, reference iteration with the VAE excluded "
             f"(latents = 8x8-pooled render): raster fwd + SD-2.1 UNet SDS grad (batch {2 * per_gpu}) + sparsity + raster bwd + Adam")
    else:
        w = (f"raster: {a.P} Gaussians, {per_gpu}x{a.res}^2 views per GPU, cameras + activations + rasterise fwd + sparsity loss "
             f"+ rasterise bwd (seeded upstream colour gradient) + densification stats + Adam; no UNet/VAE")
    return {"workload": w, "P": a.P, "views_per_gpu": per_gpu, "views_total": per_gpu * world, "res": a.res,
            "l2": "flushed between timed iterations (256 MiB write)",
            "unit_of_work": f"one SDS iteration = {per_gpu} views per rank; value = ranks x steps / max-over-ranks time"
                            if not a.strong else f"one iteration over the job's {per_gpu * world} views; value = steps / max-over-ranks time",
            "parallelism": f"views sharded x{world}, NCCL all-reduce of [P,17] grads + radii" if world > 1 else "single GPU"}


# ------------------------------------------------------------------------------------------
# CPU legs (the ONLY places of this file that execute oracle/ code, besides gpu_reference_leg)
def cpu_sample(a, with_unet):
    """A bounded sample of one iteration on the host cores: 1 of the B views rasterised fwd+bwd by the C
    oracle (OpenMP), the UNet restatement on ONE cond+uncond pair (batch 2), the VAE restatement encode +
    input gradient on ONE image -- i.e. exactly 1/B of the iteration's work. Returns (seconds, note)."""
    from garmentdreamer_b200.synthetic import garment, sample_cameras
    from oracle import raster_oracle as ro
    g = garment(a.P, 0)
    c = sample_cameras(a.views, a.res, a.res)[0]
    gen = torch.Generator().manual_seed(7)
    dc = torch.randn(3, a.res, a.res, generator=gen).numpy()
    dd = torch.randn(1, a.res, a.res, generator=gen).numpy()
    da = torch.randn(1, a.res, a.res, generator=gen).numpy()
    t0 = time.perf_counter()
    st = ro.forward(g["xyz"].numpy(), g["opacity"].numpy(), c.viewmatrix.numpy(), c.projmatrix.numpy(),
                    c.campos.numpy(), a.res, a.res, c.tanfovx, c.tanfovy, np.ones(3, np.float32),
                    shs=g["shs"].numpy(), scales=g["scales"].numpy(), rotations=g["rotations"].numpy())
    ro.backward(st, dc, dd, da)
    t_r = time.perf_counter() - t0
    t_u = t_v = 0.0
    if with_unet:
        from oracle import unet_ref
        torch.set_num_threads(os.cpu_count())
        t_u = unet_ref.time_cpu_forward(batch=2)
        if not a.no_vae:
            from oracle import vae_ref
            t_v = vae_ref.time_cpu_encode(batch=1, res=a.res)
    note = (f"1 of {a.views} views: C-oracle raster fwd+bwd {t_r:.2f}s" +
            (f", fp32 eager UNet batch 2 {t_u:.2f}s" if with_unet else "") +
            (f", fp32 eager VAE encode+input-grad 1 image {t_v:.2f}s" if t_v else "") + f" = 1/{a.views} of an iteration")
    return t_r + t_u + t_v, note


def run_reference(a, rank, world):
    """--impl reference: the reference path on the host cores (oracle port; the reference has no CPU
    code of its own and its CUDA core needs a GPU). Rank 0 only. A step is a bounded SAMPLE (1/B of an
    iteration, see cpu_sample): ms_per_step is the truly measured time of a sample, `value` extrapolates."""
    if rank != 0:
        return
    with_unet = unet_available() and a.phase in ("auto", "sds")
    for _ in range(min(a.warmup, 1)):
        cpu_sample(a, with_unet)
    ts, note = [], ""
    for _ in range(a.steps):
        s, note = cpu_sample(a, with_unet)
        ts.append(s)
    sample_s = float(np.mean(ts))
    val = 1.0 / (sample_s * a.views)
    line = {
        "impl": "reference", "metric": "SDS iterations/sec", "value": val, "unit": "it/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": sample_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "extrapolated": True, "sample_fraction_of_iteration": 1.0 / a.views,
        "config": workload_config(a, 1, with_unet),
        "cpu_baseline": {"value": val, "unit": "it/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": note + f"; value = 1 / ({a.views} x sample seconds); ms_per_step = one sample"},
        "e2e": {"value": val, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def gpu_reference_leg(a, dev, packed, cams, B, S):
    """BASELINE.md s.4 'GPU reference', measured in the same run: the UNMODIFIED reference CUDA rasteriser
    (oracle/_ref, built from /root/reference) called once per view like GaussianDreamer.forward (:189-191),
    and the PyTorch-eager fp16 restatements of the UNet (F.scaled_dot_product_attention, what the reference
    executes on torch >= 2) and of the VAE encoder with autograd. A baseline leg, never the product path."""
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    P = a.P
    try:
        from oracle.ref_cuda import RefRasterizer, available
        if available():
            xyz, shs, op, sc, rot = (packed[0:3 * P].view(P, 3), packed[3 * P:6 * P].view(P, 1, 3), packed[6 * P:7 * P].contiguous(),
                                     packed[7 * P:10 * P].view(P, 3), packed[10 * P:14 * P].view(P, 4))
            rr = [RefRasterizer() for _ in range(B)]
            bg = torch.ones(3, device=dev)
            gen = torch.Generator().manual_seed(7)
            dc, dd, da = (torch.randn(B, n, S, S, generator=gen).to(dev) for n in (3, 1, 1))
            kw = dict(shs=shs, scales=sc, rotations=rot)
            tf, tb = [], []
            for it in range(6):
                torch.cuda.synchronize()
                e0, e1, e2 = ev(), ev(), ev()
                e0.record()
                outs = [rr[b].forward(xyz, op, c.viewmatrix, c.projmatrix, c.campos, S, S, c.tanfovx, c.tanfovy, bg, **kw)
                        for b, c in enumerate(cams)]
                e1.record()
                for b, c in enumerate(cams):
                    rr[b].backward(xyz, outs[b]["radii"], outs[b]["alpha"], c.viewmatrix, c.projmatrix, c.campos, c.tanfovx,
                                   c.tanfovy, bg, dc[b], dd[b], da[b], **kw)
                e2.record()
                torch.cuda.synchronize()
                if it >= 2:
                    tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
            out["raster_fwd_ms"], out["raster_bwd_ms"] = float(np.median(tf)), float(np.median(tb))
            del rr
        if unet_available() and a.phase != "raster":
            from oracle import unet_ref, vae_ref
            from garmentdreamer_b200.unet_init import random_state_dict, random_vae_state_dict
            sd = random_state_dict(0, dev, torch.float16)
            g = torch.Generator().manual_seed(1)
            x = torch.randn(2 * B, 4, S // 8, S // 8, generator=g).to(dev).half()
            t = torch.randint(20, 981, (2 * B,), generator=g).to(dev).half()
            ctx = torch.randn(2 * B, 77, 1024, generator=g).to(dev).half()
            ts = []
            with torch.no_grad():
                for it in range(5):
                    e0, e1 = ev(), ev()
                    e0.record(); unet_ref.unet_forward(sd, x, t, ctx); e1.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ts.append(e0.elapsed_time(e1))
            out["unet_fp16_eager_ms"] = float(np.median(ts))
            del sd
            if not a.no_vae:
                sdv = {k: v.to(dev).half() for k, v in random_vae_state_dict(0, dev).items()}
                img = torch.rand(B, 3, S, S, generator=g).to(dev)
                n = torch.randn(B, 4, S // 8, S // 8, generator=g).to(dev)
                gl = torch.randn(B, 4, S // 8, S // 8, generator=g).to(dev)
                ts = []
                for it in range(4):
                    e0, e1 = ev(), ev()
                    e0.record(); vae_ref.encode_with_grad(sdv, img, n, gl); e1.record()
                    torch.cuda.synchronize()
                    if it >= 1:
                        ts.append(e0.elapsed_time(e1))
                out["vae_fp16_eager_autograd_ms"] = float(np.median(ts))
                del sdv
            torch.cuda.empty_cache()
        tot = sum(v for k, v in out.items() if k.endswith("_ms"))
        out["ms_per_step"] = tot
        out["value"] = 1e3 / tot if tot > 0 else None
        out["unit"] = "it/s"
        out["what"] = ("unmodified reference CUDA rasteriser (per-view calls, incl. its blocking read-backs) + PyTorch-eager fp16 "
                       "UNet/VAE restatements on this GPU; excludes cameras / activations / Adam (eager PyTorch in the reference)")
    except Exception as e:   # noqa: BLE001  -- a baseline leg never takes the bench line down
        out["error"] = str(e)[:300]
    return out


# ------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    import torch.distributed as dist
    from garmentdreamer_b200 import _lib, parallel, raster
    from garmentdreamer_b200.gaussians import GaussianParams
    from garmentdreamer_b200.synthetic import garment, raw_params, sample_batch, sample_cameras
    from garmentdreamer_b200.system import GaussianDreamerB200
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.raster_lib()
    with_unet = unet_available() and a.phase in ("auto", "sds")
    if a.strong and a.views % world:
        raise SystemExit("--strong needs views divisible by the number of ranks")
    B = a.views // world if a.strong else a.views          # views of this rank
    Btot = B * world
    guidance = None
    if with_unet:
        from garmentdreamer_b200 import sds_step
        guidance = sds_step.make_bench_guidance(dev, B, use_vae=not a.no_vae)
    P, S = a.P, a.res
    N, T = S * S, ((S + 15) // 16) ** 2
    raw = {k: v.to(dev) for k, v in raw_params(garment(P, 0)).items()}

    def new_system():
        gp = GaussianParams(raw["xyz"], raw["f_dc"], raw["opacity"], raw["scaling"], raw["rotation"], spatial_lr_scale=4.0)
        gp.training_setup()
        return GaussianDreamerB200(gp, guidance)

    system = new_system()
    lo, hi = parallel.shard_views(Btot, rank, world)
    batch_host = sample_batch(Btot, S, S, lo=lo, hi=hi)
    for k in ("c2w_3dgs", "fovy", "elevation", "azimuth", "camera_distances"):
        batch_host[k] = batch_host[k].contiguous().pin_memory()
    batch_dev = dict(batch_host)
    batch_dev["c2w_3dgs"] = batch_host["c2w_3dgs"].to(dev)
    if guidance is None:   # rasteriser-only runs: seeded upstream colour gradient
        gen = torch.Generator().manual_seed(7 + rank)
        dcol = torch.randn(B, 3, S, S, generator=gen).to(dev)
        batch_host["dL_dcolor"] = batch_dev["dL_dcolor"] = dcol
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(batch_host[k].numel() * 4 for k in ("c2w_3dgs", "fovy", "elevation", "azimuth", "camera_distances"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ----
    for _ in range(max(3, a.warmup)):
        out = system.training_step(batch_dev)
    torch.cuda.synchronize()
    R_total, overflow = raster.read_counters(out["state"])
    assert not overflow, "instance arena overflow in warm-up"
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = lib.gd_launch_count()
    if guidance is not None:
        guidance.reset_counters()
    system.timers = {}
    step_ms = []
    for _ in range(a.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the timed events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = system.training_step(batch_dev)
        e1.record()
        step_ms.append((e0, e1))
    barrier()
    launches = lib.gd_launch_count() - launches0
    if guidance is not None:
        launches += guidance.launch_count_delta()
    dev_ms = sum(x.elapsed_time(y) for x, y in step_ms)
    phases = system.phase_ms()
    system.timers = None
    bwd_ms = phases["raster_bwd"]
    sds_roofline = guidance.roofline(*measured_peaks()) if guidance is not None else None
    vae_ms = guidance.vae_ms() if guidance is not None else (0.0, 0.0)
    # ---- end to end through the public API with host buffers (`e2e`): the camera batch arrives in pinned
    # host memory every step (as from the reference's DataLoader) and the step's loss is read back ----
    barrier()
    t_e2e = []
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = system.training_step(batch_host)             # H2D of c2w / fovy / elevation / azimuth / distances inside
        loss_host.copy_(out["loss_sparsity"], non_blocking=True)
        e1.record()
        e1.synchronize()                                   # the host reads the loss every step (Lightning's self.log)
        t_e2e.append((e0, e1))
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_ms = sum(x.elapsed_time(y) for x, y in t_e2e)
    # ---- the collective alone (after a barrier: no rank skew in it) ----
    ar_us = px_us = None
    if world > 1:
        buf = torch.zeros_like(system.grad)
        for _ in range(3):
            dist.all_reduce(buf)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_reduce(buf)
        e1.record()
        torch.cuda.synchronize()
        ar_us = e0.elapsed_time(e1) / 10 * 1e3
        px_us = None
        if getattr(system, "peers", None) is not None:   # the exchange the step really uses: peer memory, fused with Adam
            px, g = system.peers, system.gaussian
            step0 = g.step_count
            for _ in range(3):
                px.allreduce(); g.adam_step_peers(px, densify=False)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                px.allreduce(); g.adam_step_peers(px, densify=False)
            e1.record()
            torch.cuda.synchronize()
            px_us = e0.elapsed_time(e1) / 10 * 1e3
    t = torch.tensor([dev_ms, e2e_ms, phases["allreduce"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, ar_phase_ms = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_kind = measured_peaks()
    units = 1 if a.strong else world      # weak: every rank does one iteration over its own B views per step
    value = a.steps * units / (dev_ms * 1e-3)
    e2e_value = a.steps * units / (e2e_ms * 1e-3)
    abytes = algorithmic_bytes_bwd(P * B, R_total, N * B, T * B)
    achieved = abytes / (bwd_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_render_bwd + k_bwd_epilogue (raster backward)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": 175.3e6 if (P, B, S) == (100000, 4, 512) else None,
                "traffic_source": "profiles/r02_final_raster_ncu_summary.txt (dram read+write of k_render_bwd 73.0+21.8 MB + k_bwd_epilogue 76.7+3.8 MB, c2)",
                "limiter": "instruction issue (ncu: 71 % of the issue slots, 353 M warp instructions; DRAM 2.4 % busy) -- the HBM roofline is the contract's yardstick",
                "peak_source": peak_kind, "algorithmic_bytes": abytes, "launch_ms": bwd_ms, "R": R_total}
    if sds_roofline is not None:
        sds_roofline["raster_bwd"] = roofline
        roofline = sds_roofline
        from garmentdreamer_b200 import sds_step as _probe
        try:   # an extra measurement after the timed region: never let it take the bench line down
            roofline["dominant_kernels"] = _probe.dominant_gemm_probe(dev, peaks)
        except Exception as e:   # noqa: BLE001
            roofline["dominant_kernels"] = []
            roofline["dominant_kernels_error"] = str(e)[:200]
        if guidance.vae is not None:
            from garmentdreamer_b200 import sds_step as _s
            roofline["vae"] = {"bound": "tensor", "kernel": "VAE encode + input-gradient backward (k_gemm_tcgen05 conv-GEMM + GroupNorm sweeps)",
                               "encode_ms": vae_ms[0], "backward_ms": vae_ms[1],
                               "achieved": (_s.VAE_FWD_FLOPS_PER_IMAGE + _s.VAE_BWD_FLOPS_PER_IMAGE) * B * (S / 512.0) ** 2
                               / ((vae_ms[0] + vae_ms[1]) * 1e-3) / 1e12,
                               "peak": peaks.get("bf16_tflops_sustained"), "unit": "TFLOP/s"}
            roofline["vae"]["frac"] = roofline["vae"]["achieved"] / roofline["vae"]["peak"]
    cfg = workload_config(a, world, with_unet)
    cfg["num_rendered_rank0"] = R_total
    line = {
        "metric": "SDS iterations/sec", "value": value, "unit": "it/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(3, a.warmup), "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
        "scaling": "strong" if a.strong else "weak", "vs_baseline": None,
        "dtype": "f32 raster / f16 UNet+VAE" if guidance else "f32", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "note": "parameters and optimiser state live on the GPU (as in the reference); per step the camera batch is copied "
                        "from pinned host memory and the loss is read back (one host sync per step)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "raster_bwd_gbs": achieved,
        "phase_ms": phases,
    }
    if world > 1:
        line["allreduce_us"] = {"nccl_allreduce_isolated_after_barrier": ar_us, "inside_step_incl_rank_skew_max_over_ranks": ar_phase_ms * 1e3,
                                "bytes": int(system.grad.numel() * 4 + system.radii_max.numel() * 4),
                                "path": "peer memory (gd_peer_allreduce + gd_params_adam_peers)" if px_us is not None else "nccl",
                                "peer_exchange_plus_adam_isolated_after_barrier": px_us,
                                "multicast": bool(system.peers.multicast) if px_us is not None else None}
    if not a.no_gpu_reference and world == 1:
        cams = sample_cameras(Btot, S, S)[lo:hi]
        for c in cams:
            c.viewmatrix, c.projmatrix, c.campos = c.viewmatrix.to(dev), c.projmatrix.to(dev), c.campos.to(dev)
        line["gpu_reference"] = gpu_reference_leg(a, dev, system.gaussian.activated(), cams, B, S)
        if line["gpu_reference"].get("ms_per_step"):
            line["gpu_reference"]["ours_ms_per_step"] = dev_ms / a.steps
    if not a.no_cpu_baseline and world == 1:
        s, note = cpu_sample(a, guidance is not None)
        s, note = cpu_sample(a, guidance is not None)   # second call: warm (allocator, OpenMP threads)
        line["cpu_baseline"] = {"value": 1.0 / (s * a.views), "unit": "it/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": note + "; extrapolated x" + str(a.views)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
