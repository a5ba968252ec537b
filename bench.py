#!/usr/bin/env python
"""Contract benchmark: SDS iterations/s on the synthetic garment (BASELINE.json configs[1]:
100k Gaussians, 4 camera views 512^2 per GPU), one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference path on the host cores (oracle port)

One "step" = one pass of the hot path over one batch of B views:
  rasterise forward (B views) -> SDS gradient of the UNet step (when the UNet library is built;
  see config.workload) -> rasterise backward -> (N > 1) NCCL all-reduce of the packed per-Gaussian
  gradient [P,14]. Views are sharded over ranks (weak scaling: B views per GPU); the Gaussians are
  replicated. Prints ONE JSON line on rank 0.
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
    ap.add_argument("--views", type=int, default=4, help="views per GPU")
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--phase", default="auto", choices=["auto", "raster", "sds"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


# ------------------------------------------------------------------------------------------
def cpu_reference_sample(a, n_views=1):
    """The reference path on the host cores: C oracle (OpenMP) rasterise fwd+bwd of `n_views`
    views of the same workload. Returns seconds per view and the thread count used."""
    from garmentdreamer_b200.synthetic import garment, sample_cameras
    from oracle import raster_oracle as ro
    g = garment(a.P, 0)
    cams = sample_cameras(a.views, a.res, a.res)
    gen = torch.Generator().manual_seed(7)
    dc = torch.randn(3, a.res, a.res, generator=gen).numpy()
    dd = torch.randn(1, a.res, a.res, generator=gen).numpy()
    da = torch.randn(1, a.res, a.res, generator=gen).numpy()
    ts = []
    for v in range(n_views):
        c = cams[v % len(cams)]
        t0 = time.perf_counter()
        st = ro.forward(g["xyz"].numpy(), g["opacity"].numpy(), c.viewmatrix.numpy(), c.projmatrix.numpy(),
                        c.campos.numpy(), a.res, a.res, c.tanfovx, c.tanfovy, np.ones(3, np.float32),
                        shs=g["shs"].numpy(), scales=g["scales"].numpy(), rotations=g["rotations"].numpy())
        ro.backward(st, dc, dd, da)
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)), os.cpu_count()


def run_reference(a, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the
    reference has no CPU code of its own and its CUDA core needs a GPU). Rank 0 only."""
    if rank != 0:
        return
    per_view, cores = None, os.cpu_count()
    for _ in range(max(1, a.warmup // 3)):
        cpu_reference_sample(a, 1)
    t = []
    for _ in range(a.steps):
        s, cores = cpu_reference_sample(a, 1)
        t.append(s)
    per_view = float(np.mean(t))
    unet_s, unet_note = cpu_unet_sample(a)
    step_s = per_view * a.views + unet_s
    val = 1.0 / step_s
    workload, _ = workload_name(a, unet_s > 0)
    line = {
        "impl": "reference", "metric": "SDS iterations/sec", "value": val, "unit": "it/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload, "P": a.P, "views_per_gpu": a.views, "res": a.res},
        "cpu_baseline": {"value": val, "unit": "it/s", "cores": cores, "kind": "port",
                         "sample": f"1 of {a.views} views rasterised fwd+bwd by the C oracle per step "
                                   f"(x{a.views} extrapolated){unet_note}"},
        "e2e": {"value": val, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_unet_sample(a):
    """Seconds the PyTorch-eager fp32 UNet restatement needs on the host cores for the step's
    2B samples (bounded: one cond+uncond pair is timed and scaled). 0 if the UNet is not built."""
    try:
        from oracle import unet_ref
    except Exception:
        return 0.0, ""
    if not unet_available():
        return 0.0, ""
    torch.set_num_threads(os.cpu_count())
    s = unet_ref.time_cpu_forward(batch=2)
    total, note = s * a.views, f"; UNet fp32 eager on CPU: batch 2 timed ({s:.1f}s), x{a.views} extrapolated"
    if not a.no_vae:
        from oracle import vae_ref
        v = vae_ref.time_cpu_encode(batch=1, res=a.res)
        total += v * a.views
        note += f"; VAE encode + input-gradient fp32 eager on CPU: 1 image timed ({v:.1f}s), x{a.views} extrapolated"
    return total, note


def unet_available():
    return os.path.exists(os.path.join(ROOT, "garmentdreamer_b200", "lib", "libgd_unet.so"))


def workload_name(a, with_unet):
    if with_unet and not a.no_vae:
        return (f"c2: {a.P} Gaussians, {a.views}x{a.res}^2 views per GPU, full SDS loop: raster fwd + VAE encode + SD-2.1 UNet "
                f"SDS grad (batch {2 * a.views}) + VAE input-gradient bwd + raster bwd (random-init fp16 networks)"), True
    if with_unet:
        return (f"c2: {a.P} Gaussians, {a.views}x{a.res}^2 views per GPU, raster fwd + SD-2.1 UNet SDS grad "
                f"(batch {2 * a.views}, random-init fp16) + raster bwd; VAE excluded (latents = 8x8-pooled render)"), True
    return (f"c2-raster: {a.P} Gaussians, {a.views}x{a.res}^2 views per GPU, rasterise fwd+bwd only "
            f"(UNet step not in this build)"), False


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
    from garmentdreamer_b200.synthetic import garment, sample_cameras
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.raster_lib()
    with_unet = unet_available() and a.phase in ("auto", "sds")
    guidance = None
    if with_unet:
        from garmentdreamer_b200 import sds_step
        guidance = sds_step.make_bench_guidance(dev, a.views, use_vae=not a.no_vae)
    workload, _ = workload_name(a, with_unet)

    P, B, S = a.P, a.views, a.res
    N, T = S * S, ((S + 15) // 16) ** 2
    host = garment(P, 0)
    # packed activated parameters, 14 floats per Gaussian, struct-of-arrays in one flat buffer:
    # xyz 3P | f_dc 3P | opacity P | scales 3P | rotation 4P  (contiguous slices, one H2D copy)
    packed_host = torch.cat([host["xyz"].reshape(-1), host["shs"].reshape(-1), host["opacity"].reshape(-1),
                             host["scales"].reshape(-1), host["rotations"].reshape(-1)]).contiguous().pin_memory()
    lo, hi = parallel.shard_views(B * world, rank, world)
    cams = sample_cameras(B * world, S, S)[lo:hi]
    cam_host = torch.stack([torch.cat([c.viewmatrix.reshape(-1), c.projmatrix.reshape(-1), c.campos])
                            for c in cams]).contiguous().pin_memory()  # [B,35]
    bg = torch.ones(3, device=dev)
    gen = torch.Generator().manual_seed(7 + rank)
    dL_dcolor = torch.randn(B, 3, S, S, generator=gen).to(dev)
    zeros1 = torch.zeros(B, 1, S, S, device=dev)
    grad_host = torch.empty(P * 14, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def unpack(p):
        return (p[0:3 * P].view(P, 3), p[3 * P:6 * P].view(P, 1, 3), p[6 * P:7 * P].view(P, 1),
                p[7 * P:10 * P].view(P, 3), p[10 * P:14 * P].view(P, 4))

    def make_views(cm):
        return [raster.View(cm[b, 0:16], cm[b, 16:32], cm[b, 32:35], cams[b].tanfovx, cams[b].tanfovy)
                for b in range(B)]

    timers = {"bwd_ms": [], "R": 0}

    def step(packed_dev, cam_dev, time_bwd=False):
        xyz, shs, op, sc, rot = unpack(packed_dev)
        views = make_views(cam_dev)
        color, depth, alpha, radii, st = raster.forward_views(xyz, op, views, S, S, bg, shs=shs, scales=sc,
                                                              rotations=rot, sync=False)
        if guidance is not None:
            dcol = guidance.image_grad(color, cams)
        else:
            dcol = dL_dcolor
        if time_bwd:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        out = torch.empty(P * 14, dtype=torch.float32, device=dev)  # same SoA packing as the parameters
        o3, osh, oop, osc, orot = unpack(out)
        raster.backward_views(st, xyz, radii, alpha, bg, dcol, zeros1, zeros1, shs=shs, scales=sc,
                              rotations=rot, sum_views=True,
                              out={"means3D": o3, "sh": osh, "opacity": oop, "scales": osc, "rotations": orot})
        if time_bwd:
            e1.record()
            timers["ev"].append((e0, e1))
        parallel.allreduce_gradients(out)  # the only data-path collective (NCCL over NVLink)
        return out, st

    packed_dev = packed_host.to(dev, non_blocking=True)
    cam_dev = cam_host.to(dev, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ----
    for _ in range(max(3, a.warmup)):
        out, st = step(packed_dev, cam_dev)
    torch.cuda.synchronize()
    R_total, overflow = raster.read_counters(st)
    assert not overflow, "instance arena overflow in warm-up"
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = lib.gd_launch_count()
    if guidance is not None:
        guidance.reset_counters()
    timers["ev"] = []
    step_ms = []
    for _ in range(a.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the timed events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, st = step(packed_dev, cam_dev, time_bwd=True)
        e1.record()
        step_ms.append((e0, e1))
    barrier()
    launches = lib.gd_launch_count() - launches0
    if guidance is not None:
        launches += guidance.launch_count_delta()
    dev_ms = sum(x.elapsed_time(y) for x, y in step_ms)
    bwd_ms = float(np.mean([x.elapsed_time(y) for x, y in timers["ev"]]))
    sds_roofline = guidance.roofline(*measured_peaks()) if guidance is not None else None
    vae_ms = guidance.vae_ms() if guidance is not None else (0.0, 0.0)
    # ---- end to end through the public API with host buffers (`e2e`) ----
    barrier()
    t_e2e = []
    for _ in range(a.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pd = packed_host.to(dev, non_blocking=True)
        cd = cam_host.to(dev, non_blocking=True)
        out, st = step(pd, cd)
        grad_host.copy_(out, non_blocking=True)
        e1.record()
        t_e2e.append((e0, e1))
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_ms = sum(x.elapsed_time(y) for x, y in t_e2e)
    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_kind = measured_peaks()
    # one unit = one SDS iteration over B views (BASELINE metric: "4 x 512^2 views"); every rank
    # processes one unit per step on its own shard of the view batch, so the job does `world` units
    value = a.steps * world / (dev_ms * 1e-3)
    e2e_value = a.steps * world / (e2e_ms * 1e-3)
    abytes = algorithmic_bytes_bwd(P * B, R_total, N * B, T * B)
    achieved = abytes / (bwd_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_render_bwd + k_bwd_epilogue (raster backward)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": None, "peak_source": peak_kind, "algorithmic_bytes": abytes,
                "launch_ms": bwd_ms, "R": R_total}
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
    line = {
        "metric": "SDS iterations/sec", "value": value, "unit": "it/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(3, a.warmup), "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 raster / f16 UNet" if guidance else "f32",
        "data": "synthetic",
        "config": {"workload": workload, "P": P, "views_per_gpu": B, "views_total": B * world, "res": S,
                   "num_rendered_rank0": R_total, "l2": "flushed between timed iterations (256 MiB write)",
                   "unit_of_work": f"one SDS iteration = {B} views; value = ranks x steps / max-over-ranks time",
                   "parallelism": f"views sharded x{world}, NCCL all-reduce of [P,14] grads" if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": packed_host.numel() * 4 + cam_host.numel() * 4,
                "d2h_bytes_per_step": grad_host.numel() * 4},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "raster_bwd_gbs": achieved,
    }
    if not a.no_cpu_baseline and world == 1:
        per_view, cores = cpu_reference_sample(a, 1)
        unet_s, unet_note = cpu_unet_sample(a) if guidance is not None else (0.0, "")
        line["cpu_baseline"] = {"value": 1.0 / (per_view * B + unet_s), "unit": "it/s", "cores": cores, "kind": "port",
                                "sample": f"1 of {B} views rasterised fwd+bwd by the C oracle ({per_view:.2f}s), "
                                          f"x{B} extrapolated{unet_note}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
