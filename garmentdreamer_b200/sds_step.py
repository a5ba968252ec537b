"""One SDS step on rendered views, as bench.py times it (BASELINE config 2):

    colour [B,3,S,S] --VAE encoder (VAEEncoderB200.encode: 2c-1, encoder, sample, x0.18215)-->
      latents [B,4,S/8,S/8]
      --> StableDiffusionGuidance.compute_grad_sds (B200 UNet, batch 2B) --> grad
      --> nan_to_num / clamp / 1/B (guidance __call__, :418-427)
      --> VAE encoder input-gradient backward --> dL/dcolour [B,3,S,S]

With ``use_vae=False`` the encoder is replaced by a linear stand-in (8x8 mean of 2*rgb-1 and a
fixed 3->4 channel mix); numbers produced that way say "VAE excluded" in their config.
"""
import torch

from . import unet_ops as ops
from .guidance import PromptProcessorOutput, StableDiffusionGuidance
from .unet import UNetB200

UNET_FLOPS_PER_SAMPLE = 804.3e9  # SURVEY.md Appendix B (conv 418.4 + linear 259.8 + attention 126.1 GFLOP)
# VAE encoder at 512^2, per image, forward (conv 1057.6 + linear/1x1 17.2 + attention matmuls 34.4 GFLOP);
# the input-gradient backward repeats every contraction once (dgrad) and the attention matmuls twice
VAE_FWD_FLOPS_PER_IMAGE = 1109.2e9
VAE_BWD_FLOPS_PER_IMAGE = 1109.2e9 + 34.4e9


def _random_state_dict(seed, device):
    from .unet_init import random_state_dict
    return random_state_dict(seed, device, torch.float16)


class SdsBenchStep:
    def __init__(self, dev, views, state_dict=None, seed=0, use_cuda_graph=True, use_vae=True, guidance_res=None):
        self.dev = torch.device(dev)
        self.B = views
        self.guidance_res = guidance_res   # e.g. 512: renders of another size are resized (bilinear) before the VAE
        self.vae = None
        if use_vae:
            from .unet_init import random_vae_state_dict
            from .vae import VAEEncoderB200
            self.vae_state_dict = random_vae_state_dict(seed, self.dev)
            self.vae = VAEEncoderB200(self.vae_state_dict, self.dev)
        self._vae_ms = []
        sd = state_dict if state_dict is not None else _random_state_dict(seed, self.dev)
        self.unet_state_dict = sd
        self.unet = UNetB200(sd, self.dev, use_cuda_graph=use_cuda_graph)
        g = torch.Generator().manual_seed(1234 + seed)
        bank = lambda n: torch.randn(n, 77, 1024, generator=g).to(self.dev)
        self.prompt = PromptProcessorOutput(bank(1), bank(1), bank(4), bank(4))
        self.gen = torch.Generator(device=self.dev).manual_seed(99 + seed)
        self.guidance = StableDiffusionGuidance(self.unet, self.dev, generator=self.gen)
        self.guidance.grad_clip_val = 1.5
        self.mix = (torch.tensor([[0.6, 0.3, 0.1], [-0.3, 0.5, -0.2], [0.2, -0.4, 0.6], [0.3, 0.3, -0.6]]) * 2.0).to(self.dev).contiguous()
        self._launch0 = ops.lib().gd_unet_launch_count()
        self._unet_ms = []
        self.last_grad = None

    def set_min_max_steps(self, min_step_percent=0.02, max_step_percent=0.98):
        self.guidance.set_min_max_steps(min_step_percent, max_step_percent)

    def image_grad(self, color, elevation, azimuth, distances, scale=None):
        """color [B,3,S,S] fp32 (rasteriser output) -> dL_sds/dcolor [B,3,S,S] fp32 of
        loss_sds = 0.5 * mse(latents, (latents - grad).detach(), 'sum') * scale (scale = 1/B by default:
        stable_diffusion_guidance.py:427; sharded views pass 1/(B*world))."""
        L = ops.lib()
        stream = torch.cuda.current_stream().cuda_stream
        full = None
        if self.guidance_res and tuple(color.shape[-2:]) != (self.guidance_res, self.guidance_res):
            # the reference renders at data.height x data.width (1024^2 shipped) and resizes to 512^2 for the guidance
            # (stable_diffusion_guidance.py:387-396): bilinear forward here, its transpose on the way back
            full = color
            Bc, Cc, Hf, Wf = color.shape
            color = torch.empty((Bc, Cc, self.guidance_res, self.guidance_res), dtype=torch.float32, device=full.device)
            ops._chk(L.gd_resize_bilinear(full.data_ptr(), color.data_ptr(), Bc * Cc, Hf, Wf, self.guidance_res, self.guidance_res, stream),
                     "resize_bilinear")
        B, _, H, W = color.shape
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        if self.vae is not None:
            noise = torch.randn((B, 4, H // 8, W // 8), device=color.device, dtype=torch.float32, generator=self.gen)
            lat = self.vae.encode(color, noise)
        else:
            lat = torch.empty((B, 4, H // 8, W // 8), dtype=torch.float32, device=color.device)
            ops._chk(L.gd_unet_pool_latents(color.data_ptr(), self.mix.data_ptr(), lat.data_ptr(), B, H, W, stream), "pool_latents")
        t = torch.randint(self.guidance.min_step, self.guidance.max_step + 1, [B], dtype=torch.long, device=color.device,
                          generator=self.gen)
        to_dev = lambda x: torch.as_tensor(x, dtype=torch.float32).to(color.device, non_blocking=True)
        elev, azim, dist = to_dev(elevation), to_dev(azimuth), to_dev(distances)
        scale = 1.0 / B if scale is None else float(scale)
        ev[1].record()
        grad, _ = self.guidance.compute_grad_sds(lat, t, self.prompt, elev, azim, dist)
        ev[2].record()
        self._unet_ms.append((ev[1], ev[2]))
        self.last_grad = grad
        clip = float(self.guidance.grad_clip_val or 0.0)
        if self.vae is not None:
            dcol = self.vae.backward(grad, clip=clip, scale=scale)
        else:
            dcol = torch.empty_like(color)
            ops._chk(L.gd_unet_pool_latents_bwd(grad.data_ptr(), self.mix.data_ptr(), dcol.data_ptr(), B, H, W, clip, scale,
                                                stream), "pool_latents_bwd")
        if full is not None:
            dfull = torch.empty_like(full)
            ops._chk(L.gd_resize_bilinear_bwd(dcol.data_ptr(), dfull.data_ptr(), B * 3, full.shape[2], full.shape[3], H, W, stream),
                     "resize_bilinear_bwd")
            dcol = dfull
        ev[3].record()
        self._vae_ms.append((ev[0], ev[1], ev[2], ev[3]))
        return dcol

    def reset_counters(self):
        self._launch0 = ops.lib().gd_unet_launch_count()
        self._unet_ms = []
        self._vae_ms = []

    def vae_ms(self):
        """(encode ms, backward ms) averaged over the steps since reset (0, 0 with the stand-in)."""
        n = max(1, len(self._vae_ms))
        return (float(sum(e[0].elapsed_time(e[1]) for e in self._vae_ms) / n),
                float(sum(e[2].elapsed_time(e[3]) for e in self._vae_ms) / n))

    def launch_count_delta(self):
        """Kernels of libgd_unet.so launched since reset; a CUDA-graph replay re-launches the
        captured kernels without passing through the C ABI, so those are counted from the graph."""
        direct = ops.lib().gd_unet_launch_count() - self._launch0
        return int(direct + self.unet.graph_kernel_launches_since_reset())

    def compute_grad_ms(self):
        return float(sum(a.elapsed_time(b) for a, b in self._unet_ms) / max(1, len(self._unet_ms)))

    def roofline(self, peaks, peak_kind, raster_bwd=None):
        ms = self.compute_grad_ms()
        flops = UNET_FLOPS_PER_SAMPLE * 2 * self.B
        achieved = flops / (ms * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        r = {"bound": "tensor", "kernel": "compute_grad_sds = UNet forward batch 2B (k_gemm_tcgen05 + k_flash_attn "
                                          "carry ~85% of its time, profiles/) + noise/SDS epilogue kernels",
             "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
             "peak_source": peak_kind + " (sustained cuBLAS bf16; fp16 tcgen05 has the same peak)",
             "algorithmic_flops": flops, "launch_ms": ms}
        if raster_bwd is not None:
            r["raster_bwd"] = raster_bwd
        return r


def dominant_gemm_probe(dev, peaks):
    """Live CUDA-event timing of the two GEMM shapes that carry the largest share of the step
    (ncu launch list in profiles/): the 128->128 3x3 convolution of the VAE at 4x512^2 (8 launches
    per step) and the 1280->1280 3x3 convolution of the UNet at 8x16^2 (7 launches). 20 launches of
    each are captured in a CUDA graph; burst peak = a kernel timed alone."""
    out = []
    peak = peaks.get("bf16_tflops", peaks.get("bf16_tflops_sustained"))
    for name, (N_, H, W, Ci, Co), traffic in (("k_gemm_tcgen05<0,1> conv 128->128 @4x512^2 (VAE)", (4, 512, 512, 128, 128), 502.0e6),
                                              ("k_gemm_tcgen05<0,1> conv 1280->1280 @8x16^2 (UNet)", (8, 16, 16, 1280, 1280), None)):
        x = torch.randn(N_, H, W, Ci, device=dev).half()
        w = (torch.randn(Co, 9 * Ci, device=dev) * (9 * Ci) ** -0.5).half()
        b = torch.randn(Co, device=dev).half()
        y = torch.empty(N_, H, W, Co, device=dev, dtype=torch.float16)
        ops.conv3x3(x, w, b, out=y)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                ops.conv3x3(x, w, b, out=y)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 60 * 1e3
        flops = 2.0 * N_ * H * W * 9 * Ci * Co
        ach = flops / us / 1e6
        out.append({"kernel": name, "bound": "tensor", "launch_us": us, "algorithmic_flops": flops, "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic,   # dram read+write bytes per launch from the committed ncu --set full capture, or None
                    "traffic_source": "profiles/r02_final_gemm_halo128_summary.txt (dram 268.8 MB read + 233.2 MB written; algorithmic 537 MB)" if traffic else None})
        del g
    return out


def make_bench_guidance(dev, views, use_vae=True, guidance_res=None):
    return SdsBenchStep(dev, views, use_vae=use_vae, guidance_res=guidance_res)
