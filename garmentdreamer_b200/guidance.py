"""Host-side mirror of the reference SDS guidance for the hot path, on the B200 UNet.

Mirrors (names, argument meaning, return values, error behaviour):
  * PromptProcessorOutput / DirectionConfig / shift_azimuth_deg
        Garment_3DGS/threestudio/models/prompt_processors/base.py:27-170, 255-294
  * StableDiffusionGuidance.{set_min_max_steps, forward_unet, compute_grad_sds, __call__,
    update_step}
        Garment_3DGS/threestudio/models/guidance/stable_diffusion_guidance.py:141-157,185-276,
        374-448,581-591
The UNet behind ``self.unet`` is garmentdreamer_b200.unet.UNetB200 (tcgen05 kernels). The noise
add and the CFG/SDS epilogue run as fused CUDA kernels of libgd_unet.so. encode_images (SURVEY.md
s.8 row f1) runs on garmentdreamer_b200.vae.VAEEncoderB200 (forward + input-gradient backward
behind a torch.autograd.Function) when one is passed through ``vae=``; ``rgb_as_latents=True``
skips it as in the reference.
"""
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import unet_ops as ops


def shift_azimuth_deg(azimuth):
    # shift azimuth angle (in degrees), to [-180, 180]
    return (azimuth + 180) % 360 - 180


def shifted_expotional_decay(a, b, c, r):
    return a * torch.exp(-b * r) + c


def perpendicular_component(x, y):
    # component of x that is perpendicular to y (threestudio/utils/ops.py:431-441)
    eps = torch.ones_like(x[:, 0, 0, 0]) * 1e-6
    return x - (torch.mul(x, y).sum(dim=[1, 2, 3]) / torch.maximum(torch.mul(y, y).sum(dim=[1, 2, 3]), eps)).view(-1, 1, 1, 1) * y


@dataclass
class DirectionConfig:
    name: str
    prompt: Callable[[str], str]
    negative_prompt: Callable[[str], str]
    condition: Callable[[Any, Any, Any], Any]


def default_directions(overhead_threshold=60.0, front_threshold=45.0, back_threshold=45.0):
    """side / front / back / overhead, later entries overwrite earlier ones (base.py:262-294)."""
    return [
        DirectionConfig("side", lambda s: f"{s}, side view", lambda s: s,
                        lambda ele, azi, dis: torch.ones_like(ele, dtype=torch.bool)),
        DirectionConfig("front", lambda s: f"{s}, front view", lambda s: s,
                        lambda ele, azi, dis: (shift_azimuth_deg(azi) > -front_threshold)
                        & (shift_azimuth_deg(azi) < front_threshold)),
        DirectionConfig("back", lambda s: f"{s}, back view", lambda s: s,
                        lambda ele, azi, dis: (shift_azimuth_deg(azi) > 180 - back_threshold)
                        | (shift_azimuth_deg(azi) < -180 + back_threshold)),
        DirectionConfig("overhead", lambda s: f"{s}, overhead view", lambda s: s,
                        lambda ele, azi, dis: ele > overhead_threshold),
    ]


@dataclass
class PromptProcessorOutput:
    text_embeddings: torch.Tensor             # [1,77,D]
    uncond_text_embeddings: torch.Tensor      # [1,77,D]
    text_embeddings_vd: torch.Tensor          # [4,77,D]
    uncond_text_embeddings_vd: torch.Tensor   # [4,77,D]
    directions: List[DirectionConfig] = field(default_factory=default_directions)
    direction2idx: Dict[str, int] = field(default_factory=lambda: {"side": 0, "front": 1, "back": 2, "overhead": 3})
    use_perp_neg: bool = False
    perp_neg_f_sb: Tuple[float, float, float] = (1, 0.5, -0.606)
    perp_neg_f_fsb: Tuple[float, float, float] = (1, 0.5, +0.967)
    perp_neg_f_fs: Tuple[float, float, float] = (4, 0.5, -2.426)
    perp_neg_f_sf: Tuple[float, float, float] = (4, 0.5, -2.426)

    def _direction_idx(self, elevation, azimuth, camera_distances):
        direction_idx = torch.zeros_like(elevation, dtype=torch.long)
        for d in self.directions:
            direction_idx[d.condition(elevation, azimuth, camera_distances)] = self.direction2idx[d.name]
        return direction_idx

    def get_text_embeddings(self, elevation, azimuth, camera_distances, view_dependent_prompting=True):
        batch_size = elevation.shape[0]
        if view_dependent_prompting:
            idx = self._direction_idx(elevation, azimuth, camera_distances)
            text_embeddings = self.text_embeddings_vd[idx]
            uncond_text_embeddings = self.uncond_text_embeddings_vd[idx]
        else:
            text_embeddings = self.text_embeddings.expand(batch_size, -1, -1)
            uncond_text_embeddings = self.uncond_text_embeddings.expand(batch_size, -1, -1)
        # IMPORTANT: (cond, uncond) order, as in the reference
        return torch.cat([text_embeddings, uncond_text_embeddings], dim=0)

    def get_text_embeddings_perp_neg(self, elevation, azimuth, camera_distances, view_dependent_prompting=True):
        assert view_dependent_prompting, "Perp-Neg only works with view-dependent prompting"
        batch_size = elevation.shape[0]
        direction_idx = self._direction_idx(elevation, azimuth, camera_distances)
        pos, neg, wts, unc = [], [], [], []
        side_emb, front_emb, back_emb, overhead_emb = (self.text_embeddings_vd[i] for i in range(4))
        for idx, ele, azi, dis in zip(direction_idx, elevation, azimuth, camera_distances):
            azi = shift_azimuth_deg(azi)
            unc.append(self.uncond_text_embeddings_vd[idx])
            if idx.item() == 3:  # overhead view
                pos.append(overhead_emb)
                neg += [self.uncond_text_embeddings_vd[idx], self.uncond_text_embeddings_vd[idx]]
                wts += [0.0, 0.0]
            elif torch.abs(azi) < 90:  # front-side interpolation
                r = 1 - torch.abs(azi) / 90
                pos.append(r * front_emb + (1 - r) * side_emb)
                neg += [front_emb, side_emb]
                wts += [-shifted_expotional_decay(*self.perp_neg_f_fs, r),
                        -shifted_expotional_decay(*self.perp_neg_f_sf, 1 - r)]
            else:  # side-back interpolation
                r = 2.0 - torch.abs(azi) / 90
                pos.append(r * side_emb + (1 - r) * back_emb)
                neg += [side_emb, front_emb]
                wts += [-shifted_expotional_decay(*self.perp_neg_f_sb, r),
                        -shifted_expotional_decay(*self.perp_neg_f_fsb, r)]
        text_embeddings = torch.cat([torch.stack(pos, 0), torch.stack(unc, 0), torch.stack(neg, 0)], dim=0)
        return text_embeddings, torch.as_tensor(wts, device=elevation.device).reshape(batch_size, 2)


def C(value, epoch, global_step):
    """Scheduled scalar [start_step, start_value, end_value, end_step] (utils/misc.py:65-86)."""
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) == 3:
        value = [0] + value
    assert len(value) == 4
    start_step, start_value, end_value, end_step = value
    cur = global_step if isinstance(end_step, int) else epoch
    return start_value + (end_value - start_value) * max(min(1.0, (cur - start_step) / (end_step - start_step)), 0.0)


def ddim_alphas_cumprod(n=1000, beta_start=0.00085, beta_end=0.012, device="cpu"):
    """DDIMScheduler(beta_schedule='scaled_linear').alphas_cumprod of the SD-2.1-base config."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, 0).to(device)


class StableDiffusionGuidance:
    @dataclass
    class Config:
        guidance_scale: float = 100.0
        grad_clip: Optional[Any] = None
        half_precision_weights: bool = True
        min_step_percent: float = 0.02
        max_step_percent: float = 0.98
        weighting_strategy: str = "sds"
        view_dependent_prompting: bool = True

    def __init__(self, unet, device="cuda", cfg=None, vae=None, generator=None):
        self.cfg = cfg or self.Config()
        self.device = torch.device(device)
        self.unet = unet  # callable: unet(x, t, encoder_hidden_states=...).sample
        self.vae = vae
        self.weights_dtype = torch.float16 if self.cfg.half_precision_weights else torch.float32
        self.num_train_timesteps = 1000
        self.set_min_max_steps()
        self.alphas = ddim_alphas_cumprod(device=self.device)
        self.grad_clip_val = None
        self.generator = generator

    def set_min_max_steps(self, min_step_percent=0.02, max_step_percent=0.98):
        self.min_step = int(self.num_train_timesteps * min_step_percent)
        self.max_step = int(self.num_train_timesteps * max_step_percent)

    def forward_unet(self, latents, t, encoder_hidden_states):
        input_dtype = latents.dtype
        return self.unet(latents.to(self.weights_dtype), t.to(self.weights_dtype),
                         encoder_hidden_states=encoder_hidden_states.to(self.weights_dtype)).sample.to(input_dtype)

    def _randn_like(self, x):
        if self.generator is None:
            return torch.randn_like(x)
        return torch.randn(x.shape, generator=self.generator, device=x.device, dtype=x.dtype)

    def compute_grad_sds(self, latents, t, prompt_utils, elevation, azimuth, camera_distances):
        batch_size = elevation.shape[0]
        lat = latents.detach().float().contiguous()
        chw = lat[0].numel()
        a_t = self.alphas[t].float()
        sqrt_ab, sqrt_1mab = a_t.sqrt().contiguous(), (1 - a_t).sqrt().contiguous()
        if self.cfg.weighting_strategy == "sds":
            w = (1 - a_t).contiguous()
        elif self.cfg.weighting_strategy == "uniform":
            w = torch.ones_like(a_t)
        elif self.cfg.weighting_strategy == "fantasia3d":
            w = (a_t ** 0.5 * (1 - a_t)).contiguous()
        else:
            raise ValueError(f"Unknown weighting strategy: {self.cfg.weighting_strategy}")
        stream = torch.cuda.current_stream().cuda_stream
        L = ops.lib()
        reps = 4 if prompt_utils.use_perp_neg else 2
        noise = self._randn_like(lat)
        latents_noisy = torch.empty_like(lat)
        unet_in = torch.empty((reps * batch_size,) + tuple(lat.shape[1:]), dtype=torch.float16, device=lat.device)
        # noise add (scheduler.add_noise) + duplication of the batch, one fused kernel
        ops._chk(L.gd_unet_add_noise(lat.data_ptr(), noise.data_ptr(), sqrt_ab.data_ptr(), sqrt_1mab.data_ptr(),
                                     latents_noisy.data_ptr(), unet_in.data_ptr(), batch_size, reps, chw, stream),
                 "add_noise")
        if prompt_utils.use_perp_neg:
            text_embeddings, neg_guidance_weights = prompt_utils.get_text_embeddings_perp_neg(
                elevation, azimuth, camera_distances, self.cfg.view_dependent_prompting)
            with torch.no_grad():
                noise_pred = self.forward_unet(unet_in, torch.cat([t] * 4), encoder_hidden_states=text_embeddings).float()
            noise_pred_text = noise_pred[:batch_size]
            noise_pred_uncond = noise_pred[batch_size:batch_size * 2]
            noise_pred_neg = noise_pred[batch_size * 2:]
            e_pos = noise_pred_text - noise_pred_uncond
            accum_grad = 0
            n_negative_prompts = neg_guidance_weights.shape[-1]
            for i in range(n_negative_prompts):
                e_i_neg = noise_pred_neg[i::n_negative_prompts] - noise_pred_uncond
                accum_grad += neg_guidance_weights[:, i].view(-1, 1, 1, 1) * perpendicular_component(e_i_neg, e_pos)
            noise_pred = noise_pred_uncond + self.cfg.guidance_scale * (e_pos + accum_grad)
            grad = w.view(-1, 1, 1, 1) * (noise_pred - noise)
        else:
            neg_guidance_weights = None
            text_embeddings = prompt_utils.get_text_embeddings(elevation, azimuth, camera_distances,
                                                               self.cfg.view_dependent_prompting)
            with torch.no_grad():  # predict the noise residual with unet, NO grad!
                eps = self.forward_unet(unet_in, torch.cat([t] * 2), encoder_hidden_states=text_embeddings)
            eps = eps.float().contiguous()
            noise_pred = torch.empty_like(lat)
            grad = torch.empty_like(lat)
            # noise_pred = e_text + s * (e_text - e_uncond); grad = w * (noise_pred - noise)
            ops._chk(L.gd_unet_sds_grad(eps.data_ptr(), noise.data_ptr(), w.data_ptr(), float(self.cfg.guidance_scale),
                                        noise_pred.data_ptr(), grad.data_ptr(), batch_size, chw, stream), "sds_grad")
        guidance_eval_utils = {
            "use_perp_neg": prompt_utils.use_perp_neg,
            "neg_guidance_weights": neg_guidance_weights,
            "text_embeddings": text_embeddings,
            "t_orig": t,
            "latents_noisy": latents_noisy,
            "noise_pred": noise_pred,
        }
        return grad, guidance_eval_utils

    def encode_images(self, imgs):
        """stable_diffusion_guidance.py:160-167. With a VAEEncoderB200 attached the whole function
        (imgs*2-1, encoder, sample, scaling) and its backward are CUDA kernels of libgd_unet.so;
        any other object is driven through the diffusers surface exactly like the reference."""
        if self.vae is None:
            raise RuntimeError("no VAE attached: pass vae=VAEEncoderB200(...) (or an object with the diffusers "
                               "encode(...).latent_dist.sample() surface), or use rgb_as_latents=True")
        from .vae import VAEEncoderB200
        if isinstance(self.vae, VAEEncoderB200):
            return self.vae.encode_images(imgs, generator=self.generator)
        input_dtype = imgs.dtype
        imgs = imgs * 2.0 - 1.0
        posterior = self.vae.encode(imgs.to(self.weights_dtype)).latent_dist
        latents = posterior.sample() * self.vae.config.scaling_factor
        return latents.to(input_dtype)

    def __call__(self, rgb, prompt_utils, elevation, azimuth, camera_distances, rgb_as_latents=False,
                 guidance_eval=False, **kwargs):
        batch_size = rgb.shape[0]
        rgb_BCHW = rgb.permute(0, 3, 1, 2)
        if rgb_as_latents:
            latents = ops.resize_bilinear(rgb_BCHW, (64, 64))      # F.interpolate(..., "bilinear", align_corners=False)
        else:
            latents = self.encode_images(ops.resize_bilinear(rgb_BCHW, (512, 512)))
        t = torch.randint(self.min_step, self.max_step + 1, [batch_size], dtype=torch.long, device=self.device,
                          generator=self.generator)
        grad, _ = self.compute_grad_sds(latents, t, prompt_utils, elevation, azimuth, camera_distances)
        grad = torch.nan_to_num(grad)
        if self.grad_clip_val is not None:
            grad = grad.clamp(-self.grad_clip_val, self.grad_clip_val)
        target = (latents - grad).detach()
        loss_sds = 0.5 * F.mse_loss(latents, target, reduction="sum") / batch_size
        return {"loss_sds": loss_sds, "grad_norm": grad.norm(), "min_step": self.min_step, "max_step": self.max_step}

    def update_step(self, epoch, global_step, on_load_weights=False):
        if self.cfg.grad_clip is not None:
            self.grad_clip_val = C(self.cfg.grad_clip, epoch, global_step)
        self.set_min_max_steps(min_step_percent=C(self.cfg.min_step_percent, epoch, global_step),
                               max_step_percent=C(self.cfg.max_step_percent, epoch, global_step))
