"""TEST INFRASTRUCTURE ONLY -- plain PyTorch restatement of the SD-2.1-base AutoencoderKL *encoder*
as the reference executes it through diffusers 0.19.0 in ``encode_images``
(Garment_3DGS/threestudio/models/guidance/stable_diffusion_guidance.py:160-167):

    imgs = imgs * 2 - 1
    latents = vae.encode(imgs.half()).latent_dist.sample() * vae.config.scaling_factor

and differentiated by autograd for the SDS loss (:424-427, `loss_sds.backward()` flows through
``latents`` into the rendered image).

PARITY UNPINNED: diffusers 0.19.0 (requirements.txt:12) is not under /root/reference and not
installable offline; the reference has no tests at this boundary. The restatement follows the
published module tree of AutoencoderKL(block_out_channels=(128,256,512,512), layers_per_block=2,
norm_num_groups=32, latent_channels=4, scaling_factor=0.18215): Encoder = conv_in, 4
DownEncoderBlock2D (ResnetBlock2D x2, eps 1e-6, no time embedding; Downsample2D = pad (0,1,0,1)
+ 3x3 stride-2 conv on the first three), UNetMidBlock2D (resnet, single-head 512-d attention with
GroupNorm + residual, resnet), GroupNorm + SiLU + conv_out(512 -> 8), quant_conv 1x1 (8 -> 8),
DiagonalGaussianDistribution (logvar clamped to [-30, 20]). It is checked structurally
(34.16 M encoder parameters, diffusers key scheme, output shape). The product kernels are compared
against THIS restatement and its autograd backward.
"""
import time

import torch
import torch.nn.functional as F

from .unet_ref import _conv, _linear, _norm

CH = (128, 256, 512, 512)
SCALING = 0.18215


def make_state_dict(seed=0):
    """Random-init encoder + quant_conv weights, diffusers key names, fp32 on CPU."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    _conv(sd, "encoder.conv_in", 3, CH[0], 3, g)
    cin = CH[0]
    for i, c in enumerate(CH):
        for j in range(2):
            p = f"encoder.down_blocks.{i}.resnets.{j}"
            rin = cin if j == 0 else c
            _norm(sd, p + ".norm1", rin, g); _conv(sd, p + ".conv1", rin, c, 3, g)
            _norm(sd, p + ".norm2", c, g); _conv(sd, p + ".conv2", c, c, 3, g)
            if rin != c:
                _conv(sd, p + ".conv_shortcut", rin, c, 1, g)
        if i < 3:
            _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3, g)
        cin = c
    for j in range(2):
        p = f"encoder.mid_block.resnets.{j}"
        _norm(sd, p + ".norm1", 512, g); _conv(sd, p + ".conv1", 512, 512, 3, g)
        _norm(sd, p + ".norm2", 512, g); _conv(sd, p + ".conv2", 512, 512, 3, g)
    a = "encoder.mid_block.attentions.0"
    _norm(sd, a + ".group_norm", 512, g)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _linear(sd, f"{a}.{n}", 512, 512, g)
    _norm(sd, "encoder.conv_norm_out", 512, g)
    _conv(sd, "encoder.conv_out", 512, 8, 3, g)
    _conv(sd, "quant_conv", 8, 8, 1, g)
    return sd


def param_count(sd):
    return sum(v.numel() for v in sd.values())


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def resnet(sd, p, x):
    h = F.conv2d(F.silu(_gn(sd, p + ".norm1", x)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(sd, p + ".norm2", h)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def mid_attention(sd, p, x):
    B, C, H, W = x.shape
    h = _gn(sd, p + ".group_norm", x).view(B, C, H * W).transpose(1, 2)
    lin = lambda n, t: F.linear(t, sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"])
    q, k, v = lin("to_q", h), lin("to_k", h), lin("to_v", h)
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]   # one head of 512
    return lin("to_out.0", o).transpose(1, 2).reshape(B, C, H, W) + x


def encoder_moments(sd, imgs_pm1):
    """imgs in [-1,1], [B,3,H,W] -> moments [B,8,H/8,W/8] (mean | logvar), dtype of sd."""
    x = F.conv2d(imgs_pm1.to(sd["encoder.conv_in.weight"].dtype), sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(4):
        for j in range(2):
            x = resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", x)
        if i < 3:
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            x = F.conv2d(F.pad(x, (0, 1, 0, 1)), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    x = resnet(sd, "encoder.mid_block.resnets.0", x)
    x = mid_attention(sd, "encoder.mid_block.attentions.0", x)
    x = resnet(sd, "encoder.mid_block.resnets.1", x)
    x = F.silu(_gn(sd, "encoder.conv_norm_out", x))
    x = F.conv2d(x, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(x, sd["quant_conv.weight"], sd["quant_conv.bias"])


def encode_images(sd, imgs01, noise):
    """encode_images of the reference: imgs in [0,1]; `noise` [B,4,h,w] replaces the sampler's
    randn so that both sides see the same draw. Returns fp32 latents (differentiable)."""
    moments = encoder_moments(sd, imgs01 * 2.0 - 1.0)
    mean, logvar = moments.chunk(2, dim=1)
    logvar = logvar.clamp(-30.0, 20.0)
    std = torch.exp(0.5 * logvar)
    return ((mean + std * noise.to(mean.dtype)) * SCALING).to(imgs01.dtype)


def encode_with_grad(sd, imgs01, noise, grad_latents):
    """(latents, d<latents, grad_latents>/d imgs01) by autograd."""
    x = imgs01.detach().clone().requires_grad_(True)
    lat = encode_images(sd, x, noise)
    lat.backward(grad_latents.to(lat.dtype))
    return lat.detach(), x.grad.detach()


_CPU_SD = None


def time_cpu_encode(batch=1, res=512):
    """Seconds for one fp32 eager encode + input-gradient backward on the host cores."""
    global _CPU_SD
    if _CPU_SD is None:
        _CPU_SD = make_state_dict(0)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(batch, 3, res, res, generator=g)
    n = torch.randn(batch, 4, res // 8, res // 8, generator=g)
    gl = torch.randn(batch, 4, res // 8, res // 8, generator=g)
    t0 = time.perf_counter()
    encode_with_grad(_CPU_SD, x, n, gl)
    return time.perf_counter() - t0
