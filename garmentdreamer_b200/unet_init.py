"""Parameter schema of the SD-2.1-base UNet2DConditionModel (diffusers state-dict key scheme) and
a random initialiser for benchmarks / smoke tests (there is no network for checkpoints).
Module tree as in Garment_Deformer_NeTF/netf/vsd/lora_unet.py:119-422 with the SD-2.1-base config
(SURVEY.md Appendix B): 865,910,724 parameters."""
import math
from collections import OrderedDict

import torch

CH = (320, 640, 1280, 1280)
DOWN_ATTN = (True, True, True, False)
UP_ATTN = (False, True, True, True)
CTX = 1024


def unet_param_shapes():
    s = OrderedDict()

    def lin(n, cin, cout, bias=True):
        s[n + ".weight"] = (cout, cin)
        if bias:
            s[n + ".bias"] = (cout,)

    def conv(n, cin, cout, k):
        s[n + ".weight"] = (cout, cin, k, k)
        s[n + ".bias"] = (cout,)

    def norm(n, c):
        s[n + ".weight"] = (c,)
        s[n + ".bias"] = (c,)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin); conv(p + ".conv1", cin, cout, 3); lin(p + ".time_emb_proj", 1280, cout)
        norm(p + ".norm2", cout); conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cin, cout, 1)

    def xf(p, c):
        norm(p + ".norm", c); lin(p + ".proj_in", c, c)
        b = p + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{b}.{n}", c)
        for a, kv in (("attn1", c), ("attn2", CTX)):
            lin(f"{b}.{a}.to_q", c, c, False); lin(f"{b}.{a}.to_k", kv, c, False); lin(f"{b}.{a}.to_v", kv, c, False)
            lin(f"{b}.{a}.to_out.0", c, c)
        lin(f"{b}.ff.net.0.proj", c, 8 * c); lin(f"{b}.ff.net.2", 4 * c, c); lin(p + ".proj_out", c, c)

    conv("conv_in", 4, CH[0], 3)
    lin("time_embedding.linear_1", CH[0], 1280); lin("time_embedding.linear_2", 1280, 1280)
    out_c = CH[0]
    for i, c in enumerate(CH):
        in_c, out_c = out_c, c
        for j in range(2):
            resnet(f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
            if DOWN_ATTN[i]:
                xf(f"down_blocks.{i}.attentions.{j}", out_c)
        if i < 3:
            conv(f"down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    resnet("mid_block.resnets.0", 1280, 1280); xf("mid_block.attentions.0", 1280); resnet("mid_block.resnets.1", 1280, 1280)
    rev = CH[::-1]
    out_c = rev[0]
    for i in range(4):
        prev, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, 3)]
        for j in range(3):
            resnet(f"up_blocks.{i}.resnets.{j}", (prev if j == 0 else out_c) + (in_c if j == 2 else out_c), out_c)
            if UP_ATTN[i]:
                xf(f"up_blocks.{i}.attentions.{j}", out_c)
        if i < 3:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    norm("conv_norm_out", CH[0]); conv("conv_out", CH[0], 4, 3)
    return s


def random_state_dict(seed=0, device="cpu", dtype=torch.float16):
    """PyTorch-default style init (U(+-1/sqrt(fan_in)) for conv / linear, ~(1, 0) for norms)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    shapes = unet_param_shapes()
    for name, shp in shapes.items():
        is_norm = ".norm" in name or name.startswith("conv_norm_out")
        if is_norm:
            v = (1.0 if name.endswith("weight") else 0.0) + 0.05 * torch.randn(shp, generator=g)
        else:
            wshape = shapes[name[:-4] + "weight"] if name.endswith("bias") else shp
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            v = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
        sd[name] = v.to(device=device, dtype=dtype)
    return sd


# ---- SD-2.1 AutoencoderKL encoder (+ quant_conv): 34,163,664 parameters -------------------------
VAE_CH = (128, 256, 512, 512)


def vae_encoder_param_shapes():
    """diffusers key scheme of AutoencoderKL.encoder + quant_conv (block_out_channels
    (128,256,512,512), layers_per_block 2, latent_channels 4)."""
    s = OrderedDict()

    def conv(n, cin, cout, k):
        s[n + ".weight"] = (cout, cin, k, k)
        s[n + ".bias"] = (cout,)

    def norm(n, c):
        s[n + ".weight"] = (c,)
        s[n + ".bias"] = (c,)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin); conv(p + ".conv1", cin, cout, 3)
        norm(p + ".norm2", cout); conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cin, cout, 1)

    conv("encoder.conv_in", 3, VAE_CH[0], 3)
    cin = VAE_CH[0]
    for i, c in enumerate(VAE_CH):
        for j in range(2):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c)
        if i < 3:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
        cin = c
    resnet("encoder.mid_block.resnets.0", 512, 512)
    a = "encoder.mid_block.attentions.0"
    norm(a + ".group_norm", 512)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[f"{a}.{n}.weight"] = (512, 512)
        s[f"{a}.{n}.bias"] = (512,)
    resnet("encoder.mid_block.resnets.1", 512, 512)
    norm("encoder.conv_norm_out", 512)
    conv("encoder.conv_out", 512, 8, 3)
    conv("quant_conv", 8, 8, 1)
    return s


def random_vae_state_dict(seed=0, device="cpu", dtype=torch.float32):
    g = torch.Generator().manual_seed(1000 + seed)
    sd = OrderedDict()
    shapes = vae_encoder_param_shapes()
    for name, shp in shapes.items():
        if "norm" in name:
            v = (1.0 if name.endswith("weight") else 0.0) + 0.05 * torch.randn(shp, generator=g)
        else:
            wshape = shapes[name[:-4] + "weight"] if name.endswith("bias") else shp
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            v = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
        sd[name] = v.to(device=device, dtype=dtype)
    return sd
