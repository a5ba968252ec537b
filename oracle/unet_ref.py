"""TEST INFRASTRUCTURE ONLY -- plain PyTorch restatement of the SD-2.1-base UNet2DConditionModel
forward as the reference executes it through diffusers 0.19.0
(Garment_3DGS/threestudio/models/guidance/stable_diffusion_guidance.py:146-157).

PARITY UNPINNED: diffusers 0.19.0 (requirements.txt:12) is neither vendored under /root/reference
nor installed here, and the reference has no tests at this boundary, so this restatement is
checked only structurally (865.9 M parameters, diffusers state-dict key scheme, output shape) and
against the in-tree top-level forward NETF/netf/vsd/lora_unet.py:553-744 (block order, skip
connections, time embedding). Config values are those of stabilityai/stable-diffusion-2-1-base
(SURVEY.md Appendix B). The product kernels are compared against THIS restatement.
"""
import math
import time

import torch
import torch.nn.functional as F

CFG = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
           cross_attention_dim=1024, heads=(5, 10, 20, 20), norm_groups=32, norm_eps=1e-5,
           down_attn=(True, True, True, False), up_attn=(False, True, True, True))


# ---- parameter construction (diffusers key names, PyTorch default initialisers) -------------
def _linear(sd, name, cin, cout, g, bias=True):
    bound = 1.0 / math.sqrt(cin)
    sd[name + ".weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * bound
    if bias:
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound


def _conv(sd, name, cin, cout, k, g):
    bound = 1.0 / math.sqrt(cin * k * k)
    sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
    sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound


def _norm(sd, name, c, g):
    # PyTorch initialises affine norms to (1, 0); a small seeded perturbation keeps the parity
    # tests sensitive to gamma/beta handling
    sd[name + ".weight"] = 1.0 + 0.05 * torch.randn(c, generator=g)
    sd[name + ".bias"] = 0.05 * torch.randn(c, generator=g)


def _resnet(sd, p, cin, cout, g):
    _norm(sd, p + ".norm1", cin, g); _conv(sd, p + ".conv1", cin, cout, 3, g)
    _linear(sd, p + ".time_emb_proj", 1280, cout, g)
    _norm(sd, p + ".norm2", cout, g); _conv(sd, p + ".conv2", cout, cout, 3, g)
    if cin != cout:
        _conv(sd, p + ".conv_shortcut", cin, cout, 1, g)


def _transformer(sd, p, c, ctx, g):
    _norm(sd, p + ".norm", c, g)
    _linear(sd, p + ".proj_in", c, c, g)
    b = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        _norm(sd, f"{b}.{n}", c, g)
    for attn, kv in (("attn1", c), ("attn2", ctx)):
        _linear(sd, f"{b}.{attn}.to_q", c, c, g, bias=False)
        _linear(sd, f"{b}.{attn}.to_k", kv, c, g, bias=False)
        _linear(sd, f"{b}.{attn}.to_v", kv, c, g, bias=False)
        _linear(sd, f"{b}.{attn}.to_out.0", c, c, g)
    _linear(sd, f"{b}.ff.net.0.proj", c, 8 * c, g)
    _linear(sd, f"{b}.ff.net.2", 4 * c, c, g)
    _linear(sd, p + ".proj_out", c, c, g)


def make_state_dict(seed=0):
    """Random-init weights with the diffusers key scheme, fp32 on CPU (865.9 M parameters)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    ch = CFG["block_out_channels"]
    ctx = CFG["cross_attention_dim"]
    _conv(sd, "conv_in", 4, ch[0], 3, g)
    _linear(sd, "time_embedding.linear_1", ch[0], 1280, g)
    _linear(sd, "time_embedding.linear_2", 1280, 1280, g)
    out_c = ch[0]
    for i, c in enumerate(ch):
        in_c, out_c = out_c, c
        for j in range(2):
            _resnet(sd, f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c, g)
            if CFG["down_attn"][i]:
                _transformer(sd, f"down_blocks.{i}.attentions.{j}", out_c, ctx, g)
        if i < 3:
            _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3, g)
    _resnet(sd, "mid_block.resnets.0", 1280, 1280, g)
    _transformer(sd, "mid_block.attentions.0", 1280, ctx, g)
    _resnet(sd, "mid_block.resnets.1", 1280, 1280, g)
    rev = list(reversed(ch))
    out_c = rev[0]
    for i in range(4):
        prev, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, 3)]
        for j in range(3):
            skip = in_c if j == 2 else out_c
            rin = prev if j == 0 else out_c
            _resnet(sd, f"up_blocks.{i}.resnets.{j}", rin + skip, out_c, g)
            if CFG["up_attn"][i]:
                _transformer(sd, f"up_blocks.{i}.attentions.{j}", out_c, ctx, g)
        if i < 3:
            _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3, g)
    _norm(sd, "conv_norm_out", ch[0], g)
    _conv(sd, "conv_out", ch[0], 4, 3, g)
    return sd


def param_count(sd):
    return sum(v.numel() for v in sd.values())


# ---- forward -------------------------------------------------------------------------------
def timestep_embedding(t, dim=320):
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t.float()[:, None] * freq[None]
    return torch.cat([torch.cos(a), torch.sin(a)], -1)


def _gn(sd, p, x, eps, groups=32):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def resnet(sd, p, x, emb):
    h = F.conv2d(F.silu(_gn(sd, p + ".norm1", x, 1e-5)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = h + _lin(sd, p + ".time_emb_proj", F.silu(emb))[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, p + ".norm2", h, 1e-5)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def attention(sd, p, x, ctx, heads):
    B, T, C = x.shape
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    sp = lambda t: t.view(B, -1, heads, C // heads).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))  # AttnProcessor2_0, scale 1/sqrt(64)
    return _lin(sd, p + ".to_out.0", o.transpose(1, 2).reshape(B, T, C))


def transformer(sd, p, x, ctx, heads):
    B, C, H, W = x.shape
    res = x
    h = _gn(sd, p + ".norm", x, 1e-6).permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = _lin(sd, p + ".proj_in", h)
    b = p + ".transformer_blocks.0"
    ln = lambda n, t: F.layer_norm(t, (C,), sd[f"{b}.{n}.weight"], sd[f"{b}.{n}.bias"], 1e-5)
    h = attention(sd, b + ".attn1", ln("norm1", h), ln("norm1", h), heads) + h
    h = attention(sd, b + ".attn2", ln("norm2", h), ctx, heads) + h
    f = _lin(sd, b + ".ff.net.0.proj", ln("norm3", h))
    val, gate = f.chunk(2, -1)
    h = _lin(sd, b + ".ff.net.2", val * F.gelu(gate)) + h
    h = _lin(sd, p + ".proj_out", h).reshape(B, H, W, C).permute(0, 3, 1, 2)
    return h + res


def unet_forward(sd, sample, timestep, encoder_hidden_states, trace=None):
    """sample [B,4,H,W], timestep [B], encoder_hidden_states [B,77,1024]; dtype = that of sd.
    trace: optional callable(name, output NCHW) invoked after every block (error-growth tables)."""
    tr = trace if trace is not None else (lambda name, out: None)
    dt = sd["conv_in.weight"].dtype
    x = sample.to(dt)
    ctx = encoder_hidden_states.to(dt)
    temb = timestep_embedding(timestep.to(dt)).to(dt)  # timesteps arrive in weights dtype (:155)
    emb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", temb)))
    x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    tr("conv_in", x)
    skips = [x]
    for i in range(4):
        for j in range(2):
            x = resnet(sd, f"down_blocks.{i}.resnets.{j}", x, emb)
            tr(f"down_blocks.{i}.resnets.{j}", x)
            if CFG["down_attn"][i]:
                x = transformer(sd, f"down_blocks.{i}.attentions.{j}", x, ctx, CFG["heads"][i])
                tr(f"down_blocks.{i}.attentions.{j}", x)
            skips.append(x)
        if i < 3:
            p = f"down_blocks.{i}.downsamplers.0.conv"
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=2, padding=1)
            tr(p, x)
            skips.append(x)
    x = resnet(sd, "mid_block.resnets.0", x, emb)
    tr("mid_block.resnets.0", x)
    x = transformer(sd, "mid_block.attentions.0", x, ctx, 20)
    tr("mid_block.attentions.0", x)
    x = resnet(sd, "mid_block.resnets.1", x, emb)
    tr("mid_block.resnets.1", x)
    rev_heads = list(reversed(CFG["heads"]))
    for i in range(4):
        for j in range(3):
            x = torch.cat([x, skips.pop()], 1)
            x = resnet(sd, f"up_blocks.{i}.resnets.{j}", x, emb)
            tr(f"up_blocks.{i}.resnets.{j}", x)
            if CFG["up_attn"][i]:
                x = transformer(sd, f"up_blocks.{i}.attentions.{j}", x, ctx, rev_heads[i])
                tr(f"up_blocks.{i}.attentions.{j}", x)
        if i < 3:
            p = f"up_blocks.{i}.upsamplers.0.conv"
            x = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), sd[p + ".weight"], sd[p + ".bias"], padding=1)
            tr(p, x)
    x = F.silu(_gn(sd, "conv_norm_out", x, 1e-5))
    x = F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    tr("conv_out", x)
    return x


def alphas_cumprod(n=1000, beta_start=0.00085, beta_end=0.012):
    """DDIMScheduler (scaled_linear) alphas_cumprod as used at stable_diffusion_guidance.py:129."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, 0)


_CPU_SD = None


def time_cpu_forward(batch=2, latent=64):
    """Seconds for one fp32 eager forward of `batch` samples on the host cores (bench baseline)."""
    global _CPU_SD
    if _CPU_SD is None:
        _CPU_SD = make_state_dict(0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 4, latent, latent, generator=g)
    t = torch.randint(20, 981, (batch,), generator=g)
    ctx = torch.randn(batch, 77, 1024, generator=g)
    with torch.no_grad():
        t0 = time.perf_counter()
        unet_forward(_CPU_SD, x, t, ctx)
        return time.perf_counter() - t0
