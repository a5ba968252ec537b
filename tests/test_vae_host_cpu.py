"""CPU: the host-side weight re-layouts of VAEEncoderB200 (dgrad = flipped/transposed weights,
stride-2 conv on the space-to-depth tensor and its per-phase backward, conv_out+quant_conv fold,
conv_in as im2col / tap-product GEMMs) against torch conv2d + autograd, using a torch emulation of
the implicit-GEMM tap addressing of gd_unet_gemm (no compute kernels are called)."""
import torch
import torch.nn.functional as F

from garmentdreamer_b200 import vae as V
from oracle import vae_ref

T3 = [(t % 3 - 1, t // 3 - 1, 0) for t in range(9)]


def conv_taps_emul(x, w, taps, Ck, bias=None):
    """x [N,H,W,Cx]; w [Cout, ntaps*Ck]; tap (dx, dy, c): reads x[n, y+dy, x+dx, c:c+Ck], zero outside."""
    N, H, W, _ = x.shape
    out = torch.zeros(N, H, W, w.shape[0])
    for t, (dx, dy, c) in enumerate(taps):
        xs = torch.zeros(N, H, W, Ck)
        ys, xsl = slice(max(0, -dy), min(H, H - dy)), slice(max(0, -dx), min(W, W - dx))
        yd, xd = slice(max(0, -dy) + dy, min(H, H - dy) + dy), slice(max(0, -dx) + dx, min(W, W - dx) + dx)
        xs[:, ys, xsl] = x[:, yd, xd, c:c + Ck]
        out += xs @ w[:, t * Ck:(t + 1) * Ck].t()
    return out if bias is None else out + bias


def s2d(x):
    N, H, W, C = x.shape
    y = torch.zeros(N, H // 2, W // 2, 4 * C)
    for py in range(2):
        for px in range(2):
            y[..., (py * 2 + px) * C:(py * 2 + px + 1) * C] = x[:, py::2, px::2]
    return y


def d2s(y):
    N, Ho, Wo, C4 = y.shape
    C = C4 // 4
    x = torch.zeros(N, 2 * Ho, 2 * Wo, C)
    for py in range(2):
        for px in range(2):
            x[:, py::2, px::2] = y[..., (py * 2 + px) * C:(py * 2 + px + 1) * C]
    return x


def _enc():
    sd = vae_ref.make_state_dict(0)
    return sd, V.VAEEncoderB200(sd, "cpu")


def test_weight_relayouts_match_conv2d_and_autograd():
    sd, enc = _enc()
    g = torch.Generator().manual_seed(0)
    nchw = lambda t: t.permute(0, 3, 1, 2)
    # stride-2 downsample with padding (0,1,0,1): forward on the space-to-depth tensor, backward per phase
    p, C = "encoder.down_blocks.0.downsamplers.0.conv", 128
    x = torch.randn(1, 8, 8, C, generator=g)
    xr = nchw(x).clone().requires_grad_(True)
    ref = F.conv2d(F.pad(xr, (0, 1, 0, 1)), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    out = conv_taps_emul(s2d(x), enc.w[p + ".fwd"].float(), [(dx, dy, ph * C) for dx, dy, ph in V._DOWN_TAPS], C, enc.w[p + ".bias"].float())
    assert (nchw(out) - ref).abs().max() < 2e-3
    dy_ = torch.randn(1, 4, 4, C, generator=g)
    ref.backward(nchw(dy_))
    ds = torch.zeros(1, 4, 4, 4 * C)
    for ph in range(4):
        taps = [(-(kx >> 1), -(ky >> 1), 0) for ky in range(3) for kx in range(3) if (ky & 1) * 2 + (kx & 1) == ph]
        ds[..., ph * C:(ph + 1) * C] = conv_taps_emul(dy_, enc.w[f"{p}.bwd{ph}"].float(), taps, C)
    assert (nchw(d2s(ds)) - xr.grad).abs().max() < 2e-3
    # 3x3 conv + dgrad with flipped, transposed weights (Cin != Cout)
    p = "encoder.down_blocks.1.resnets.0.conv1"
    x = torch.randn(1, 8, 8, 128, generator=g)
    xr = nchw(x).clone().requires_grad_(True)
    ref = F.conv2d(xr, sd[p + ".weight"], sd[p + ".bias"], padding=1)
    assert (nchw(conv_taps_emul(x, enc.w[p + ".fwd"].float(), T3, 128, enc.w[p + ".bias"].float())) - ref).abs().max() < 2e-3
    dy_ = torch.randn(1, 8, 8, 256, generator=g)
    ref.backward(nchw(dy_))
    assert (nchw(conv_taps_emul(dy_, enc.w[p + ".bwd"].float(), T3, 256)) - xr.grad).abs().max() < 4e-3
    # conv_out folded with quant_conv, and its dgrad on the 64-channel padded moments gradient
    x = torch.randn(1, 8, 8, 512, generator=g)
    xr = nchw(x).clone().requires_grad_(True)
    ref = F.conv2d(F.conv2d(xr, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1), sd["quant_conv.weight"], sd["quant_conv.bias"])
    assert (nchw(conv_taps_emul(x, enc.w["conv_out.fwd"].float(), T3, 512, enc.w["conv_out.bias"].float())) - ref).abs().max() < 2e-3
    dm = torch.zeros(1, 8, 8, 64)
    dm[..., :8] = torch.randn(1, 8, 8, 8, generator=g)
    ref.backward(nchw(dm[..., :8]))
    assert (nchw(conv_taps_emul(dm, enc.w["conv_out.bwd"].float(), T3, 64)) - xr.grad).abs().max() < 1e-3


def test_conv_in_im2col_and_tap_gather():
    sd, enc = _enc()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, 8, 8, generator=g).requires_grad_(True)
    ref = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    # im2col rows as gd_vae_im2col writes them: column (ky*3+kx)*3+c
    xp = F.pad(x.detach(), (1, 1, 1, 1))
    cols = torch.zeros(64, 64)
    for ky in range(3):
        for kx in range(3):
            for c in range(3):
                cols[:, (ky * 3 + kx) * 3 + c] = xp[0, c, ky:ky + 8, kx:kx + 8].reshape(-1)
    out = cols @ enc.w["conv_in.fwd"].float().t() + enc.w["encoder.conv_in.bias"].float()
    assert (out.view(8, 8, 128).permute(2, 0, 1) - ref[0]).abs().max() < 2e-3
    # data gradient: per-pixel tap products, then the 9-tap gather of gd_vae_dimg_gather
    dy_ = torch.randn(1, 8, 8, 128, generator=g)
    ref.backward(dy_.permute(0, 3, 1, 2))
    z = (dy_.view(64, 128) @ enc.w["conv_in.bwd"].float().t()).view(8, 8, 32)
    dx = torch.zeros(3, 8, 8)
    for y in range(8):
        for xx in range(8):
            for ky in range(3):
                for kx in range(3):
                    yy, xs = y - ky + 1, xx - kx + 1
                    if 0 <= yy < 8 and 0 <= xs < 8:
                        dx[:, y, xx] += z[yy, xs, (ky * 3 + kx) * 3:(ky * 3 + kx) * 3 + 3]
    assert (dx - x.grad[0]).abs().max() < 5e-3 * x.grad.abs().max()


def test_param_schema_matches_restatement():
    from garmentdreamer_b200 import unet_init
    a, b = unet_init.vae_encoder_param_shapes(), vae_ref.make_state_dict(0)
    assert set(a) == set(b) and all(tuple(b[k].shape) == tuple(a[k]) for k in a)
    assert vae_ref.param_count(b) == 34163664
