"""GPU: every UNet operator of the C ABI (include/gd_unet.h) against a plain PyTorch fp32
reference of the same op on the same fp16 inputs. Tolerance: 2e-3 relative (fp16 output rounding
is 4.9e-4; north_star allows 1e-3 rel on the final SDS gradient, checked in test_unet_gpu.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def maxrel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


@pytest.fixture(scope="module")
def ops():
    from garmentdreamer_b200 import unet_ops
    unet_ops.lib()
    return unet_ops


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g, device="cuda") * scale).half()


@pytest.mark.parametrize("M,K,N", [(128, 64, 16), (256, 320, 320), (1000, 1280, 640), (4096, 320, 2560),
                                   (77 * 2, 1024, 1280), (64, 1280, 1280), (8192, 2560, 320), (512, 1280, 5120), (300, 640, 2560)])
def test_linear(ops, M, K, N):
    x, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N)
    y = ops.linear(x, w, b)
    ref = x.float() @ w.float().t() + b.float()
    assert maxrel(y, ref) < 2e-3 and rel(y, ref) < 1e-3


def test_linear_epilogues(ops):
    M, K, N = 512, 640, 640
    x, w, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    ref = x.float() @ w.float().t() + b.float()
    assert rel(ops.linear(x, w, b, residual=r), ref + r.float()) < 1e-3
    assert rel(ops.linear(x, w, b, flags=ops.EPI_SILU), F.silu(ref)) < 1e-3
    assert rel(ops.linear(x, w, None, alpha=0.125), 0.125 * (x.float() @ w.float().t())) < 1e-3
    # GEGLU: weight rows interleaved in blocks of 16 (value, gate)
    H = N // 2
    val, gate = ref[:, :H], ref[:, H:]
    perm = torch.cat([torch.cat([torch.arange(i, i + 16), torch.arange(H + i, H + i + 16)]) for i in range(0, H, 16)]).cuda()
    y = ops.linear(x, w[perm].contiguous(), b[perm].contiguous(), flags=ops.EPI_GEGLU)
    assert y.shape == (M, H)
    assert rel(y, val * F.gelu(gate)) < 1e-3
    assert rel(ops.geglu(ref.half()), ref.half().float()[:, :H] * F.gelu(ref.half().float()[:, H:])) < 1e-3


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 320), (3, 16, 16, 1280, 640),
                                            (2, 8, 8, 1280, 1280), (1, 8, 8, 2560, 1280), (1, 64, 64, 64, 128),
                                            (8, 8, 8, 1280, 1280), (8, 8, 8, 2560, 1280)])  # last two: split-K path
def test_conv3x3(ops, N, H, W, Cin, Cout):
    x = rnd(N, H, W, Cin)
    w = rnd(Cout, 3, 3, Cin, scale=(9 * Cin) ** -0.5)
    b, temb, res = rnd(Cout), rnd(N, Cout), rnd(N, H, W, Cout)
    y = ops.conv3x3(x, w, b, row_bias=temb, residual=res)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b.float(), padding=1)
    ref = (ref + temb.float()[:, :, None, None]).permute(0, 2, 3, 1) + res.float()
    assert maxrel(y, ref) < 2e-3 and rel(y, ref) < 1e-3


def test_conv3x3_patch_tiles_odd_tile_count_stays_in_bounds(ops):
    """Patch tiles (8 px x 16 rows) with an ODD number of tiles: the second CTA of the last pair owns a tile that does not exist
    and must neither store nor count statistics for it (guard words behind the output stay untouched)."""
    N, H, W, Cin, Cout = 3, 16, 8, 128, 64
    x, w, b = rnd(N, H, W, Cin), rnd(Cout, 3, 3, Cin, scale=(9 * Cin) ** -0.5), rnd(Cout)
    buf = torch.full((N * H * W * Cout + 4096,), 7.0, dtype=torch.float16, device="cuda")
    out = buf[:N * H * W * Cout].view(N, H, W, Cout)
    y = ops.conv3x3(x, w, b, out=out, want_stats=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b.float(), padding=1).permute(0, 2, 3, 1)
    assert rel(y, ref) < 1e-3
    assert bool((buf[N * H * W * Cout:] == 7.0).all()), "stores behind the end of the output tensor"


@pytest.mark.parametrize("N,H,W,C", [(2, 64, 64, 320), (2, 32, 32, 640), (2, 16, 16, 1280)])
def test_conv3x3_stride2(ops, N, H, W, C):
    x, w, b = rnd(N, H, W, C), rnd(C, 3, 3, C, scale=(9 * C) ** -0.5), rnd(C)
    y = ops.conv3x3_stride2(x, w, b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b.float(), stride=2, padding=1)
    assert rel(y, ref.permute(0, 2, 3, 1)) < 1e-3


@pytest.mark.parametrize("B,Tq,Tk,heads", [(2, 4096, 4096, 5), (2, 1024, 1024, 10), (2, 256, 77, 20), (1, 64, 64, 20), (3, 256, 256, 20), (2, 4096, 77, 5), (1, 1024, 200, 10)])
def test_attention(ops, B, Tq, Tk, heads):
    C = heads * 64
    q, k, xv = rnd(B, Tq, C), rnd(B, Tk, C, seed=1), rnd(B, Tk, C, seed=2)
    wv = rnd(C, C, scale=C ** -0.5)
    ldv = (Tk + 7) // 8 * 8
    vt = ops.linear_transposed(xv, wv, ldv)
    v_ref = (xv.float() @ wv.float().t())
    assert rel(vt[:, :, :Tk].transpose(1, 2), v_ref) < 1e-3
    s = ops.attn_scores(q, k, heads, 0.125)
    qh = q.float().view(B, Tq, heads, 64).permute(0, 2, 1, 3)
    kh = k.float().view(B, Tk, heads, 64).permute(0, 2, 1, 3)
    s_ref = 0.125 * qh @ kh.transpose(-1, -2)
    assert rel(s[:, :, :Tk].view(B, heads, Tq, Tk), s_ref) < 1e-3
    ops.softmax_(s, Tk)
    assert rel(s[:, :, :Tk].view(B, heads, Tq, Tk), s_ref.softmax(-1)) < 2e-3
    o = torch.empty(B, Tq, C, dtype=torch.float16, device="cuda")
    ops.attn_values(s, vt, heads, Tk, o)
    vh = vt[:, :, :Tk].float().view(B, heads, 64, Tk).transpose(-1, -2)
    o_ref = (s_ref.softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(B, Tq, C)
    assert rel(o, o_ref) < 3e-3
    # fused kernel (the one the UNet uses): same result without materialising the scores
    of = ops.flash_attention(q, k, vt, heads, Tk, 0.125)
    assert rel(of, o_ref) < 3e-3 and maxrel(of, o_ref) < 1e-2


@pytest.mark.parametrize("B,Tq,Tk,heads,gain", [(1, 384, 300, 5, 1.0), (2, 1024, 1024, 10, 3.0), (1, 4096, 4096, 5, 4.0), (1, 200, 640, 2, 2.0)])
def test_attention_single_pass_lazy_rescale(ops, B, Tq, Tk, heads, gain):
    """k_flash_attn2 (key ranges of >= 2 tiles): ragged Tq / Tk, and logits whose row maxima keep growing along the key
    axis (keys sorted by norm, large gain) so the lazy running maximum moves and O is rescaled in TMEM several times."""
    C = heads * 64
    q, k, v = rnd(B, Tq, C, scale=gain), rnd(B, Tk, C, seed=1), rnd(B, Tk, C, seed=2)
    ramp = torch.linspace(0.2, gain, Tk, device="cuda").view(1, Tk, 1)
    k = (k.float() * ramp).half()
    ldv = (Tk + 7) // 8 * 8
    vt = torch.zeros(B, C, ldv, dtype=torch.float16, device="cuda")
    vt[:, :, :Tk] = v.transpose(1, 2)
    qh = q.float().view(B, Tq, heads, 64).permute(0, 2, 1, 3)
    kh = k.float().view(B, Tk, heads, 64).permute(0, 2, 1, 3)
    vh = v.float().view(B, Tk, heads, 64).permute(0, 2, 1, 3)
    o_ref = ((0.125 * qh @ kh.transpose(-1, -2)).softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(B, Tq, C)
    o = ops.flash_attention(q, k, vt, heads, Tk, 0.125)
    assert torch.isfinite(o).all()
    assert rel(o, o_ref) < 3e-3 and maxrel(o, o_ref) < 1e-2


def test_norms_and_elementwise(ops):
    x = rnd(2, 32, 32, 640)
    g, b = rnd(640), rnd(640)
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), 32, g.float(), b.float(), 1e-5)
    assert rel(ops.groupnorm(x, g, b, eps=1e-5), ref.permute(0, 2, 3, 1)) < 1e-3
    assert rel(ops.groupnorm(x, g, b, eps=1e-5, silu=True), F.silu(ref).permute(0, 2, 3, 1)) < 1e-3
    x320 = rnd(1, 64, 64, 320)
    g3, b3 = rnd(320), rnd(320)
    ref = F.group_norm(x320.float().permute(0, 3, 1, 2), 32, g3.float(), b3.float(), 1e-6)
    assert rel(ops.groupnorm(x320, g3, b3, eps=1e-6), ref.permute(0, 2, 3, 1)) < 1e-3
    t = rnd(300, 1280)
    g2, b2 = rnd(1280), rnd(1280)
    assert rel(ops.layernorm(t, g2, b2), F.layer_norm(t.float(), (1280,), g2.float(), b2.float(), 1e-5)) < 1e-3
    a, c = rnd(1000, 64), rnd(1000, 64, seed=3)
    assert torch.equal(ops.add(a, c), a + c)
    u = rnd(2, 8, 8, 1280)
    assert torch.equal(ops.upsample2x(u), F.interpolate(u.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1))
    p, q = rnd(2, 16, 16, 640), rnd(2, 16, 16, 1280)
    assert torch.equal(ops.concat(p, q), torch.cat([p, q], -1))


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 32, 32, 320, 640), (8, 16, 16, 1280, 1280), (1, 64, 64, 320, 320), (2, 128, 128, 128, 128)])
def test_gemm_colstats_feed_groupnorm(ops, N, H, W, Cin, Cout):
    """GroupNorm statistics emitted by the producing GEMM epilogue (GdGemmArgs.colstats): the per-32-row column sums match
    the stored fp16 output, and groupnorm() consuming them equals groupnorm() with its own statistics pass."""
    x, w, b = rnd(N, H, W, Cin), rnd(Cout, 3, 3, Cin, scale=(9 * Cin) ** -0.5), rnd(Cout)
    y = ops.conv3x3(x, w, b, want_stats=True)
    cs = getattr(y, "_gd_colstats", None)
    assert cs is not None and cs[0] is not None, "the plain conv epilogue must produce column statistics"
    st = cs[0]
    assert st.shape == (N * H * W // 32, 2, Cout)
    # the blocks partition each image's pixels (which 32 pixels form a block depends on the tile shape): per-image totals
    per_img = st.view(N, H * W // 32, 2, Cout).sum(1)
    yf = y.float().view(N, H * W, Cout)
    assert (per_img[:, 0] - yf.sum(1)).abs().max() < 1e-3 * max(1.0, float(yf.sum(1).abs().max()))
    assert (per_img[:, 1] - (yf * yf).sum(1)).abs().max() < 1e-3 * float((yf * yf).sum(1).abs().max())
    g, be = rnd(Cout), rnd(Cout)
    fused = ops.groupnorm(y, g, be, eps=1e-5, silu=True)
    y2 = y.clone()
    plain = ops.groupnorm(y2, g, be, eps=1e-5, silu=True)
    ref = F.silu(F.group_norm(y.float().permute(0, 3, 1, 2), 32, g.float(), be.float(), 1e-5)).permute(0, 2, 3, 1)
    assert rel(fused, ref) < 1e-3 and rel(fused, plain) < 2e-4
    # concatenated sources (UNet up path): both producers' statistics are used
    y3 = ops.conv3x3(x, w, b, want_stats=True)
    cat = ops.concat(y, y3)
    assert getattr(cat, "_gd_colstats", None) is not None
    g2, b2 = rnd(2 * Cout), rnd(2 * Cout)
    fused = ops.groupnorm(cat, g2, b2, eps=1e-5, silu=True)
    ref = F.silu(F.group_norm(cat.float().permute(0, 3, 1, 2), 32, g2.float(), b2.float(), 1e-5)).permute(0, 2, 3, 1)
    assert rel(fused, ref) < 1e-3
    # VAE flavour: (mean, rstd) table for the backward
    out, stats = ops.groupnorm_stats(y, g, be, eps=1e-6, silu=True)
    out2, stats2 = ops.groupnorm_stats(y.clone(), g, be, eps=1e-6, silu=True)
    assert rel(out, out2) < 2e-4 and rel(stats, stats2) < 1e-4


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 128, 128, 128, 128), (2, 64, 64, 256, 256), (4, 64, 64, 512, 256), (1, 128, 128, 256, 256),
                                            (1, 32, 32, 64, 512), (1, 16, 16, 512, 512)])
def test_gemm_groupnorm_backward_producer_epilogue(ops, N, H, W, Cin, Cout):
    """Data-gradient conv whose epilogue multiplies by silu'(GN(x)) and emits sum g | sum g*xh (GdGemmArgs.gn_coef), followed by
    groupnorm_bwd on those sums, against the unfused conv + two-sweep GroupNorm backward and against torch autograd. The last
    shape is too small for the fused epilogue (split-K): the call must fall back and still be right."""
    x = rnd(N, H, W, Cout)                                  # GroupNorm input (forward activation)
    gam, bet = rnd(Cout), rnd(Cout)
    dz_in, w = rnd(N, H, W, Cin, seed=5), rnd(Cout, 3, 3, Cin, scale=(9 * Cin) ** -0.5, seed=6)
    add = rnd(N, H, W, Cout, seed=7)
    _, stats = ops.groupnorm_stats(x, gam, bet, eps=1e-6, silu=True)
    plain = ops.conv3x3(dz_in, w)
    d_plain = ops.groupnorm_bwd(x, plain, gam, bet, stats, silu=True, add=add)
    g = ops.conv3x3(dz_in, w, gn_bwd=(x, stats, gam, bet))
    fused_taken = bool(getattr(g, "_gd_is_g", False))
    # (1,128,128,256->256) runs a 256-wide tile with two staging tiles per warp and few ring stages; (1,32,32,64->512) has 9 k-blocks only
    assert fused_taken == (N * H * W >= 1024), "fused epilogue expected for every shape but the split-K one"
    d_fused = ops.groupnorm_bwd(x, g, gam, bet, stats, silu=True, add=add)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    z = F.silu(F.group_norm(xr, 32, gam.float(), bet.float(), 1e-6))
    z.backward(plain.float().permute(0, 3, 1, 2))
    ref = xr.grad.permute(0, 2, 3, 1) + add.float()
    assert rel(d_plain, ref) < 2e-3
    assert rel(d_fused, ref) < 2e-3 and rel(d_fused, d_plain) < 2e-3


def test_time_embedding_and_small_ops(ops):
    t = torch.tensor([20.0, 500.0, 979.0, 980.0], device="cuda")
    e = ops.timestep_embedding(t, 320)
    th = t.half().float()
    freq = torch.exp(-math.log(10000.0) * torch.arange(160, device="cuda").float() / 160)
    ref = torch.cat([torch.cos(th[:, None] * freq), torch.sin(th[:, None] * freq)], -1)
    assert (e.float() - ref).abs().max() < 2e-3
    x, w, b = rnd(8, 1280), rnd(640, 1280, scale=1280 ** -0.5), rnd(640)
    ref = F.silu(x.float()) @ w.float().t() + b.float()
    assert rel(ops.small_linear(x, w, b, silu_in=True), ref) < 1e-3
    assert rel(ops.small_linear(x, w, b, silu_out=True), F.silu(x.float() @ w.float().t() + b.float())) < 1e-3
    xi = rnd(2, 4, 64, 64)
    wi, bi = rnd(320, 3, 3, 4, scale=1 / 6), rnd(320)
    ref = F.conv2d(xi.float(), wi.float().permute(0, 3, 1, 2), bi.float(), padding=1).permute(0, 2, 3, 1)
    assert rel(ops.conv_in(xi, wi, bi), ref) < 1e-3
    xo = rnd(2, 64, 64, 320)
    wo, bo = rnd(4, 3, 3, 320, scale=(9 * 320) ** -0.5), rnd(4)
    ref = F.conv2d(xo.float().permute(0, 3, 1, 2), wo.float().permute(0, 3, 1, 2), bo.float(), padding=1)
    assert rel(ops.conv_out(xo, wo, bo), ref) < 1e-3


@pytest.mark.parametrize("hi,wi,ho,wo", [(1024, 1024, 512, 512), (96, 80, 64, 64), (48, 40, 64, 72), (64, 64, 64, 64)])
def test_resize_bilinear_matches_interpolate_forward_and_backward(hi, wi, ho, wo):
    """gd_resize_bilinear / _bwd == F.interpolate(mode="bilinear", align_corners=False) and its autograd
    (stable_diffusion_guidance.py:387-396; 1024^2 -> 512^2 in the shipped config; up- and down-scaling, odd ratios)."""
    import torch.nn.functional as F
    from garmentdreamer_b200 import unet_ops as ops
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, hi, wi, generator=g).cuda()
    dy = torch.randn(2, 3, ho, wo, generator=g).cuda()
    xr = x.clone().requires_grad_(True)
    ref = F.interpolate(xr, (ho, wo), mode="bilinear", align_corners=False)
    ref.backward(dy)
    xo = x.clone().requires_grad_(True)
    y = ops.resize_bilinear(xo, (ho, wo))
    if (hi, wi) == (ho, wo):
        assert y is xo
        return
    y.backward(dy)
    assert float((y - ref).abs().max()) < 1e-5
    assert float((xo.grad - xr.grad).abs().max()) < 1e-4 * float(xr.grad.abs().max())
