"""Torch-facing wrappers of the UNet operator C ABI (include/gd_unet.h). fp16, channels-last.

Every dense contraction goes through ONE kernel (gd_unet_gemm: tcgen05.mma + TMEM + TMA); the
wrappers only describe the operand tensors (dims / strides / TMA boxes / conv taps).
"""
import ctypes
import os

import torch

from . import _lib

EPI_SILU, EPI_TRANSPOSED, EPI_GEGLU = 1, 2, 4


class GdGemmArgs(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("batch", ctypes.c_int), ("heads", ctypes.c_int),
        ("A", ctypes.c_void_p),
        ("a_dim", ctypes.c_int * 4), ("a_stride", ctypes.c_longlong * 3), ("a_box", ctypes.c_int * 4),
        ("ntaps", ctypes.c_int), ("Ck", ctypes.c_int),
        ("tap_dx", ctypes.c_int * 9), ("tap_dy", ctypes.c_int * 9), ("tap_c", ctypes.c_int * 9),
        ("rows_per_image", ctypes.c_int), ("img_w", ctypes.c_int), ("img_h", ctypes.c_int),
        ("a_head_k", ctypes.c_int), ("a_zflat", ctypes.c_int),
        ("B", ctypes.c_void_p),
        ("b_dim", ctypes.c_int * 3), ("b_stride", ctypes.c_longlong * 2),
        ("b_head_k", ctypes.c_int), ("b_head_n", ctypes.c_int),
        ("C", ctypes.c_void_p),
        ("ldc", ctypes.c_longlong), ("c_batch_stride", ctypes.c_longlong), ("c_head_stride", ctypes.c_longlong),
        ("bias", ctypes.c_void_p), ("row_bias", ctypes.c_void_p), ("residual", ctypes.c_void_p),
        ("alpha", ctypes.c_float), ("flags", ctypes.c_uint), ("block_n", ctypes.c_int),
        ("row_bias_ld", ctypes.c_longlong),
        ("colstats", ctypes.c_void_p),
        ("gn_coef", ctypes.c_void_p),
        ("c_up2_w", ctypes.c_int),
        ("a_yscale", ctypes.c_int),
    ]


_unet = None


def lib():
    """libgd_unet.so; raises if it has not been built (no fallback path exists)."""
    global _unet
    if _unet is None:
        path = os.path.join(_lib.LIB_DIR, "libgd_unet.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C garmentdreamer_b200/csrc unet`. "
                               "There is no CPU or library fallback.")
        L = ctypes.CDLL(path)
        L.gd_unet_last_error.restype = ctypes.c_char_p
        L.gd_unet_version.restype = ctypes.c_char_p
        L.gd_unet_launch_count.restype = ctypes.c_uint64
        L.gd_unet_pair_launch_count.restype = ctypes.c_uint64
        L.gd_unet_gemm.argtypes = [ctypes.POINTER(GdGemmArgs), ctypes.c_void_p]
        vp, i, f, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
        L.gd_unet_flash_attn.argtypes = [vp, vp, vp, vp, i, i, i, i, ll, ll, ll, ll, f, vp]
        L.gd_unet_flash_attn_ex.argtypes = [vp, vp, vp, vp, i, i, i, i, ll, ll, ll, ll, ll, f, vp]
        L.gd_unet_flash_attn_ex.restype = ctypes.c_int
        L.gd_unet_groupnorm.argtypes = [vp, vp, vp, vp, i, i, i, i, f, i, vp]
        L.gd_unet_layernorm.argtypes = [vp, vp, vp, vp, i, i, f, vp]
        L.gd_unet_softmax.argtypes = [vp, ll, i, ll, vp]
        L.gd_unet_geglu.argtypes = [vp, vp, ll, i, vp]
        L.gd_unet_add.argtypes = [vp, vp, vp, ll, vp]
        L.gd_unet_upsample2x.argtypes = [vp, vp, i, i, i, i, vp]
        L.gd_unet_space_to_depth.argtypes = [vp, vp, i, i, i, i, vp]
        L.gd_unet_concat.argtypes = [vp, vp, vp, ll, i, i, vp]
        L.gd_unet_small_linear.argtypes = [vp, vp, vp, vp, i, i, i, i, i, vp]
        L.gd_unet_timestep_embedding.argtypes = [vp, vp, i, i, vp]
        L.gd_unet_conv_in.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
        L.gd_unet_conv_out.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
        L.gd_unet_add_noise.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, vp]
        L.gd_unet_sds_grad.argtypes = [vp, vp, vp, f, vp, vp, i, i, vp]
        L.gd_unet_pool_latents.argtypes = [vp, vp, vp, i, i, i, vp]
        L.gd_unet_pool_latents_bwd.argtypes = [vp, vp, vp, i, i, i, f, f, vp]
        L.gd_unet_groupnorm_stats.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, f, i, vp]
        L.gd_unet_groupnorm_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp]
        L.gd_unet_softmax_bwd.argtypes = [vp, vp, ll, i, ll, vp]
        L.gd_unet_transpose.argtypes = [vp, vp, i, i, i, vp]
        L.gd_unet_depth_to_space.argtypes = [vp, vp, i, i, i, i, vp]
        L.gd_vae_prep.argtypes = [vp, vp, i, i, i, f, f, vp]
        L.gd_vae_sample.argtypes = [vp, vp, vp, i, i, f, vp]
        L.gd_vae_sample_bwd.argtypes = [vp, vp, vp, vp, i, i, i, f, f, f, vp]
        L.gd_vae_dimg.argtypes = [vp, vp, i, i, i, i, f, vp]
        for name in ("groupnorm_stats", "groupnorm_bwd", "softmax_bwd", "transpose", "depth_to_space"):
            getattr(L, "gd_unet_" + name).restype = ctypes.c_int
        L.gd_vae_im2col.argtypes = [vp, vp, i, i, i, f, f, vp]
        L.gd_vae_dimg_gather.argtypes = [vp, vp, i, i, i, f, vp]
        L.gd_vae_dimg_gather_dyn.argtypes = [vp, vp, i, i, i, f, vp, vp]
        L.gd_vae_grad_scale.argtypes = [vp, ll, f, f, f, vp, vp, vp]
        L.gd_vae_sample_bwd_dyn.argtypes = [vp, vp, vp, vp, i, i, i, f, f, f, vp, vp]
        for name in ("dimg_gather_dyn", "grad_scale", "sample_bwd_dyn"):
            getattr(L, "gd_vae_" + name).restype = ctypes.c_int
        for name in ("prep", "sample", "sample_bwd", "dimg", "im2col", "dimg_gather"):
            getattr(L, "gd_vae_" + name).restype = ctypes.c_int
        for name in ("gemm", "flash_attn", "groupnorm", "layernorm", "softmax", "geglu", "add", "upsample2x", "space_to_depth",
                     "concat", "small_linear", "timestep_embedding", "conv_in", "conv_out", "add_noise", "sds_grad",
                     "pool_latents", "pool_latents_bwd"):
            getattr(L, "gd_unet_" + name).restype = ctypes.c_int
        L.gd_unet_init.restype = ctypes.c_int
        L.gd_unet_groupnorm_colstats.argtypes = [vp, vp, vp, vp, vp, vp, i, vp, i, i, i, i, i, f, i, vp]
        L.gd_unet_groupnorm_colstats.restype = ctypes.c_int
        L.gd_unet_im2col4.argtypes = [vp, vp, i, i, i, vp]
        L.gd_unet_unpack4_nchw.argtypes = [vp, vp, i, ctypes.c_longlong, i, vp]
        L.gd_unet_im2col4.restype = L.gd_unet_unpack4_nchw.restype = ctypes.c_int
        L.gd_unet_gn_bwd_coef.argtypes = [vp, vp, vp, vp, i, i, i, vp]
        L.gd_unet_groupnorm_bwd_g.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, vp]
        L.gd_unet_gn_bwd_coef.restype = L.gd_unet_groupnorm_bwd_g.restype = ctypes.c_int
        L.gd_resize_bilinear.argtypes = [vp, vp, i, i, i, i, i, vp]
        L.gd_resize_bilinear_bwd.argtypes = [vp, vp, i, i, i, i, i, vp]
        L.gd_resize_bilinear.restype = L.gd_resize_bilinear_bwd.restype = ctypes.c_int
        _unet = L
    return _unet


_inited = set()


def init_device(device=None):
    """Allocate the per-device scratch of libgd_unet.so for `device` (before any CUDA-graph capture)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":   # host-side weight re-layout checks build the wrappers on the CPU; nothing to allocate
        return
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx in _inited:
        return
    with torch.cuda.device(idx):
        _chk(lib().gd_unet_init(), "gd_unet_init")
    _inited.add(idx)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().gd_unet_last_error().decode()}")


def _p(t):
    return None if t is None else t.data_ptr()


def _h(t):
    assert t.dtype == torch.float16 and t.is_cuda and t.is_contiguous(), (t.dtype, t.device, t.is_contiguous())
    return t


def _gemm(a: GdGemmArgs, out=None, want_stats=False):
    """Launches the GEMM. want_stats: ask the epilogue for the GroupNorm column statistics of the output
    ([ceil(M/32), 2, N] fp32) and hang them on `out` (``out._gd_colstats = (stats, N)``) for the consuming
    groupnorm(); silently absent when the kernel variant cannot produce them (split-K, ragged N)."""
    stats = None
    if want_stats and _FUSE_GN_STATS:
        stats = torch.empty(((a.M + 31) // 32, 2, a.N), dtype=torch.float32, device=out.device)
        a.colstats = stats.data_ptr()
    rc = lib().gd_unet_gemm(ctypes.byref(a), _stream())
    if rc not in (0, 1):
        _chk(rc, "gd_unet_gemm")
    if out is not None:
        out._gd_colstats = (stats, a.N) if (stats is not None and rc == 0) else None
        out._gd_is_g = bool(a.gn_coef) and rc == 0   # the output already carries silu'(GN(x)) (GroupNorm-backward producer)


_FUSE_GN_STATS = os.environ.get("GD_FUSE_GN_STATS", "1") != "0"


def carry_stats(src, dst):
    """Views made of a GEMM output (reshape to NHWC, ...) keep its column statistics."""
    dst._gd_colstats = getattr(src, "_gd_colstats", None)
    return dst


def _colstats_of(x):
    """(statsA, Ca, statsB, Cb) if the tensor's producer(s) left column statistics, else None."""
    cs = getattr(x, "_gd_colstats", None)
    if cs is None:
        return None
    if len(cs) == 2:
        return cs[0], cs[1], None, 0
    return cs


def linear(x, w, bias=None, *, residual=None, out=None, flags=0, alpha=1.0, block_n=0, want_stats=False):
    """y[M,N] = x[M,K] @ w[N,K]^T (+bias) (+residual); GEGLU flag halves N."""
    _h(x); _h(w)
    M, K = x.shape[-2] * (x.numel() // (x.shape[-1] * x.shape[-2])), x.shape[-1]
    N = w.shape[0]
    n_out = N // 2 if flags & EPI_GEGLU else N
    if out is None:
        out = torch.empty(x.shape[:-1] + (n_out,), dtype=torch.float16, device=x.device)
    a = GdGemmArgs()
    a.M, a.N, a.K, a.batch, a.heads = M, N, K, 1, 1
    a.A = x.data_ptr()
    a.a_dim[:] = [K, M, 1, 1]
    a.a_stride[:] = [K * 2, M * K * 2, M * K * 2]
    a.a_box[:] = [64, 128, 1, 1]
    a.B = w.data_ptr()
    a.b_dim[:] = [K, N, 1]
    a.b_stride[:] = [K * 2, N * K * 2]
    a.C, a.ldc = out.data_ptr(), (out.stride(-2) if out.dim() >= 2 else n_out)
    a.bias, a.residual = _p(bias), _p(residual)
    a.alpha, a.flags, a.block_n = alpha, flags, block_n
    _gemm(a, out, want_stats)
    return out


def bmm_nt(x, y, *, alpha=1.0, out=None):
    """Batched C[b] = alpha * x[b] @ y[b]^T ; x [B,M,K], y [B,N,K] (K contiguous) -> [B,M,N]."""
    _h(x); _h(y)
    B, M, K = x.shape
    N = y.shape[1]
    if out is None:
        out = torch.empty((B, M, N), dtype=torch.float16, device=x.device)
    a = GdGemmArgs()
    a.M, a.N, a.K, a.batch, a.heads = M, N, K, B, 1
    a.A = x.data_ptr()
    a.a_dim[:] = [K, M, B, 1]
    a.a_stride[:] = [K * 2, M * K * 2, B * M * K * 2]
    a.a_box[:] = [64, 128, 1, 1]
    a.B = y.data_ptr()
    a.b_dim[:] = [K, N, B]
    a.b_stride[:] = [K * 2, N * K * 2]
    a.C, a.ldc = out.data_ptr(), N
    a.c_batch_stride = M * N
    a.alpha = alpha
    _gemm(a)
    return out


def _conv_box(H, W, N):
    if W >= 128:   # tile = 128 pixels of one image row
        if W % 128:
            raise ValueError(f"conv width {W} must be a multiple of 128")
        return 128, 1, 1
    if 128 % W:
        raise ValueError(f"conv width {W} must divide 128")
    rows = 128 // W
    if rows <= H:
        if H % rows:
            raise ValueError("conv height must be a multiple of the tile rows")
        return W, rows, 1
    if rows % H:
        raise ValueError("conv tile must hold whole images")
    return W, H, rows // H


def conv_taps(x, w, taps, Ck, bias=None, *, row_bias=None, residual=None, out=None, flags=0, want_stats=False, gn_bwd=None,
              up2=None):
    """Implicit-GEMM convolution over NHWC x[N,H,W,Cx]: out[n,y,x,:] = sum_t x[n, y+dy_t, x+dx_t,
    c_t : c_t+Ck] @ w[:, t*Ck:(t+1)*Ck]^T (zero outside the image). taps = [(dx, dy, c)], w
    [Cout, len(taps)*Ck]. `out` may be a channel slice of a wider NHWC tensor."""
    _h(x); _h(w)
    N, H, W, Cx = x.shape
    Cout = w.shape[0]
    nt = len(taps)
    if out is None:
        out = torch.empty((N, H, W, Cout), dtype=torch.float16, device=x.device)
    a = GdGemmArgs()
    a.M, a.N, a.K, a.batch, a.heads = N * H * W, Cout, nt * Ck, 1, 1
    a.A = x.data_ptr()
    a.a_dim[:] = [Cx, W, H, N]
    a.a_stride[:] = [Cx * 2, W * Cx * 2, H * W * Cx * 2]
    bw, bh, bn = _conv_box(H, W, N)
    a.a_box[:] = [64, bw, bh, bn]
    a.ntaps, a.Ck = nt, Ck
    for t, (dx, dy, c) in enumerate(taps):
        a.tap_dx[t], a.tap_dy[t], a.tap_c[t] = dx, dy, c
    a.rows_per_image, a.img_w, a.img_h = H * W, W, H
    a.B = w.data_ptr()
    a.b_dim[:] = [nt * Ck, Cout, 1]
    a.b_stride[:] = [nt * Ck * 2, Cout * nt * Ck * 2]
    a.C, a.ldc = out.data_ptr(), out.stride(2)
    a.bias, a.row_bias, a.residual = _p(bias), _p(row_bias), _p(residual)
    if row_bias is not None:
        a.row_bias_ld = row_bias.stride(0)
    a.alpha, a.flags = 1.0, flags
    if up2 is not None:
        # `out` is the FULL-resolution tensor [N,2H,2W,Cout]; this GEMM's pixel (n,y,x) lands at (n, 2y+py, 2x+px): the phase
        # GEMMs of a stride-2 data gradient write the upsampled tensor directly
        py, px = up2
        assert out is not None and out.shape == (N, 2 * H, 2 * W, w.shape[0]) and out.is_contiguous()
        a.C, a.ldc = out.data_ptr() + 2 * (py * 2 * W + px) * w.shape[0], w.shape[0]
        a.c_up2_w = W
    if gn_bwd is not None and _FUSE_GN_BWD:
        # this GEMM is the data gradient in front of a GroupNorm+SiLU backward: gn_bwd = (x, stats, gamma, beta) of that GroupNorm
        gx, gstats, ggamma, gbeta = gn_bwd
        Nimg, C = gx.shape[0], gx.shape[-1]
        coef = torch.empty((Nimg, C, 4), dtype=torch.float32, device=gx.device)
        _chk(lib().gd_unet_gn_bwd_coef(gstats.data_ptr(), ggamma.data_ptr(), gbeta.data_ptr(), coef.data_ptr(), Nimg, C, 32, _stream()),
             "gn_bwd_coef")
        a.residual, a.gn_coef = _h(gx).data_ptr(), coef.data_ptr()
        want_stats = True
        if not _FUSE_GN_STATS:   # the column statistics are part of this epilogue
            stats = torch.empty(((a.M + 31) // 32, 2, a.N), dtype=torch.float32, device=out.device)
            a.colstats = stats.data_ptr()
            rc = lib().gd_unet_gemm(ctypes.byref(a), _stream())
            if rc not in (0, 1):
                _chk(rc, "gd_unet_gemm")
            out._gd_colstats = (stats, a.N) if rc == 0 else None
            out._gd_is_g = rc == 0
            return out
    _gemm(a, out, want_stats)
    return out


def conv_stride2_view(x, w, offsets, bias=None, *, out=None, want_stats=False):
    """Stride-2 convolution over NHWC x[N,H,W,C] WITHOUT a space-to-depth copy (needs W/2 >= 128): output pixel (y, x) sums
    taps t over input pixels (2y + oy_t, 2x + ox_t), offsets = [(ox, oy)] in the order of w[Cout, len(offsets)*C]; zero
    outside the image. A is read through a strided view {2C, W/2, H, N} of x (GdGemmArgs.a_yscale = 2)."""
    _h(x); _h(w)
    N, H, W, C = x.shape
    Ho, Wo, Cout, nt = H // 2, W // 2, w.shape[0], len(offsets)
    assert Wo % 128 == 0 and H % 2 == 0 and C % 64 == 0
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), dtype=torch.float16, device=x.device)
    a = GdGemmArgs()
    a.M, a.N, a.K, a.batch, a.heads = N * Ho * Wo, Cout, nt * C, 1, 1
    a.A = x.data_ptr()
    a.a_dim[:] = [2 * C, Wo, H, N]
    a.a_stride[:] = [2 * C * 2, W * C * 2, H * W * C * 2]
    a.a_box[:] = [64, 128, 1, 1]
    a.ntaps, a.Ck = nt, C
    for t, (ox, oy) in enumerate(offsets):
        a.tap_dx[t], a.tap_dy[t], a.tap_c[t] = ox >> 1, oy, (ox & 1) * C
    a.rows_per_image, a.img_w, a.img_h = Ho * Wo, Wo, Ho
    a.B = w.data_ptr()
    a.b_dim[:] = [nt * C, Cout, 1]
    a.b_stride[:] = [nt * C * 2, Cout * nt * C * 2]
    a.C, a.ldc = out.data_ptr(), Cout
    a.bias = _p(bias)
    a.alpha, a.a_yscale = 1.0, 2
    _gemm(a, out, want_stats)
    return out


_FUSE_GN_BWD = os.environ.get("GD_FUSE_GN_BWD", "1") != "0"
_TAPS_3X3 = [(t % 3 - 1, t // 3 - 1, 0) for t in range(9)]


def conv3x3(x, w, bias=None, *, row_bias=None, residual=None, out=None, flags=0, want_stats=False, gn_bwd=None):
    """3x3, stride 1, pad 1 over NHWC x[N,H,W,Cin]; w[Cout,3,3,Cin]."""
    return conv_taps(x, w, _TAPS_3X3, x.shape[-1], bias, row_bias=row_bias, residual=residual, out=out, flags=flags,
                     want_stats=want_stats, gn_bwd=gn_bwd)


def conv3x3_stride2(x, w, bias=None, *, s2d=None, out=None, want_stats=False):
    """3x3, stride 2, pad 1 (diffusers Downsample2D): space-to-depth, then 9 taps with shifts in
    {-1,0} over the 4 phase images (zero fill at the top/left border = the padding)."""
    _h(x); _h(w)
    N, H, W, Cin = x.shape
    Cout = w.shape[0]
    Ho, Wo = H // 2, W // 2
    if s2d is None:
        s2d = torch.empty((N, Ho, Wo, 4 * Cin), dtype=torch.float16, device=x.device)
    _chk(lib().gd_unet_space_to_depth(x.data_ptr(), s2d.data_ptr(), N, H, W, Cin, _stream()), "space_to_depth")
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), dtype=torch.float16, device=x.device)
    a = GdGemmArgs()
    a.M, a.N, a.K, a.batch, a.heads = N * Ho * Wo, Cout, 9 * Cin, 1, 1
    a.A = s2d.data_ptr()
    a.a_dim[:] = [4 * Cin, Wo, Ho, N]
    a.a_stride[:] = [4 * Cin * 2, Wo * 4 * Cin * 2, Ho * Wo * 4 * Cin * 2]
    bw, bh, bn = _conv_box(Ho, Wo, N)
    a.a_box[:] = [64, bw, bh, bn]
    a.ntaps, a.Ck = 9, Cin
    for t in range(9):
        ky, kx = t // 3, t % 3
        py, dy = (1, -1) if ky == 0 else ((0, 0) if ky == 1 else (1, 0))
        px, dx = (1, -1) if kx == 0 else ((0, 0) if kx == 1 else (1, 0))
        a.tap_dx[t], a.tap_dy[t], a.tap_c[t] = dx, dy, (py * 2 + px) * Cin
    a.rows_per_image, a.img_w, a.img_h = Ho * Wo, Wo, Ho
    a.B = w.data_ptr()
    a.b_dim[:] = [9 * Cin, Cout, 1]
    a.b_stride[:] = [9 * Cin * 2, Cout * 9 * Cin * 2]
    a.C, a.ldc = out.data_ptr(), Cout
    a.bias = _p(bias)
    a.alpha = 1.0
    _gemm(a, out, want_stats)
    return out


def attn_scores(q, k, heads, scale, out=None):
    """S[b,h] = scale * Q[b,:,h] K[b,:,h]^T ; q [B,Tq,C], k [B,Tk,C] -> S [B*heads,Tq,Tk] fp16."""
    _h(q); _h(k)
    B, Tq, C = q.shape
    Tk = k.shape[1]
    d = C // heads
    assert d == 64, "head_dim 64 (SD-2.1)"
    ld = (Tk + 7) // 8 * 8
    if out is None:
        out = torch.empty((B * heads, Tq, ld), dtype=torch.float16, device=q.device)
    a = GdGemmArgs()
    a.M, a.N, a.K, a.batch, a.heads = Tq, Tk, d, B * heads, heads
    a.A = q.data_ptr()
    a.a_dim[:] = [C, Tq, B, 1]
    a.a_stride[:] = [C * 2, Tq * C * 2, B * Tq * C * 2]
    a.a_box[:] = [64, 128, 1, 1]
    a.a_head_k, a.a_zflat = d, 0
    a.B = k.data_ptr()
    a.b_dim[:] = [C, Tk, B]
    a.b_stride[:] = [C * 2, Tk * C * 2]
    a.b_head_k, a.b_head_n = d, 0
    a.C, a.ldc = out.data_ptr(), ld
    a.c_batch_stride, a.c_head_stride = heads * Tq * ld, Tq * ld
    a.alpha = scale
    _gemm(a)
    return out


def attn_values(p, vt, heads, Tk, out):
    """O[b,:,h] = P[b,h] V[b,:,h]; p [B*heads,Tq,ld] (rows padded to ld), vt = V^T [B,C,ldv]."""
    _h(p); _h(vt)
    BH, Tq, ld = p.shape
    B, C, ldv = vt.shape
    d = C // heads
    Kp = (Tk + 63) // 64 * 64  # K loop in 64-blocks; the tensor maps bound reads to Tk (zero fill)
    a = GdGemmArgs()
    a.M, a.N, a.K, a.batch, a.heads = Tq, d, Kp, BH, heads
    a.A = p.data_ptr()
    a.a_dim[:] = [Tk, Tq, BH, 1]
    a.a_stride[:] = [ld * 2, Tq * ld * 2, BH * Tq * ld * 2]
    a.a_box[:] = [64, 128, 1, 1]
    a.a_head_k, a.a_zflat = 0, 1
    a.B = vt.data_ptr()
    a.b_dim[:] = [Tk, C, B]
    a.b_stride[:] = [ldv * 2, C * ldv * 2]
    a.b_head_k, a.b_head_n = 0, d
    a.C, a.ldc = out.data_ptr(), C
    a.c_batch_stride, a.c_head_stride = Tq * C, d
    a.alpha = 1.0
    a.block_n = 64
    _gemm(a)
    return out


def flash_attention(q, k, vt, heads, Tk, scale=0.125, out=None):
    """Fused softmax(scale q k^T) v; q [B,Tq,C], k [B,Tk,C] (last dim may be a wider row: strides are
    taken from the tensors), vt = V^T [B,C,ldv]; returns [B,Tq,C]."""
    B, Tq, C = q.shape[0], q.shape[1], heads * 64
    if out is None:
        out = torch.empty((B, Tq, C), dtype=torch.float16, device=q.device)
    _chk(lib().gd_unet_flash_attn_ex(q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr(), B, heads, Tq, Tk,
                                     q.stride(1), k.stride(1), vt.stride(1), out.stride(1), vt.stride(0), scale, _stream()),
         "flash_attn")
    return out


def linear_transposed(x, w, Tk_pad, out=None, bias=None):
    """V^T projection: x [B,T,K] @ w[N,K]^T stored transposed per batch -> [B,N,Tk_pad]."""
    _h(x); _h(w)
    B, T, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.zeros((B, N, Tk_pad), dtype=torch.float16, device=x.device)
    a = GdGemmArgs()
    a.M, a.N, a.K, a.batch, a.heads = T, N, K, B, 1
    a.A = x.data_ptr()
    a.a_dim[:] = [K, T, B, 1]
    a.a_stride[:] = [K * 2, T * K * 2, B * T * K * 2]
    a.a_box[:] = [64, 128, 1, 1]
    a.B = w.data_ptr()
    a.b_dim[:] = [K, N, 1]
    a.b_stride[:] = [K * 2, N * K * 2]
    a.C, a.ldc = out.data_ptr(), Tk_pad
    a.c_batch_stride = N * Tk_pad
    a.bias = _p(bias)
    a.alpha, a.flags = 1.0, EPI_TRANSPOSED
    _gemm(a)
    return out


def groupnorm(x, gamma, beta, groups=32, eps=1e-5, silu=False, out=None):
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    cs = _colstats_of(x)
    out = torch.empty_like(x) if out is None else out
    out._gd_colstats = None
    if cs is not None and HW % 32 == 0 and cs[1] + cs[3] == C:
        # statistics came out of the producing GEMM epilogue(s): no statistics pass over x
        _chk(lib().gd_unet_groupnorm_colstats(_h(x).data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None,
                                              cs[0].data_ptr(), cs[1], _p(cs[2]), cs[3], N, HW, C, groups, eps, int(silu), _stream()),
             "groupnorm_colstats")
        return out
    _chk(lib().gd_unet_groupnorm(_h(x).data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), N, HW, C, groups,
                                 eps, int(silu), _stream()), "groupnorm")
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    C = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    _chk(lib().gd_unet_layernorm(_h(x).data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                 x.numel() // C, C, eps, _stream()), "layernorm")
    return out


def softmax_(s, cols):
    ld = s.shape[-1]
    _chk(lib().gd_unet_softmax(_h(s).data_ptr(), s.numel() // ld, cols, ld, _stream()), "softmax")
    return s


def geglu(x, out=None):
    H = x.shape[-1] // 2
    out = torch.empty(x.shape[:-1] + (H,), dtype=torch.float16, device=x.device) if out is None else out
    _chk(lib().gd_unet_geglu(_h(x).data_ptr(), out.data_ptr(), x.numel() // (2 * H), H, _stream()), "geglu")
    return out


def add(a, b, out=None):
    out = torch.empty_like(a) if out is None else out
    _chk(lib().gd_unet_add(_h(a).data_ptr(), _h(b).data_ptr(), out.data_ptr(), a.numel(), _stream()), "add")
    return out


def upsample2x(x, out=None):
    N, H, W, C = x.shape
    out = torch.empty((N, 2 * H, 2 * W, C), dtype=torch.float16, device=x.device) if out is None else out
    _chk(lib().gd_unet_upsample2x(_h(x).data_ptr(), out.data_ptr(), N, H, W, C, _stream()), "upsample2x")
    return out


def concat(a, b, out=None):
    Ca, Cb = a.shape[-1], b.shape[-1]
    out = torch.empty(a.shape[:-1] + (Ca + Cb,), dtype=torch.float16, device=a.device) if out is None else out
    sa, sb = getattr(a, "_gd_colstats", None), getattr(b, "_gd_colstats", None)
    out._gd_colstats = (sa[0], Ca, sb[0], Cb) if (sa is not None and sb is not None and len(sa) == 2 and len(sb) == 2) else None
    _chk(lib().gd_unet_concat(_h(a).data_ptr(), _h(b).data_ptr(), out.data_ptr(), a.numel() // Ca, Ca, Cb, _stream()), "concat")
    return out


def small_linear(x, w, bias, silu_in=False, silu_out=False):
    Bm, K = x.shape
    N = w.shape[0]
    out = torch.empty((Bm, N), dtype=torch.float16, device=x.device)
    _chk(lib().gd_unet_small_linear(_h(x).data_ptr(), _h(w).data_ptr(), _p(bias), out.data_ptr(), Bm, K, N,
                                    int(silu_in), int(silu_out), _stream()), "small_linear")
    return out


def timestep_embedding(t_f32, dim):
    Bm = t_f32.shape[0]
    out = torch.empty((Bm, dim), dtype=torch.float16, device=t_f32.device)
    _chk(lib().gd_unet_timestep_embedding(t_f32.data_ptr(), out.data_ptr(), Bm, dim, _stream()), "timestep_embedding")
    return out


def conv_in(x_nchw, w, bias):
    N, _, H, W = x_nchw.shape
    Cout = w.shape[0]
    out = torch.empty((N, H, W, Cout), dtype=torch.float16, device=x_nchw.device)
    _chk(lib().gd_unet_conv_in(_h(x_nchw).data_ptr(), _h(w).data_ptr(), bias.data_ptr(), out.data_ptr(), N, H, W, Cout,
                               _stream()), "conv_in")
    return out


def conv_out(x, w, bias):
    N, H, W, Cin = x.shape
    out = torch.empty((N, 4, H, W), dtype=torch.float32, device=x.device)
    _chk(lib().gd_unet_conv_out(_h(x).data_ptr(), _h(w).data_ptr(), bias.data_ptr(), out.data_ptr(), N, H, W, Cin,
                                _stream()), "conv_out")
    return out


def im2col4(x_nchw):
    """[N,4,H,W] fp16 latents -> [N*H*W, 64] im2col rows of the 3x3 conv_in (columns (ky*3+kx)*4 + c, 36..63 zero)."""
    N, _, H, W = x_nchw.shape
    out = torch.empty((N * H * W, 64), dtype=torch.float16, device=x_nchw.device)
    _chk(lib().gd_unet_im2col4(_h(x_nchw).data_ptr(), out.data_ptr(), N, H, W, _stream()), "im2col4")
    return out


def unpack4_nchw(x):
    """[N,H,W,C>=4] fp16 NHWC -> [N,4,H,W] fp32 (first four channels)."""
    N, H, W, C = x.shape
    out = torch.empty((N, 4, H, W), dtype=torch.float32, device=x.device)
    _chk(lib().gd_unet_unpack4_nchw(_h(x).data_ptr(), out.data_ptr(), N, H * W, C, _stream()), "unpack4_nchw")
    return out


# ---- VAE encoder support (include/gd_unet.h, "VAE" section) ---------------------------------
def groupnorm_stats(x, gamma, beta, groups=32, eps=1e-6, silu=False, out=None, apply=True):
    """GroupNorm(+SiLU) that also returns the (mean, rstd) table [N*groups, 2] fp32 for the backward."""
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    stats = torch.empty((N * groups, 2), dtype=torch.float32, device=x.device)
    if apply and out is None:
        out = torch.empty_like(x)
    cs = _colstats_of(x)
    if apply and cs is not None and HW % 32 == 0 and cs[1] + cs[3] == C:
        out._gd_colstats = None
        _chk(lib().gd_unet_groupnorm_colstats(_h(x).data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(),
                                              cs[0].data_ptr(), cs[1], _p(cs[2]), cs[3], N, HW, C, groups, eps, int(silu), _stream()),
             "groupnorm_colstats")
        return out, stats
    _chk(lib().gd_unet_groupnorm_stats(_h(x).data_ptr(), out.data_ptr() if apply else None, gamma.data_ptr(), beta.data_ptr(),
                                       stats.data_ptr(), N, HW, C, groups, eps, int(silu), _stream()), "groupnorm_stats")
    return out, stats


def groupnorm_bwd(x, dz, gamma, beta, stats, groups=32, silu=False, add=None, out=None):
    """dx of z = act(GN(x)) given dz (+ add)."""
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    out = torch.empty_like(x) if out is None else out
    cs = getattr(dz, "_gd_colstats", None)
    if getattr(dz, "_gd_is_g", False) and cs is not None and silu and HW % 32 == 0:
        # dz came out of a GEMM whose epilogue already applied silu'(GN(x)) and summed g | g*xh per column: no statistics sweep
        _chk(lib().gd_unet_groupnorm_bwd_g(_h(x).data_ptr(), _h(dz).data_ptr(), _p(add), out.data_ptr(), gamma.data_ptr(),
                                           beta.data_ptr(), stats.data_ptr(), cs[0].data_ptr(), N, HW, C, groups, _stream()),
             "groupnorm_bwd_g")
        out._gd_colstats, out._gd_is_g = None, False
        return out
    if getattr(dz, "_gd_is_g", False):
        raise RuntimeError("groupnorm_bwd: the gradient was produced by a GroupNorm-backward GEMM epilogue but cannot be consumed here")
    _chk(lib().gd_unet_groupnorm_bwd(_h(x).data_ptr(), _h(dz).data_ptr(), _p(add), out.data_ptr(), gamma.data_ptr(),
                                     beta.data_ptr(), stats.data_ptr(), N, HW, C, groups, int(silu), _stream()), "groupnorm_bwd")
    out._gd_colstats, out._gd_is_g = None, False
    return out


def softmax_bwd_(p, dp):
    ld = p.shape[-1]
    _chk(lib().gd_unet_softmax_bwd(_h(p).data_ptr(), _h(dp).data_ptr(), p.numel() // ld, ld, ld, _stream()), "softmax_bwd")
    return dp


def transpose(x, out=None):
    """[B,R,C] -> [B,C,R]"""
    B, R, C = x.shape
    out = torch.empty((B, C, R), dtype=torch.float16, device=x.device) if out is None else out
    _chk(lib().gd_unet_transpose(_h(x).data_ptr(), out.data_ptr(), B, R, C, _stream()), "transpose")
    return out


def space_to_depth(x, out=None):
    N, H, W, C = x.shape
    out = torch.empty((N, H // 2, W // 2, 4 * C), dtype=torch.float16, device=x.device) if out is None else out
    _chk(lib().gd_unet_space_to_depth(_h(x).data_ptr(), out.data_ptr(), N, H, W, C, _stream()), "space_to_depth")
    return out


def depth_to_space(x, out=None):
    N, Ho, Wo, C4 = x.shape
    C = C4 // 4
    out = torch.empty((N, 2 * Ho, 2 * Wo, C), dtype=torch.float16, device=x.device) if out is None else out
    _chk(lib().gd_unet_depth_to_space(_h(x).data_ptr(), out.data_ptr(), N, 2 * Ho, 2 * Wo, C, _stream()), "depth_to_space")
    return out


class _ResizeBilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Ho, Wo):
        B, C, Hi, Wi = x.shape
        ctx.shape = (B, C, Hi, Wi, Ho, Wo)
        xin = x.detach().float().contiguous()
        out = torch.empty((B, C, Ho, Wo), dtype=torch.float32, device=x.device)
        _chk(lib().gd_resize_bilinear(xin.data_ptr(), out.data_ptr(), B * C, Hi, Wi, Ho, Wo, _stream()), "resize_bilinear")
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        B, C, Hi, Wi, Ho, Wo = ctx.shape
        gin = g.detach().float().contiguous()
        out = torch.empty((B, C, Hi, Wi), dtype=torch.float32, device=g.device)
        _chk(lib().gd_resize_bilinear_bwd(gin.data_ptr(), out.data_ptr(), B * C, Hi, Wi, Ho, Wo, _stream()), "resize_bilinear_bwd")
        return out.to(g.dtype), None, None


def resize_bilinear(x, size):
    """F.interpolate(x, size, mode="bilinear", align_corners=False) for NCHW CUDA tensors (differentiable);
    the identity when the size already matches."""
    Ho, Wo = int(size[0]), int(size[1])
    if tuple(x.shape[-2:]) == (Ho, Wo):
        return x
    if not x.is_cuda:
        raise RuntimeError("resize_bilinear is CUDA-only (no CPU fallback)")
    return _ResizeBilinear.apply(x, Ho, Wo)
