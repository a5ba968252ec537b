"""B200-native SD-2.1 VAE encoder: forward AND input-gradient backward (fp16, channels-last).

Replaces ``StableDiffusionGuidance.encode_images`` of the reference
(Garment_3DGS/threestudio/models/guidance/stable_diffusion_guidance.py:160-167)

    imgs = imgs * 2 - 1
    latents = vae.encode(imgs.half()).latent_dist.sample() * vae.config.scaling_factor

together with the autograd backward the SDS loss drives through it (:424-427): the only trainable
quantity upstream is the rendered image, so the backward is data-gradient only (frozen weights,
`p.requires_grad_(False)` at :99-102). The reference executes diffusers 0.19.0 AutoencoderKL
(not in the reference tree; module tree restated in oracle/vae_ref.py).

Every convolution / Linear / attention matmul, forward and backward, is the tcgen05 implicit-GEMM
kernel of libgd_unet.so (dgrad = the same kernel with flipped, transposed weights prepared once);
GroupNorm(+SiLU) forward/backward, softmax backward, transposes and the sampler are fused sweeps
(csrc/unet/gd_vae.cuh). No torch math on the path; Python only orders the calls.
"""
import torch

from . import unet_ops as ops

CH = (128, 256, 512, 512)
SCALING = 0.18215
EPS = 1e-6
GRAD_TARGET = 16.0   # the fp16 backward chain starts with max |d moments| in [8, 16): a power-of-two loss scale chosen
                     # on the device from the incoming gradient (gd_vae_grad_scale), removed again in gd_vae_dimg_gather_dyn

# stride-2 3x3 conv with padding (0,1,0,1) on the space-to-depth tensor: tap (ky,kx) reads phase
# (ky&1, kx&1) at shift (ky>>1, kx>>1)
_DOWN_TAPS = [((kx >> 1), (ky >> 1), ((ky & 1) * 2 + (kx & 1))) for ky in range(3) for kx in range(3)]


_DOWN_VIEW = __import__("os").environ.get("GD_VAE_DOWN_VIEW", "1") != "0"
_DGRAD_DIRECT = __import__("os").environ.get("GD_VAE_DGRAD_DIRECT", "1") != "0"
_GN_BWD_FUSE_MIN_C = int(__import__("os").environ.get("GD_GN_BWD_FUSE_MIN_C", "256"))


class VAEEncoderB200:
    """encode(imgs01, noise) -> latents; backward(grad_latents) -> d imgs01 (both fp32 NCHW)."""

    def __init__(self, state_dict, device="cuda"):
        self.device = torch.device(device)
        ops.lib()  # fail loudly if the CUDA library is missing
        ops.init_device(self.device)
        sd = {k: v.detach().to(self.device, torch.float32) for k, v in state_dict.items()}
        self.w = {}
        h = lambda t: t.to(torch.float16).contiguous()
        for k, v in sd.items():
            if v.dim() == 1:
                self.w[k] = h(v)
        # conv_in as a GEMM over im2col rows: K index (ky*3+kx)*3+c, 27 real columns of 64
        w = sd["encoder.conv_in.weight"]                                      # [128,3,3,3] = [co,c,ky,kx]
        wf = torch.zeros(w.shape[0], 64, device=self.device)
        wf[:, :27] = w.permute(0, 2, 3, 1).reshape(w.shape[0], 27)
        self.w["conv_in.fwd"] = h(wf)
        # data gradient: Z[p, (ky,kx,c)] = dY[p,:] . w[:,c,ky,kx]  (27 real rows of 32), then a 9-tap gather
        wz = torch.zeros(32, w.shape[0], device=self.device)
        wz[:27] = w.permute(2, 3, 1, 0).reshape(27, w.shape[0])
        self.w["conv_in.bwd"] = h(wz)
        for k in [k for k in sd if k.endswith(".weight") and sd[k].dim() == 4 and k != "encoder.conv_in.weight"]:
            base, v = k[:-len(".weight")], sd[k]
            if base in ("encoder.conv_out", "quant_conv"):
                continue
            if v.shape[2] == 1:      # conv_shortcut 1x1 -> Linear (+ its transpose for the backward)
                m = v.reshape(v.shape[0], v.shape[1])
                self.w[base + ".fwd"], self.w[base + ".bwd"] = h(m), h(m.t())
            elif "downsamplers" in base:
                self.w[base + ".fwd"] = h(v.permute(0, 2, 3, 1).reshape(v.shape[0], -1))   # [Cout, (ky,kx,Cin)]
                # dgrad w.r.t. the space-to-depth tensor, one GEMM per phase: taps with that phase
                for ph in range(4):
                    taps = [(ky, kx) for ky in range(3) for kx in range(3) if (ky & 1) * 2 + (kx & 1) == ph]
                    m = torch.stack([v[:, :, ky, kx].t() for ky, kx in taps], 1)            # [Cin, ntaps, Cout]
                    self.w[f"{base}.bwd{ph}"] = h(m.reshape(v.shape[1], -1))
            else:
                self.w[base + ".fwd"] = h(v.permute(0, 2, 3, 1).reshape(v.shape[0], -1))
                self.w[base + ".bwd"] = h(v.flip(2, 3).permute(1, 2, 3, 0).reshape(v.shape[1], -1))
        # conv_out (512 -> 8) and quant_conv (1x1, 8 -> 8) are both linear: folded into one conv
        wq = sd["quant_conv.weight"].reshape(8, 8)
        wo = torch.einsum("ab,bcyx->acyx", wq, sd["encoder.conv_out.weight"])
        self.w["conv_out.fwd"] = h(wo.permute(0, 2, 3, 1).reshape(8, -1))
        self.w["conv_out.bias"] = h(wq @ sd["encoder.conv_out.bias"] + sd["quant_conv.bias"])
        wt = torch.zeros(512, 3, 3, 64, device=self.device)                 # dgrad: 8 real input channels of 64
        wt[..., :8] = wo.flip(2, 3).permute(1, 2, 3, 0)
        self.w["conv_out.bwd"] = h(wt.reshape(512, -1))
        a = "encoder.mid_block.attentions.0"
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            m = sd[f"{a}.{n}.weight"]
            self.w[f"{a}.{n}.fwd"], self.w[f"{a}.{n}.bwd"] = h(m), h(m.t())
        self._saved = None
        # test hook: callable(kind, name, input, output) after every forward block, and
        # callable(kind + "_bwd", name, (block input, upstream gradient), input gradient) after every backward block
        self._trace = None

    # ---- module-like surface ---------------------------------------------------------------
    def eval(self):
        return self

    def parameters(self):
        return iter(self.w.values())

    def requires_grad_(self, flag=False):
        return self

    # ---- blocks ----------------------------------------------------------------------------
    def _resnet_fwd(self, p, x, saved):
        w = self.w
        n1, st1 = ops.groupnorm_stats(x, w[p + ".norm1.weight"], w[p + ".norm1.bias"], eps=EPS, silu=True)
        h1 = ops.conv3x3(n1, w[p + ".conv1.fwd"], w[p + ".conv1.bias"], want_stats=True)
        n2, st2 = ops.groupnorm_stats(h1, w[p + ".norm2.weight"], w[p + ".norm2.bias"], eps=EPS, silu=True, out=n1 if n1.shape == h1.shape else None)
        sc = x
        if p + ".conv_shortcut.fwd" in w:
            N, H, W, C = x.shape
            sc = ops.linear(x.view(N * H * W, C), w[p + ".conv_shortcut.fwd"], w[p + ".conv_shortcut.bias"]).view(N, H, W, -1)
        out = ops.conv3x3(n2, w[p + ".conv2.fwd"], w[p + ".conv2.bias"], residual=sc, want_stats=True)
        saved.append(("resnet", p, x, st1, h1, st2))
        if self._trace is not None:
            self._trace("resnet", p, x, out)
        return out

    def _resnet_bwd(self, rec, dout):
        _, p, x, st1, h1, st2 = rec
        w = self.w
        # >= 256 channels: the data-gradient GEMMs apply silu'(GN(.)) and emit the backward's column sums in their epilogue (no
        # statistics sweep). At 128 channels (512^2) the GEMM is epilogue-bound and the fused form measured SLOWER (+230 us per
        # GEMM against a 100 us sweep), so those keep the two-sweep backward.
        fuse = lambda t, st, n: (t, st, w[p + n + ".weight"], w[p + n + ".bias"]) if t.shape[-1] >= _GN_BWD_FUSE_MIN_C else None
        dn2 = ops.conv3x3(dout, w[p + ".conv2.bwd"], gn_bwd=fuse(h1, st2, ".norm2"))
        dh1 = ops.groupnorm_bwd(h1, dn2, w[p + ".norm2.weight"], w[p + ".norm2.bias"], st2, silu=True, out=dn2)
        dn1 = ops.conv3x3(dh1, w[p + ".conv1.bwd"], gn_bwd=fuse(x, st1, ".norm1"))
        add = dout
        if p + ".conv_shortcut.bwd" in w:
            N, H, W, C = dout.shape
            add = ops.linear(dout.view(N * H * W, C), w[p + ".conv_shortcut.bwd"]).view(N, H, W, -1)
        keep = dout.clone() if self._trace is not None else None
        din = ops.groupnorm_bwd(x, dn1, w[p + ".norm1.weight"], w[p + ".norm1.bias"], st1, silu=True, add=add, out=dn1)
        if self._trace is not None:
            self._trace("resnet_bwd", p, (x, keep), din)
        return din

    def _down_fwd(self, p, x, saved):
        C = x.shape[-1]
        saved.append(("down", p, C))
        if _DOWN_VIEW and (x.shape[2] // 2) % 128 == 0:
            # pad (0,1,0,1), stride 2: output (y, x) reads input (2y + ky, 2x + kx) through a strided view of x -- no space-to-depth copy
            out = ops.conv_stride2_view(x, self.w[p + ".fwd"], [(kx, ky) for ky in range(3) for kx in range(3)], self.w[p + ".bias"],
                                        want_stats=True)
        else:
            s2d = ops.space_to_depth(x)
            taps = [(dx, dy, ph * C) for dx, dy, ph in _DOWN_TAPS]
            out = ops.conv_taps(s2d, self.w[p + ".fwd"], taps, C, self.w[p + ".bias"], want_stats=True)
        if self._trace is not None:
            self._trace("down", p, x, out)
        return out

    def _down_bwd(self, rec, dout):
        _, p, C = rec
        N, Ho, Wo, Cout = dout.shape
        if _DGRAD_DIRECT:
            # the GEMM of phase (py, px) stores its pixel (y, x) at (2y + py, 2x + px) of the full-resolution gradient itself
            # (GdGemmArgs.c_up2_w): no [N,Ho,Wo,4C] intermediate and no depth-to-space pass
            din = torch.empty((N, 2 * Ho, 2 * Wo, C), dtype=torch.float16, device=dout.device)
            for ph in range(4):
                taps = [(-(kx >> 1), -(ky >> 1), 0) for ky in range(3) for kx in range(3) if (ky & 1) * 2 + (kx & 1) == ph]
                ops.conv_taps(dout, self.w[f"{p}.bwd{ph}"], taps, Cout, out=din, up2=(ph >> 1, ph & 1))
        else:
            ds2d = torch.empty((N, Ho, Wo, 4 * C), dtype=torch.float16, device=dout.device)
            for ph in range(4):
                taps = [(-(kx >> 1), -(ky >> 1), 0) for ky in range(3) for kx in range(3) if (ky & 1) * 2 + (kx & 1) == ph]
                ops.conv_taps(dout, self.w[f"{p}.bwd{ph}"], taps, Cout, out=ds2d[..., ph * C:(ph + 1) * C])
            din = ops.depth_to_space(ds2d)
        if self._trace is not None:
            self._trace("down_bwd", p, (None, dout), din)
        return din

    def _attn_fwd(self, p, x, saved):
        w = self.w
        N, H, W, C = x.shape
        T = H * W
        hn, st = ops.groupnorm_stats(x, w[p + ".group_norm.weight"], w[p + ".group_norm.bias"], eps=EPS, silu=False)
        hn = hn.view(N, T, C)
        q = ops.linear(hn, w[p + ".to_q.fwd"], w[p + ".to_q.bias"])
        k = ops.linear(hn, w[p + ".to_k.fwd"], w[p + ".to_k.bias"])
        v = ops.linear(hn, w[p + ".to_v.fwd"], w[p + ".to_v.bias"])
        scale = float(C) ** -0.5
        P = ops.softmax_(ops.bmm_nt(q, k, alpha=scale), T)     # [N,T,T]; one head of 512
        o = ops.bmm_nt(P, ops.transpose(v))                    # P @ v
        out = ops.linear(o, w[p + ".to_out.0.fwd"], w[p + ".to_out.0.bias"], residual=x.view(N, T, C), want_stats=True)
        out = ops.carry_stats(out, out.view(N, H, W, C))
        saved.append(("attn", p, x, st, q, k, v, P, scale))
        if self._trace is not None:
            self._trace("attn", p, x, out)
        return out

    def _attn_bwd(self, rec, dout):
        _, p, x, st, q, k, v, P, scale = rec
        w = self.w
        N, H, W, C = x.shape
        T = H * W
        do = ops.linear(dout.view(N, T, C), w[p + ".to_out.0.bwd"])
        dv = ops.bmm_nt(ops.transpose(P), ops.transpose(do))                   # P^T do
        dS = ops.softmax_bwd_(P, ops.bmm_nt(do, v))                            # dP = do v^T, then softmax bwd in place
        dq = ops.bmm_nt(dS, ops.transpose(k), alpha=scale)                     # dS k
        dk = ops.bmm_nt(ops.transpose(dS), ops.transpose(q), alpha=scale)      # dS^T q
        dhn = ops.linear(dq, w[p + ".to_q.bwd"])
        dhn = ops.linear(dk, w[p + ".to_k.bwd"], residual=dhn, out=dhn)
        dhn = ops.linear(dv, w[p + ".to_v.bwd"], residual=dhn, out=dhn)
        din = ops.groupnorm_bwd(x, dhn.view(N, H, W, C), w[p + ".group_norm.weight"], w[p + ".group_norm.bias"], st,
                                silu=False, add=dout)
        if self._trace is not None:
            self._trace("attn_bwd", p, (x, dout), din)
        return din

    # ---- forward / backward ------------------------------------------------------------------
    def encode(self, imgs, noise, keep_for_backward=True, input_range="01", scaling=SCALING):
        """imgs fp32 [B,3,H,W] (H, W multiples of 128), in [0,1] (``input_range="01"``: the
        `imgs * 2 - 1` of encode_images is fused in) or already in [-1,1] (``"pm1"``); noise fp32
        [B,4,H/8,W/8] replaces the sampler's randn. Returns fp32 [B,4,H/8,W/8] =
        posterior.sample() * scaling."""
        if not imgs.is_cuda:
            raise RuntimeError("garmentdreamer_b200 VAE is CUDA-only (no CPU fallback)")
        L, st = ops.lib(), ops._stream()
        B, _, H, W = imgs.shape
        if H % 128 or W % 128:
            raise ValueError("VAE encoder: image height / width must be multiples of 128")
        imgs = imgs.detach().float().contiguous()
        noise = noise.detach().float().contiguous()
        a, sh = (2.0, -1.0) if input_range == "01" else (1.0, 0.0)
        cols = torch.empty((B * H * W, 64), dtype=torch.float16, device=imgs.device)
        ops._chk(L.gd_vae_im2col(imgs.data_ptr(), cols.data_ptr(), B, H, W, a, sh, st), "vae_im2col")
        saved = []
        x = ops.linear(cols, self.w["conv_in.fwd"], self.w["encoder.conv_in.bias"], want_stats=True)
        x = ops.carry_stats(x, x.view(B, H, W, -1))
        for i in range(4):
            for j in range(2):
                x = self._resnet_fwd(f"encoder.down_blocks.{i}.resnets.{j}", x, saved)
            if i < 3:
                x = self._down_fwd(f"encoder.down_blocks.{i}.downsamplers.0.conv", x, saved)
        x = self._resnet_fwd("encoder.mid_block.resnets.0", x, saved)
        x = self._attn_fwd("encoder.mid_block.attentions.0", x, saved)
        x = self._resnet_fwd("encoder.mid_block.resnets.1", x, saved)
        n, stn = ops.groupnorm_stats(x, self.w["encoder.conv_norm_out.weight"], self.w["encoder.conv_norm_out.bias"], eps=EPS, silu=True)
        mom = ops.conv3x3(n, self.w["conv_out.fwd"], self.w["conv_out.bias"])     # [B,h,w,8] = mean | logvar
        h, w_ = H // 8, W // 8
        lat = torch.empty((B, 4, h, w_), dtype=torch.float32, device=imgs.device)
        ops._chk(L.gd_vae_sample(mom.data_ptr(), noise.data_ptr(), lat.data_ptr(), B, h * w_, float(scaling), st), "vae_sample")
        self._saved = (saved, x, stn, mom, noise, (B, H, W), a, float(scaling)) if keep_for_backward else None
        return lat

    def backward(self, grad_latents, clip=0.0, scale=1.0):
        """d<latents, g>/d imgs for g = nan_to_num(clamp(grad_latents, +-clip)) * scale
        (stable_diffusion_guidance.py:418-427). fp32 [B,3,H,W]."""
        if self._saved is None:
            raise RuntimeError("VAEEncoderB200.backward() needs a preceding encode(keep_for_backward=True)")
        saved, x_out, stn, mom, noise, (B, H, W), a, scaling = self._saved
        self._saved = None
        L, st = ops.lib(), ops._stream()
        h, w_ = H // 8, W // 8
        g = grad_latents.detach().float().contiguous()
        dmom = torch.empty((B, h, w_, 64), dtype=torch.float16, device=g.device)
        n = g.numel()
        if getattr(self, "_dyn", None) is None:
            self._dyn = torch.ones(1, dtype=torch.float32, device=g.device)
        if getattr(self, "_dyn_scratch", None) is None or self._dyn_scratch.numel() < (n + 1023) // 1024:
            self._dyn_scratch = torch.empty((n + 1023) // 1024, dtype=torch.float32, device=g.device)
        ops._chk(L.gd_vae_grad_scale(g.data_ptr(), n, float(clip), scaling * float(scale), GRAD_TARGET, self._dyn_scratch.data_ptr(),
                                     self._dyn.data_ptr(), st), "vae_grad_scale")
        ops._chk(L.gd_vae_sample_bwd_dyn(g.data_ptr(), mom.data_ptr(), noise.data_ptr(), dmom.data_ptr(), B, h * w_, 64,
                                         scaling, float(clip), float(scale), self._dyn.data_ptr(), st), "vae_sample_bwd")
        dn = ops.conv3x3(dmom, self.w["conv_out.bwd"], gn_bwd=(x_out, stn, self.w["encoder.conv_norm_out.weight"], self.w["encoder.conv_norm_out.bias"]))
        d = ops.groupnorm_bwd(x_out, dn, self.w["encoder.conv_norm_out.weight"], self.w["encoder.conv_norm_out.bias"], stn, silu=True, out=dn)
        for rec in reversed(saved):
            if rec[0] == "resnet":
                d = self._resnet_bwd(rec, d)
            elif rec[0] == "attn":
                d = self._attn_bwd(rec, d)
            else:
                d = self._down_bwd(rec, d)
        z = ops.linear(d.view(B * H * W, -1), self.w["conv_in.bwd"])              # [B*H*W,32] per-pixel tap products
        dimg = torch.empty((B, 3, H, W), dtype=torch.float32, device=g.device)
        ops._chk(L.gd_vae_dimg_gather_dyn(z.data_ptr(), dimg.data_ptr(), B, H, W, a, self._dyn.data_ptr(), st), "vae_dimg_gather")
        return dimg

    # ---- autograd + the diffusers surface the reference calls ---------------------------------
    def encode_images(self, imgs01, noise=None, generator=None):
        """Differentiable encode_images (stable_diffusion_guidance.py:160-167): latents carry a
        grad_fn whose backward is the CUDA input-gradient chain above."""
        if noise is None:
            B, _, H, W = imgs01.shape
            noise = torch.randn((B, 4, H // 8, W // 8), device=imgs01.device, dtype=torch.float32, generator=generator)
        return _EncodeFn.apply(imgs01, noise, self, "01", SCALING)


class _EncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, imgs, noise, enc, input_range, scaling):
        ctx.enc = enc
        ctx.in_dtype = imgs.dtype
        lat = enc.encode(imgs, noise, keep_for_backward=True, input_range=input_range, scaling=scaling)
        ctx.saved, enc._saved = enc._saved, None   # the graph node owns its activations (several encodes may be alive)
        return lat

    @staticmethod
    def backward(ctx, grad_latents):
        ctx.enc._saved, ctx.saved = ctx.saved, None
        return ctx.enc.backward(grad_latents).to(ctx.in_dtype), None, None, None, None


class _LatentDist:
    def __init__(self, enc, x_pm1):
        self.enc, self.x = enc, x_pm1

    def sample(self, generator=None):
        B, _, H, W = self.x.shape
        noise = torch.randn((B, 4, H // 8, W // 8), device=self.x.device, dtype=torch.float32, generator=generator)
        return _EncodeFn.apply(self.x.float(), noise, self.enc, "pm1", 1.0).to(self.x.dtype)


class DiffusersVAEView:
    """`pipe.vae` look-alike for the two attributes the reference touches (:165-166):
    ``vae.encode(imgs_pm1).latent_dist.sample()`` and ``vae.config.scaling_factor``."""

    def __init__(self, enc):
        from types import SimpleNamespace
        self.enc = enc
        self.config = SimpleNamespace(scaling_factor=SCALING)

    def encode(self, x_pm1):
        from types import SimpleNamespace
        return SimpleNamespace(latent_dist=_LatentDist(self.enc, x_pm1))

    def eval(self):
        return self

    def parameters(self):
        return self.enc.parameters()
