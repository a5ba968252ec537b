"""GPU: per-block parity of the B200 UNet / VAE kernels.

Every ResnetBlock2D, Transformer2DModel, down/up-sampler and the in/out convolutions of the UNet,
and every encoder block of the VAE (forward AND input-gradient backward), is compared with the
fp32 PyTorch restatement (oracle/unet_ref.py, oracle/vae_ref.py; the weights are the same
fp16-rounded values) fed the SAME fp16 input the product block saw. A kernel bug therefore cannot
hide inside fp16 drift accumulated over the 60-block network: what is left per block is the fp16
rounding of the block's own intermediates, and it is asserted in absolute numbers:

    relative L2 error of every block output  < 1e-3   (north_star's fp16 tolerance)

The per-block numbers are appended to gpurun_out/r02_block_parity.jsonl (copy under profiles/).
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-3


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def _log(rows, tag):
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "r02_block_parity.jsonl"), "a") as f:
            for r in rows:
                f.write(json.dumps({"net": tag, **r}) + "\n")
    except OSError:
        pass


def test_unet_every_block_matches_fp32_restatement_on_same_input():
    from oracle import unet_ref
    from garmentdreamer_b200.unet import HEADS, UNetB200
    sd = unet_ref.make_state_dict(0)
    sd16 = {k: v.cuda().half() for k, v in sd.items()}
    sdr = {k: v.float() for k, v in sd16.items()}          # the SAME (fp16-rounded) weights, fp32 arithmetic
    net = UNetB200(sd16, "cuda", use_cuda_graph=False)
    g = torch.Generator().manual_seed(1)
    B = 2
    x = torch.randn(B, 4, 64, 64, generator=g).cuda()
    t = torch.randint(20, 981, (B,), generator=g).cuda()
    ctx = torch.randn(B, 77, 1024, generator=g).cuda()
    trace = []
    net._trace = lambda kind, name, xin, out: trace.append((kind, name, xin.clone(), out.clone()))
    with torch.no_grad():
        net(x.half(), t.half(), encoder_hidden_states=ctx.half())
        net._trace = None
        temb = unet_ref.timestep_embedding(t.half().float())
        emb = unet_ref._lin(sdr, "time_embedding.linear_2", F.silu(unet_ref._lin(sdr, "time_embedding.linear_1", temb.half().float())))
        ctx32 = ctx.half().float()
        rows = []
        for kind, name, xin, out in trace:
            if kind == "resnet":
                ref = unet_ref.resnet(sdr, name, nchw(xin), emb)
            elif kind == "transformer":
                if name.startswith("mid"):
                    heads = 20
                else:
                    i = int(name.split(".")[1])
                    heads = HEADS[i] if name.startswith("down") else HEADS[::-1][i]
                ref = unet_ref.transformer(sdr, name, nchw(xin), ctx32, heads)
            elif kind == "down":
                ref = F.conv2d(nchw(xin), sdr[name + ".weight"], sdr[name + ".bias"], stride=2, padding=1)
            elif kind == "up":
                ref = F.conv2d(F.interpolate(nchw(xin), scale_factor=2.0, mode="nearest"), sdr[name + ".weight"], sdr[name + ".bias"], padding=1)
            elif kind == "in":
                ref = F.conv2d(xin.float(), sdr["conv_in.weight"], sdr["conv_in.bias"], padding=1)
            else:
                ref = F.conv2d(F.silu(unet_ref._gn(sdr, "conv_norm_out", nchw(xin), 1e-5)), sdr["conv_out.weight"], sdr["conv_out.bias"], padding=1)
            ours = out.float() if kind == "out" else nchw(out)
            rows.append({"kind": kind, "block": name, "shape": list(out.shape), "rel": rel(ours, ref)})
    _log(rows, "unet")
    worst = sorted(rows, key=lambda r: -r["rel"])[:5]
    print("UNet blocks:", len(rows), "worst:", [(r["block"], f"{r['rel']:.2e}") for r in worst])
    assert len([r for r in rows if r["kind"] == "resnet"]) == 22 and len([r for r in rows if r["kind"] == "transformer"]) == 16
    bad = [r for r in rows if not r["rel"] < TOL]
    assert not bad, bad


def test_vae_every_block_forward_and_backward_matches_fp32_restatement_on_same_input():
    from oracle import vae_ref
    from garmentdreamer_b200.vae import VAEEncoderB200
    sd = vae_ref.make_state_dict(0)
    sd16 = {k: v.cuda().half() for k, v in sd.items()}
    sdr = {k: v.float() for k, v in sd16.items()}
    enc = VAEEncoderB200(sdr, "cuda")
    g = torch.Generator().manual_seed(2)
    B, res = 1, 256
    x = torch.rand(B, 3, res, res, generator=g).cuda()
    n = torch.randn(B, 4, res // 8, res // 8, generator=g).cuda()
    gl = torch.randn(B, 4, res // 8, res // 8, generator=g).cuda()
    fwd, bwd = [], []

    def hook(kind, name, a, b):
        if kind.endswith("_bwd"):
            bwd.append((kind[:-4], name, None if a[0] is None else a[0].clone(), a[1].clone(), b.clone()))
        else:
            fwd.append((kind, name, a.clone(), b.clone()))
    enc._trace = hook
    enc.encode(x, n)
    enc.backward(gl)
    enc._trace = None
    inputs = {name: xin for _, name, xin, _ in fwd}

    def block(kind, name, xin):
        if kind == "resnet":
            return vae_ref.resnet(sdr, name, xin)
        if kind == "attn":
            return vae_ref.mid_attention(sdr, name, xin)
        return F.conv2d(F.pad(xin, (0, 1, 0, 1)), sdr[name + ".weight"], sdr[name + ".bias"], stride=2)
    rows = []
    for kind, name, xin, out in fwd:
        with torch.no_grad():
            rows.append({"kind": kind, "dir": "fwd", "block": name, "rel": rel(nchw(out), block(kind, name, nchw(xin)))})
    for kind, name, xin, dout, din in bwd:
        xr = nchw(inputs[name]).requires_grad_(True)
        block(kind, name, xr).backward(nchw(dout))
        rows.append({"kind": kind, "dir": "bwd", "block": name, "rel": rel(nchw(din), xr.grad)})
    _log(rows, "vae")
    worst = sorted(rows, key=lambda r: -r["rel"])[:5]
    print("VAE blocks:", len(rows), "worst:", [(r["block"], r["dir"], f"{r['rel']:.2e}") for r in worst])
    assert len(rows) == 2 * (10 + 3 + 1)
    bad = [r for r in rows if not r["rel"] < TOL]
    assert not bad, bad
