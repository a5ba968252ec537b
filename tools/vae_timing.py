"""Times the B200 VAE encoder forward + input-gradient backward (B images 512^2) against the
PyTorch-eager fp16 restatement under autograd. Dev aid; --ops prints per-op CUDA-event times."""
import argparse, collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vae_ref
from garmentdreamer_b200 import unet_ops as ops
from garmentdreamer_b200.vae import VAEEncoderB200
from garmentdreamer_b200.sds_step import VAE_FWD_FLOPS_PER_IMAGE, VAE_BWD_FLOPS_PER_IMAGE

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--torch", action="store_true")
ap.add_argument("--ops", action="store_true")
a = ap.parse_args()
sd = {k: v.cuda() for k, v in vae_ref.make_state_dict(0).items()}
enc = VAEEncoderB200(sd, "cuda")
g = torch.Generator().manual_seed(2)
x = torch.rand(a.batch, 3, a.res, a.res, generator=g).cuda()
n = torch.randn(a.batch, 4, a.res // 8, a.res // 8, generator=g).cuda()
gl = torch.randn(a.batch, 4, a.res // 8, a.res // 8, generator=g).cuda()
scale = (a.res / 512.0) ** 2

def run():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); enc.encode(x, n); e[1].record(); enc.backward(gl); e[2].record()
    return e
for _ in range(2): run()
torch.cuda.synchronize()
evs = [run() for _ in range(a.iters)]
torch.cuda.synchronize()
f = sum(e[0].elapsed_time(e[1]) for e in evs) / a.iters
b = sum(e[1].elapsed_time(e[2]) for e in evs) / a.iters
print(f"ours: encode {f:.2f} ms ({VAE_FWD_FLOPS_PER_IMAGE * scale * a.batch / f / 1e9:.0f} TFLOP/s)  backward {b:.2f} ms "
      f"({VAE_BWD_FLOPS_PER_IMAGE * scale * a.batch / b / 1e9:.0f} TFLOP/s)  total {f + b:.2f} ms  (batch {a.batch}, {a.res}^2)")
if a.ops:
    recs = []
    L = ops.lib()
    names = ["gd_unet_gemm", "gd_unet_groupnorm_stats", "gd_unet_groupnorm_bwd", "gd_unet_softmax", "gd_unet_softmax_bwd", "gd_unet_transpose",
             "gd_unet_space_to_depth", "gd_unet_depth_to_space", "gd_unet_conv_in", "gd_vae_prep", "gd_vae_sample", "gd_vae_sample_bwd", "gd_vae_dimg"]
    class Spy:
        def __init__(self, name, fn): self.name, self.fn = name, fn
        def __call__(self, *args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = self.fn(*args); e1.record()
            key = self.name
            if self.name == "gd_unet_gemm":
                g_ = args[0]._obj
                key = f"gemm M={g_.M} N={g_.N} K={g_.K} b={g_.batch} taps={g_.ntaps}"
            recs.append((key, e0, e1)); return r
    class LibSpy:
        def __getattr__(self, k):
            fn = getattr(L, k)
            return Spy(k, fn) if k in names else fn
    spy = LibSpy()
    ops.lib = lambda: spy
    run(); torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, e0, e1 in recs:
        agg[k][0] += 1; agg[k][1] += e0.elapsed_time(e1)
    tot = sum(v[1] for v in agg.values())
    print(f"per-op (eager, event-bracketed; includes launch gaps): total {tot:.2f} ms over {len(recs)} calls")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {v[1]:8.3f} ms  n={v[0]:3d}  {1e3 * v[1] / v[0]:8.1f} us/call  {k}")
    ops.lib = lambda: L
if a.torch:
    sd16 = {k: v.half() for k, v in sd.items()}
    def trun():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); vae_ref.encode_with_grad(sd16, x, n, gl); e[1].record()
        return e
    for _ in range(2): trun()
    torch.cuda.synchronize()
    evs = [trun() for _ in range(a.iters)]
    torch.cuda.synchronize()
    t = sum(e[0].elapsed_time(e[1]) for e in evs) / a.iters
    print(f"torch eager fp16 autograd (cuDNN/cuBLAS/SDPA): encode+backward {t:.2f} ms")
