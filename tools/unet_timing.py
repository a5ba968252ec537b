"""Times the B200 UNet forward (batch 2B) against the PyTorch-eager fp16 restatement. Dev aid."""
import argparse, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_ref
from garmentdreamer_b200.unet import UNetB200

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--torch", action="store_true")
ap.add_argument("--dump-shapes", default="")
a = ap.parse_args()
sd16 = {k: v.cuda().half() for k, v in unet_ref.make_state_dict(0).items()}
net = UNetB200(sd16, "cuda", use_cuda_graph=not a.no_graph)
g = torch.Generator().manual_seed(1)
x = torch.randn(a.batch, 4, 64, 64, generator=g).cuda().half()
t = torch.randint(20, 981, (a.batch,), generator=g).cuda().half()
ctx = torch.randn(a.batch, 77, 1024, generator=g).cuda().half()
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters
if a.dump_shapes:
    import json
    from garmentdreamer_b200 import unet_ops as _ops
    _rec = []
    _orig = _ops._gemm
    def _spy(g):
        _rec.append([g.M, g.N, g.K, g.batch, g.ntaps, g.flags]); _orig(g)
    _ops._gemm = _spy
with torch.no_grad():
    ms = timeit(lambda: net(x, t, encoder_hidden_states=ctx))
    flops = 804.3e9 * a.batch
    print(f"ours: {ms:.2f} ms / forward (batch {a.batch}) = {flops / ms / 1e9:.1f} TFLOP/s")
    if a.torch:
        ms = timeit(lambda: unet_ref.unet_forward(sd16, x, t, ctx))
        print(f"torch eager fp16 (SDPA, cuDNN/cuBLAS): {ms:.2f} ms / forward = {flops / ms / 1e9:.1f} TFLOP/s")

if a.dump_shapes:
    per = len(_rec) // (3 + a.iters)
    json.dump(_rec[-per:], open(a.dump_shapes, "w"))
