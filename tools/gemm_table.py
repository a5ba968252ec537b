"""Per-shape timing of every GEMM launched by one UNet forward (CUDA events, eager mode)."""
import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_ref
from garmentdreamer_b200 import unet_ops as ops
from garmentdreamer_b200.unet import UNetB200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sd16 = {k: v.cuda().half() for k, v in unet_ref.make_state_dict(0).items()}
net = UNetB200(sd16, "cuda", use_cuda_graph=False)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 4, 64, 64, generator=g).cuda().half()
t = torch.randint(20, 981, (B,), generator=g).cuda().half()
ctx = torch.randn(B, 77, 1024, generator=g).cuda().half()
recs = []
orig = ops._gemm
def timed(a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(a); e1.record()
    recs.append(((a.M, a.N, a.K, a.batch, a.ntaps, a.flags), e0, e1))
with torch.no_grad():
    for _ in range(2): net(x, t, encoder_hidden_states=ctx)
    ops._gemm = timed
    net(x, t, encoder_hidden_states=ctx)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for k, e0, e1 in recs:
    agg[k][0] += 1; agg[k][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"total GEMM time {tot:.2f} ms over {len(recs)} launches")
print(f"{'M':>6} {'N':>6} {'K':>6} {'bat':>4} taps flg {'n':>3} {'ms':>8} {'us/call':>8} {'TFLOP/s':>8}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, K, b, taps, fl = k
    fl_ops = 2.0 * M * N * K * b * v[0]
    print(f"{M:6d} {N:6d} {K:6d} {b:4d} {taps:4d} {fl:3d} {v[0]:3d} {v[1]:8.3f} {1e3 * v[1] / v[0]:8.1f} {fl_ops / v[1] / 1e9:8.1f}")
