"""Device-time of gd_unet_gemm for representative shapes: 20 back-to-back launches captured in a
CUDA graph (no host overhead, L2-warm like inside the UNet graph)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garmentdreamer_b200 import unet_ops as ops

def graph_time(fn, reps=20, iters=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1e3

def lin(M, N, K, res, bns=(0,), flags=0):
    x = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda").half(); r = torch.randn(M, N, device="cuda").half() if res else None
    n_out = N // 2 if flags & ops.EPI_GEGLU else N
    out = torch.empty(M, n_out, device="cuda", dtype=torch.float16)
    for bn in bns:
        us = graph_time(lambda: ops.linear(x, w, b, residual=r, out=out, flags=flags, block_n=bn))
        print(f"linear M={M:6d} N={N:5d} K={K:5d} res={int(res)} fl={flags} bn={bn:3d} {us:8.1f} us {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)

def conv(N_, H, W, Ci, Co):
    x = torch.randn(N_, H, W, Ci, device="cuda").half(); w = (torch.randn(Co, 9 * Ci, device="cuda") * (9 * Ci) ** -0.5).half()
    b = torch.randn(Co, device="cuda").half(); out = torch.empty(N_, H, W, Co, device="cuda", dtype=torch.float16)
    us = graph_time(lambda: ops.conv3x3(x, w, b, out=out))
    print(f"conv {N_}x{H}x{W} {Ci}->{Co} {us:8.1f} us {2.0 * N_ * H * W * 9 * Ci * Co / us / 1e6:7.1f} TFLOP/s", flush=True)

if __name__ == "__main__":
    lin(32768, 320, 320, True, (0, 64, 128)); lin(32768, 320, 320, False)
    lin(8192, 640, 640, True, (0, 128)); lin(2048, 1280, 1280, True, (0, 128)); lin(32768, 320, 1280, True)
    lin(32768, 640, 320, False); lin(8192, 640, 2560, True); lin(2048, 1280, 5120, True); lin(616, 1280, 1024, False); lin(512, 1280, 1280, True)
    lin(32768, 2560, 320, False, flags=ops.EPI_GEGLU); lin(8192, 5120, 640, False, flags=ops.EPI_GEGLU); lin(2048, 10240, 1280, False, flags=ops.EPI_GEGLU)
    conv(4, 512, 512, 128, 128); conv(4, 256, 256, 256, 256); conv(4, 128, 128, 512, 512); conv(4, 64, 64, 512, 512)
    conv(8, 64, 64, 320, 320); conv(8, 32, 32, 640, 640); conv(8, 16, 16, 1280, 1280); conv(8, 8, 8, 1280, 1280)
    conv(8, 64, 64, 640, 320); conv(8, 32, 32, 1280, 640); conv(8, 16, 16, 2560, 1280)
