"""Micro-benchmark of gd_unet_gemm for given shapes / block_n (CUDA events, L2-warm loop)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garmentdreamer_b200 import unet_ops as ops

def bench(M, N, K, bn=0, residual=False, iters=20):
    x = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda").half()
    r = torch.randn(M, N, device="cuda").half() if residual else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3): ops.linear(x, w, b, residual=r, out=out, block_n=bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.linear(x, w, b, residual=r, out=out, block_n=bn)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    return us, 2.0 * M * N * K / us / 1e6

for M, N, K in [(32768, 320, 320), (8192, 640, 640), (2048, 1280, 1280), (32768, 320, 1280), (512, 1280, 1280), (32768, 640, 640)]:
    for bn in (0, 64, 96, 128, 160, 256, 320):
        if bn > 256 or (bn and N % bn and bn not in (64, 128)): continue
        for res in (False, True):
            us, tf = bench(M, N, K, bn, res)
            print(f"M={M:6d} N={N:5d} K={K:5d} bn={bn:3d} res={int(res)}  {us:8.1f} us  {tf:7.1f} TFLOP/s")
