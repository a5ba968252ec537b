"""Times the fused attention kernel on the UNet's self-attention shapes (batch 8). Dev aid.
GD_ATTN_TWO_PASS=1 selects the two-pass kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garmentdreamer_b200 import unet_ops as ops

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for B, T, heads, Tk in [(8, 4096, 5, 4096), (8, 1024, 10, 1024), (8, 256, 20, 256), (8, 64, 20, 64), (8, 4096, 5, 77), (8, 1024, 10, 77), (8, 256, 20, 77)]:
    C = heads * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, T, C, generator=g, device="cuda").half()
    k = torch.randn(B, Tk, C, generator=g, device="cuda").half()
    vt = torch.zeros(B, C, (Tk + 7) // 8 * 8, device="cuda").half()
    vt[:, :, :Tk] = torch.randn(B, C, Tk, generator=g, device="cuda").half()
    o = torch.empty(B, T, C, dtype=torch.float16, device="cuda")
    us = timeit(lambda: ops.flash_attention(q, k, vt, heads, Tk, 0.125, out=o))
    flops = 4.0 * B * heads * T * Tk * 64
    print(f"attn B={B} T={T} Tk={Tk} heads={heads}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s")
