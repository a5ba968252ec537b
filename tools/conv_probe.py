"""Times + checks the 3x3 conv-as-GEMM on the VAE's wide-image shapes (run with GD_GEMM_HALO=0/1 to compare)."""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from garmentdreamer_b200 import unet_ops as ops

dev = torch.device("cuda:0")
shapes = [(4, 512, 512, 128, 128), (4, 256, 256, 128, 256), (4, 256, 256, 256, 256), (4, 128, 128, 256, 512), (4, 128, 128, 512, 512),
          (2, 128, 256, 128, 64)]
if len(sys.argv) > 1 and sys.argv[1] == "narrow":   # images narrower than 128 pixels: GD_GEMM_PATCH=0/1
    shapes = [(8, 64, 64, 320, 320), (8, 64, 64, 640, 320), (8, 32, 32, 640, 640), (8, 32, 32, 1280, 640), (8, 16, 16, 1280, 1280),
              (4, 64, 64, 512, 512), (2, 32, 16, 128, 64), (3, 48, 32, 64, 96)]
for N_, H, W, Ci, Co in shapes:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N_, H, W, Ci, generator=g).to(dev).half()
    w4 = (torch.randn(Co, Ci, 3, 3, generator=g) * (9 * Ci) ** -0.5).to(dev).half()
    w = w4.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    b = torch.randn(Co, generator=g).to(dev).half()
    y = ops.conv3x3(x, w, b)
    ref = F.conv2d(x[-1:].float().permute(0, 3, 1, 2), w4.float(), b.float(), padding=1)
    err = float((y[-1:].float().permute(0, 3, 1, 2) - ref).norm() / ref.norm())
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            ops.conv3x3(x, w, b, out=y)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    fl = 2.0 * N_ * H * W * 9 * Ci * Co
    print(f"conv {Ci:4d}->{Co:4d} @ {N_}x{H}x{W}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  rel err {err:.2e}  halo={os.environ.get('GD_GEMM_HALO', '1')} patch={os.environ.get('GD_GEMM_PATCH', '1')}")
