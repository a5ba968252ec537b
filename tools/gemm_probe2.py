"""Experiments on one conv shape: epilogue skipped (flag 0x200), ring depth (GD_GEMM_STAGES env), CTA pairs (GD_GEMM_PAIR env)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garmentdreamer_b200 import unet_ops as ops
from tools.gemm_probe import graph_time

def conv(N_, H, W, Ci, Co, flags):
    x = torch.randn(N_, H, W, Ci, device="cuda").half(); w = (torch.randn(Co, 9 * Ci, device="cuda") * (9 * Ci) ** -0.5).half()
    b = torch.randn(Co, device="cuda").half(); out = torch.empty(N_, H, W, Co, device="cuda", dtype=torch.float16)
    n0 = ops.lib().gd_unet_pair_launch_count()
    us = graph_time(lambda: ops.conv3x3(x, w, b, out=out, flags=flags))
    print("pair launches:", ops.lib().gd_unet_pair_launch_count() - n0, end="  ")
    print(f"[pair={os.environ.get('GD_GEMM_PAIR','1')} stages={os.environ.get('GD_GEMM_STAGES','-')}] conv {N_}x{H}x{W} {Ci}->{Co} flags={flags:#x} {us:8.1f} us {2.0 * N_ * H * W * 9 * Ci * Co / us / 1e6:7.1f} TFLOP/s", flush=True)

for fl in (0,):
    conv(4, 512, 512, 128, 128, fl); conv(4, 256, 256, 256, 256, fl); conv(8, 64, 64, 320, 320, fl)
