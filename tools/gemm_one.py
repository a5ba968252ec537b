import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garmentdreamer_b200 import unet_ops as ops
M, N, K = (int(v) for v in sys.argv[1:4])
res = len(sys.argv) > 4 and sys.argv[4] == "1"
x = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
b = torch.randn(N, device="cuda").half(); r = torch.randn(M, N, device="cuda").half() if res else None
out = torch.empty(M, N, device="cuda", dtype=torch.float16)
for _ in range(5): ops.linear(x, w, b, residual=r, out=out)
torch.cuda.synchronize()
