import sys, torch, torch.nn.functional as F
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from garmentdreamer_b200 import unet_ops as ops
def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g, device="cuda") * scale).half()
rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))
for (N, H, W, Cin, Cout) in [(1, 32, 32, 64, 512), (1, 64, 64, 512, 512), (1, 256, 256, 128, 128), (1, 128, 128, 256, 256), (1, 128, 128, 256, 128), (1, 32, 32, 512, 512), (4, 64, 64, 64, 512), (1, 64, 64, 512, 256)]:
    x = rnd(N, H, W, Cout); gam, bet = rnd(Cout), rnd(Cout)
    dz_in, w = rnd(N, H, W, Cin, seed=5), rnd(Cout, 3, 3, Cin, scale=(9 * Cin) ** -0.5, seed=6)
    _, stats = ops.groupnorm_stats(x, gam, bet, eps=1e-6, silu=True)
    plain = ops.conv3x3(dz_in, w)
    d_plain = ops.groupnorm_bwd(x, plain.clone(), gam, bet, stats, silu=True)
    g = ops.conv3x3(dz_in, w, gn_bwd=(x, stats, gam, bet))
    fused = bool(getattr(g, "_gd_is_g", False))
    d_fused = ops.groupnorm_bwd(x, g, gam, bet, stats, silu=True, out=g)
    print((N, H, W, Cin, Cout), "fused" if fused else "fallback", "rel", rel(d_fused, d_plain))
