"""Times the product rasteriser (B views per call) against the unmodified reference CUDA core
(oracle/_ref, one call per view) on the synthetic garment. Development aid; bench.py is the
contract benchmark."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garmentdreamer_b200 import raster  # noqa: E402
from garmentdreamer_b200.synthetic import garment, sample_cameras  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=100000)
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ref", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = {k: v.to(dev) for k, v in garment(a.P, 0).items()}
    cams = sample_cameras(a.B, a.res, a.res)
    views = [raster.View(c.viewmatrix.to(dev), c.projmatrix.to(dev), c.campos.to(dev), c.tanfovx, c.tanfovy)
             for c in cams]
    bg = torch.ones(3, device=dev)
    gen = torch.Generator(device="cpu").manual_seed(7)
    dc = torch.randn(a.B, 3, a.res, a.res, generator=gen).to(dev)
    dd = torch.randn(a.B, 1, a.res, a.res, generator=gen).to(dev)
    da = torch.randn(a.B, 1, a.res, a.res, generator=gen).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kw = dict(shs=g["shs"], scales=g["scales"], rotations=g["rotations"])

    def fwd(sync):
        return raster.forward_views(g["xyz"], g["opacity"], views, a.res, a.res, bg, sync=sync, **kw)

    color, depth, alpha, radii, st = fwd(True)
    print(f"P={a.P} B={a.B} {a.res}^2  num_rendered={st.num_rendered} view_base={st.view_base}")
    tf, tb = [], []
    for it in range(a.iters + 3):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        color, depth, alpha, radii, st = fwd(False)
        e1.record()
        gr = raster.backward_views(st, g["xyz"], radii, alpha, bg, dc, dd, da, sum_views=True, **kw)
        e2.record()
        torch.cuda.synchronize()
        if it >= 3:
            tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
    med = lambda x: sorted(x)[len(x) // 2]
    print(f"ours: fwd {med(tf) * 1e3:.1f} us  bwd {med(tb) * 1e3:.1f} us  (all {a.B} views, median of {a.iters})")
    if a.ref:
        from oracle.ref_cuda import RefRasterizer
        rr = [RefRasterizer() for _ in range(a.B)]
        rf, rb = [], []
        for it in range(a.iters + 3):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            outs = []
            for b, c in enumerate(views):
                outs.append(rr[b].forward(g["xyz"], g["opacity"].reshape(-1), c.viewmatrix, c.projmatrix,
                                          c.campos, a.res, a.res, c.tanfovx, c.tanfovy, bg, **kw))
            e1.record()
            for b, c in enumerate(views):
                rr[b].backward(g["xyz"], outs[b]["radii"], outs[b]["alpha"], c.viewmatrix, c.projmatrix,
                               c.campos, c.tanfovx, c.tanfovy, bg, dc[b], dd[b], da[b], **kw)
            e2.record()
            torch.cuda.synchronize()
            if it >= 3:
                rf.append(e0.elapsed_time(e1)); rb.append(e1.elapsed_time(e2))
        print(f"reference CUDA (per-view calls, includes its host syncs): fwd {med(rf) * 1e3:.1f} us  "
              f"bwd {med(rb) * 1e3:.1f} us  R={[o['R'] for o in outs]}")


if __name__ == "__main__":
    main()
