"""Generates tests/golden/*.npz by running the UNMODIFIED reference CUDA rasteriser
(oracle/_ref/libgd_ref_raster.so, built by oracle/Makefile from /root/reference) on a B200.

    gpurun -- 'python tests/golden/make_golden.py --out gpurun_out/golden'
    cp gpurun_out/golden/*.npz tests/golden/

Also prints a three-way report (reference vs oracle vs product kernels) for every case.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def report(tag, ref_st, ref_out, ref_g, st, out, g, vis):
    """Prints mismatch counts between the reference (ref_*) and another implementation."""
    def cnt(name, a, b, mask=None):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            print(f"  [{tag}] {name}: SHAPE {a.shape} vs {b.shape}")
            return
        d = a != b
        if mask is not None:
            d = d[mask]
        print(f"  [{tag}] {name}: {int(d.sum())} / {d.size} differ")
    cnt("radii", ref_out["radii"], out["radii"])
    cnt("tiles_touched", ref_st["tiles_touched"], st["tiles_touched"])
    cnt("point_offsets", ref_st["point_offsets"], st["point_offsets"])
    cnt("depth bits", bits(ref_st["depths"]), bits(st["depths"]), vis)
    cnt("means2D bits", bits(ref_st["means2D"]), bits(st["means2D"]), vis)
    cnt("conic_opacity bits", bits(ref_st["conic_opacity"]), bits(st["conic_opacity"]), vis)
    cnt("rgb bits", bits(ref_st["rgb"]), bits(st["rgb"]), vis)
    cnt("cov3D bits", bits(ref_st["cov3D"]), bits(st["cov3D"]), vis)
    print(f"  [{tag}] num_rendered ref {ref_st['num_rendered']} vs {st['num_rendered']}")
    cnt("point_list", ref_st["point_list"], st["point_list"])
    cnt("ranges", ref_st["ranges"], st["ranges"])
    cnt("n_contrib", ref_st["n_contrib"], st["n_contrib"])
    for k in ("color", "depth", "alpha"):
        a, b = np.asarray(ref_out[k]), np.asarray(out[k])
        print(f"  [{tag}] {k}: bit-diff {int((bits(a) != bits(b)).sum())}/{a.size}, max abs {np.abs(a - b).max():.3e}")
    if ref_g is not None and g is not None:
        for k in ("means2D", "conic", "opacity", "colors", "depths", "means3D", "cov3D", "sh", "scales", "rotations"):
            if k not in g or k not in ref_g:
                continue
            a, b = np.asarray(ref_g[k], np.float64).reshape(-1), np.asarray(g[k], np.float64).reshape(-1)
            if a.size == 0:
                continue
            den = np.abs(a).max() + 1e-30
            print(f"  [{tag}] grad {k}: max|d|/max|ref| {np.abs(a - b).max() / den:.3e}  "
                  f"rel-L2 {np.linalg.norm(a - b) / (np.linalg.norm(a) + 1e-30):.3e}")


def np_out(o):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in o.items()}


def ours_state(o):
    from garmentdreamer_b200 import raster
    s = raster.inspect_state(o["state"])
    st = {
        "tiles_touched": s["tiles_touched"][0].numpy().astype(np.uint32),
        "point_offsets": s["point_offsets"][0].numpy().astype(np.uint32),
        "depths": s["depths"][0].numpy(), "means2D": s["means2D"][0].numpy(),
        "conic_opacity": s["conic_opacity"][0].numpy(), "rgb": s["rgb"][0].numpy(),
        "cov3D": s["cov3D"].numpy(), "num_rendered": s["view_base"][1] - s["view_base"][0],
        "point_list": s["point_list"][s["view_base"][0]:s["view_base"][1]].numpy().astype(np.uint32),
        "ranges": s["ranges"][0].numpy().astype(np.uint32),
        "n_contrib": s["n_contrib"][0].numpy().astype(np.uint32),
        "clamped": s["clamped"][0].numpy(),
    }
    out = {"radii": o["radii"][0].cpu().numpy(), "color": o["color"][0].cpu().numpy(),
           "depth": o["depth"][0].cpu().numpy(), "alpha": o["alpha"][0].cpu().numpy()}
    g = None
    if "grads" in o:
        g = {k: v[0].cpu().numpy() for k, v in o["grads"].items()}
    return st, out, g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    ap.add_argument("--big", action="store_true", help="also report on c1 / c2-sized cases")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    names = list(cases.SMALL_CASES) + (["c1"] if args.big else [])
    for name in names:
        c = cases.make_case(name)
        print(f"== case {name}: P={c['P']} {c['W']}x{c['H']} sh_degree={c['sh_degree']}")
        ref_out, ref_st, ref_g = cases.ref_run(c)
        ref_out = np_out(ref_out)
        ref_g = np_out(ref_g) if ref_g is not None else None
        vis = ref_out["radii"] > 0
        print(f"  reference: num_rendered={ref_st['num_rendered']} visible={int(vis.sum())}")
        if name in cases.SMALL_CASES:
            save = {"radii": ref_out["radii"], "color": ref_out["color"], "depth": ref_out["depth"],
                    "alpha": ref_out["alpha"]}
            for k in ("tiles_touched", "point_offsets", "depths", "means2D", "conic_opacity", "rgb",
                      "cov3D", "clamped", "point_list", "ranges", "n_contrib", "keys_sorted"):
                save["st_" + k] = ref_st[k]
            save["num_rendered"] = np.int64(ref_st["num_rendered"])
            if ref_g is not None:
                for k, v in ref_g.items():
                    save["g_" + k] = v
            for k in ("means3D", "opacities", "viewmatrix", "projmatrix", "campos"):
                save["in_" + k] = c[k]
            np.savez_compressed(os.path.join(args.out, name + ".npz"), **save)
        t0 = time.time()
        ost, og = cases.oracle_run(c)
        print(f"  oracle (CPU) took {time.time() - t0:.2f}s")
        oout = {"radii": ost["radii"], "color": ost["color"], "depth": ost["depth"], "alpha": ost["alpha"]}
        report("oracle", ref_st, ref_out, ref_g, ost, oout, og, vis)
        try:
            o = cases.ours_run(c)
            torch.cuda.synchronize()
            st, out, g = ours_state(o)
            report("ours", ref_st, ref_out, ref_g, st, out, g, vis)
        except Exception as e:  # keep generating fixtures even if the product path breaks
            import traceback
            traceback.print_exc()
            print(f"  [ours] FAILED: {e}")
    print("golden written to", args.out)


if __name__ == "__main__":
    main()
