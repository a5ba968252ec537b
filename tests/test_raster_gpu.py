"""GPU parity tests: the product CUDA rasteriser, called through the C ABI, against
 (1) golden vectors of the unmodified reference CUDA core (bit-exact forward, 1e-3 rel grads),
 (2) the C oracle on the same seeded inputs,
 (3) size-independent properties at BASELINE.json's full size (c2: 100k Gaussians, 4 x 512^2).
Tolerance for floating point follows north_star: 1e-3 relative; the index path is bit-exact."""
import os

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def relerr(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def _state(o):
    from garmentdreamer_b200 import raster
    return raster.inspect_state(o["state"])


@pytest.mark.parametrize("name", cases.SMALL_CASES)
def test_matches_reference_golden(name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    c = cases.make_case(name)
    o = cases.ours_run(c)
    s = _state(o)
    vis = gold["radii"] > 0
    assert np.array_equal(o["radii"][0].cpu().numpy(), gold["radii"])
    assert np.array_equal(s["tiles_touched"][0].numpy().astype(np.uint32), gold["st_tiles_touched"])
    assert np.array_equal(s["point_offsets"][0].numpy().astype(np.uint32), gold["st_point_offsets"])
    assert s["num_rendered"] == int(gold["num_rendered"])
    assert np.array_equal(s["point_list"].numpy().astype(np.uint32), gold["st_point_list"])
    assert np.array_equal(s["ranges"][0].numpy().astype(np.uint32), gold["st_ranges"])
    assert np.array_equal(s["n_contrib"][0].numpy().astype(np.uint32), gold["st_n_contrib"])
    assert np.array_equal(bits(s["depths"][0].numpy())[vis], bits(gold["st_depths"])[vis])
    assert np.array_equal(bits(s["means2D"][0].numpy())[vis], bits(gold["st_means2D"])[vis])
    assert np.array_equal(bits(s["conic_opacity"][0].numpy())[vis], bits(gold["st_conic_opacity"])[vis])
    if c["colors_precomp"] is None:
        assert np.array_equal(bits(s["rgb"][0].numpy())[vis], bits(gold["st_rgb"])[vis])
    if c["cov3D_precomp"] is None:
        assert np.array_equal(bits(s["cov3D"].numpy())[vis], bits(gold["st_cov3D"])[vis])
    # sorted keys of the reference are (tile << 32 | depth bits); ours are (depth bits << 32 | idx)
    if s["num_rendered"]:
        ours_depth_bits = (s["tile_keys"].numpy().view(np.uint64) >> np.uint64(32)).astype(np.uint32)
        assert np.array_equal(ours_depth_bits, (gold["st_keys_sorted"] & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    for k in ("color", "depth", "alpha"):  # same arithmetic, same MUFU.EX2 -> bit-exact
        assert np.array_equal(bits(o[k][0].cpu().numpy()), bits(gold[k])), k
    if c["P"] and int(gold["num_rendered"]):
        g = o["grads"]
        for k in ("means2D", "conic", "opacity", "colors", "depths", "means3D", "cov3D", "sh",
                  "scales", "rotations"):
            ref = gold["g_" + k]
            if ref.size == 0 or np.abs(ref).max() == 0 or k not in g:
                continue
            if (k in ("sh",) and c["shs"] is None) or (k in ("scales", "rotations") and c["scales"] is None):
                continue
            assert relerr(g[k][0].cpu().numpy().reshape(ref.shape), ref) < 1e-3, k


@pytest.mark.parametrize("name", ["garment_small", "close_big_splats", "sh3"])
def test_matches_oracle_live(name):
    c = cases.make_case(name)
    o = cases.ours_run(c)
    s = _state(o)
    st, g = cases.oracle_run(c)
    assert np.array_equal(o["radii"][0].cpu().numpy(), st["radii"])
    assert np.array_equal(s["point_list"].numpy().astype(np.uint32), st["point_list"])
    assert np.array_equal(s["ranges"][0].numpy().astype(np.uint32), st["ranges"])
    assert (s["n_contrib"][0].numpy().astype(np.uint32) != st["n_contrib"]).mean() <= 1e-3
    for k in ("color", "depth", "alpha"):
        assert np.abs(o[k][0].cpu().numpy() - st[k]).max() < 3e-6
    for k in ("means3D", "opacity", "scales", "rotations", "sh", "means2D"):
        assert relerr(o["grads"][k][0].cpu().numpy().reshape(g[k].shape), g[k]) < 1e-3, k


def test_batched_views_equal_per_view_calls():
    from garmentdreamer_b200 import raster
    from garmentdreamer_b200.synthetic import garment, sample_cameras
    dev = torch.device("cuda:0")
    g = {k: v.to(dev) for k, v in garment(5000, 0).items()}
    cams = sample_cameras(3, 160, 128)
    views = [raster.View(c.viewmatrix.to(dev), c.projmatrix.to(dev), c.campos.to(dev), c.tanfovx, c.tanfovy) for c in cams]
    bg = torch.tensor([1.0, 1.0, 1.0], device=dev)
    kw = dict(shs=g["shs"], scales=g["scales"], rotations=g["rotations"])
    W, H = 128, 160
    gen = torch.Generator().manual_seed(7)
    dc, dd, da = (torch.randn(3, n, H, W, generator=gen).to(dev) for n in (3, 1, 1))
    cb, db, ab, rb, stb = raster.forward_views(g["xyz"], g["opacity"], views, W, H, bg, **kw)
    gb = raster.backward_views(stb, g["xyz"], rb, ab, bg, dc, dd, da, sum_views=False, **kw)
    gs = raster.backward_views(stb, g["xyz"], rb, ab, bg, dc, dd, da, sum_views=True, **kw)
    sb = raster.inspect_state(stb)
    for b, v in enumerate(views):
        c1, d1, a1, r1, st1 = raster.forward_views(g["xyz"], g["opacity"], [v], W, H, bg, **kw)
        assert torch.equal(c1[0], cb[b]) and torch.equal(d1[0], db[b]) and torch.equal(a1[0], ab[b])
        assert torch.equal(r1[0], rb[b])
        s1 = raster.inspect_state(st1)
        lo, hi = sb["view_base"][b], sb["view_base"][b + 1]
        assert torch.equal(s1["point_list"], sb["point_list"][lo:hi])
        rg = sb["ranges"][b].clone()
        rg[rg[:, 1] > 0] -= lo  # global -> per-view offsets
        assert torch.equal(s1["ranges"][0], rg)
        g1 = raster.backward_views(st1, g["xyz"], r1, a1, bg, dc[b:b + 1], dd[b:b + 1], da[b:b + 1], **kw)
        for k in g1:  # deterministic reduction order -> bit-identical
            assert torch.equal(g1[k][0], gb[k][b]), k
    for k in gs:
        assert torch.allclose(gs[k], gb[k].sum(0), rtol=1e-5, atol=1e-6), k


def test_dropin_autograd_module():
    """GaussianRasterizer used exactly as gaussian_renderer/__init__.py:86-94 does."""
    import diff_gaussian_rasterization as dgr
    c = cases.make_case("garment_small")
    t = cases.to_cuda(c)
    settings = dgr.GaussianRasterizationSettings(
        image_height=c["H"], image_width=c["W"], tanfovx=c["tanfovx"], tanfovy=c["tanfovy"], bg=t["bg"],
        scale_modifier=1.0, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"], sh_degree=0,
        campos=t["campos"], prefiltered=False, debug=False)
    leaves = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    img, radii, depth, alpha = dgr.GaussianRasterizer(settings)(
        means3D=leaves["means3D"], means2D=means2D, shs=leaves["shs"], colors_precomp=None,
        opacities=leaves["opacities"], scales=leaves["scales"], rotations=leaves["rotations"],
        cov3D_precomp=None)
    assert img.shape == (3, c["H"], c["W"]) and depth.shape == (1, c["H"], c["W"]) and radii.dtype == torch.int32
    loss = (img * t["dL_dcolor"]).sum() + (depth * t["dL_ddepth"]).sum() + (alpha * t["dL_dalpha"]).sum()
    loss.backward()
    o = cases.ours_run(c)
    assert torch.equal(img, o["color"][0])
    assert torch.equal(leaves["means3D"].grad, o["grads"]["means3D"][0])
    assert torch.equal(leaves["opacities"].grad, o["grads"]["opacity"][0])
    assert torch.equal(leaves["shs"].grad, o["grads"]["sh"][0])
    assert torch.equal(means2D.grad, o["grads"]["means2D"][0])
    vis = dgr.GaussianRasterizer(settings).markVisible(t["means3D"])
    assert vis.dtype == torch.bool and bool(vis[radii > 0].all())


def test_arena_overflow_regrows():
    from garmentdreamer_b200 import raster
    c = cases.make_case("garment_small")
    t = cases.to_cuda(c)
    view = raster.View(t["viewmatrix"], t["projmatrix"], t["campos"], c["tanfovx"], c["tanfovy"])
    kw = dict(shs=t["shs"], scales=t["scales"], rotations=t["rotations"])
    col, _, _, _, st = raster.forward_views(t["means3D"], t["opacities"], [view], c["W"], c["H"], t["bg"], cap=64, sync=False, **kw)
    n, overflow = raster.read_counters(st)
    assert overflow and n > 64 and float(col.min()) == 1.0  # blank (background) image, flag set
    col2, _, _, _, st2 = raster.forward_views(t["means3D"], t["opacities"], [view], c["W"], c["H"], t["bg"], cap=64, sync=True, **kw)
    assert st2.num_rendered == n and st2.cap >= n
    assert torch.equal(col2, cases.ours_run(c, backward=False)["color"])


@pytest.mark.parametrize("P,B,S", [(100000, 4, 512), (50000, 1, 1024), (500000, 8, 1024)], ids=["c2", "c4a", "c5"])
def test_full_size_properties(P, B, S):
    """BASELINE configs at full size (c2: 100k Gaussians, 4 views 512^2; c4a: 50k, one 1024^2 view;
    c5: 500k, 8 views 1024^2) -- size-independent properties that need no oracle."""
    from garmentdreamer_b200 import raster
    from garmentdreamer_b200.synthetic import garment, sample_cameras
    dev = torch.device("cuda:0")
    g = {k: v.to(dev) for k, v in garment(P, 0).items()}
    views = [raster.View(c.viewmatrix.to(dev), c.projmatrix.to(dev), c.campos.to(dev), c.tanfovx, c.tanfovy)
             for c in sample_cameras(B, S, S)]
    bg = torch.ones(3, device=dev)
    kw = dict(shs=g["shs"], scales=g["scales"], rotations=g["rotations"])
    col, dep, alp, rad, st = raster.forward_views(g["xyz"], g["opacity"], views, S, S, bg, **kw)
    s = raster.inspect_state(st)
    R = s["num_rendered"]
    assert R == int(s["tiles_touched"].to(torch.int64).sum())          # checksum of checksums
    assert int(s["point_offsets"].view(-1)[-1]) == R                   # inclusive scan ends at R
    keys = s["tile_keys"].numpy().view(np.uint64)
    rg = s["ranges"].view(-1, 2).numpy().astype(np.int64)
    nz = rg[rg[:, 1] > 0]
    assert int((nz[:, 1] - nz[:, 0]).sum()) == R
    bad = 0
    for a, b in nz[:: max(1, len(nz) // 200)]:                          # sortedness inside tiles
        bad += int((np.diff(keys[a:b].astype(np.float64)) < 0).sum())
        assert (np.diff(keys[a:b]) > 0).all()
    slot = s["instance_slot"].numpy().astype(np.int64)
    assert np.array_equal(np.sort(slot), np.arange(R))                  # a permutation
    assert float(alp.min()) >= 0 and float(alp.max()) <= 1.0 + 1e-5
    assert torch.isfinite(col).all() and int(s["n_contrib"].max()) <= int((nz[:, 1] - nz[:, 0]).max())
    gen = torch.Generator().manual_seed(7)
    dc, dd, da = (torch.randn(B, n, S, S, generator=gen).to(dev) for n in (3, 1, 1))
    g1 = raster.backward_views(st, g["xyz"], rad, alp, bg, dc, dd, da, sum_views=True, **kw)
    col2, dep2, alp2, rad2, st2 = raster.forward_views(g["xyz"], g["opacity"], views, S, S, bg, **kw)
    g2 = raster.backward_views(st2, g["xyz"], rad2, alp2, bg, dc, dd, da, sum_views=True, **kw)
    assert torch.equal(col, col2)
    for k in g1:                                                        # idempotent + deterministic
        assert torch.equal(g1[k], g2[k]), k
        assert torch.isfinite(g1[k]).all(), k
    # linearity of the backward in the upstream gradient
    g3 = raster.backward_views(st2, g["xyz"], rad2, alp2, bg, 2 * dc, 2 * dd, 2 * da, sum_views=True, **kw)
    assert torch.allclose(g3["means3D"], 2 * g1["means3D"], rtol=1e-4, atol=1e-6)
    # invisible Gaussians get exact zeros
    inv = (rad <= 0).all(0)
    if bool(inv.any()):
        assert float(g1["means3D"][inv].abs().max()) == 0.0
