"""GPU: fused activation / Adam / densification-statistics kernels against the torch ops the
reference executes (gaussian_model.py:95-115,156-165,415-419): autograd + torch.optim.Adam on the
same inputs. fp32 elementwise math: tolerance 2e-6 relative (libm exp/rsqrt ordering), exact for
the copies and the statistics counters."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _raw(P, seed=0):
    g = torch.Generator().manual_seed(seed)
    return dict(xyz=torch.randn(P, 3, generator=g), f_dc=torch.randn(P, 1, 3, generator=g),
                opacity=torch.randn(P, 1, generator=g) * 2, scaling=torch.randn(P, 3, generator=g) - 3,
                rotation=torch.randn(P, 4, generator=g))


@pytest.mark.parametrize("P", [1, 1000, 100003])
def test_activate_matches_torch(P):
    from garmentdreamer_b200.gaussians import GaussianParams
    r = {k: v.cuda() for k, v in _raw(P).items()}
    gp = GaussianParams(**r)
    xyz, dc, op, sc, rot = GaussianParams.unpack(gp.activated(), P)
    assert torch.equal(xyz, r["xyz"]) and torch.equal(dc, r["f_dc"])
    torch.testing.assert_close(op, torch.sigmoid(r["opacity"]), rtol=2e-6, atol=1e-7)
    torch.testing.assert_close(sc, torch.exp(r["scaling"]), rtol=2e-6, atol=0)
    torch.testing.assert_close(rot, torch.nn.functional.normalize(r["rotation"]), rtol=2e-6, atol=1e-7)


def test_adam_steps_match_torch_autograd():
    from garmentdreamer_b200.gaussians import GaussianParams, OptimizationParams
    P = 5000
    r = {k: v.cuda() for k, v in _raw(P, 3).items()}
    gp = GaussianParams(**r, spatial_lr_scale=1.75)
    gp.training_setup(OptimizationParams())
    ref = {k: torch.nn.Parameter(v.clone()) for k, v in r.items()}
    a = OptimizationParams()
    opt = torch.optim.Adam([
        {"params": [ref["xyz"]], "lr": a.position_lr_init * 1.75, "name": "xyz"},
        {"params": [ref["f_dc"]], "lr": a.feature_lr, "name": "f_dc"},
        {"params": [ref["opacity"]], "lr": a.opacity_lr, "name": "opacity"},
        {"params": [ref["scaling"]], "lr": a.scaling_lr, "name": "scaling"},
        {"params": [ref["rotation"]], "lr": a.rotation_lr, "name": "rotation"}], lr=0.0, eps=1e-15)
    g = torch.Generator().manual_seed(9)
    for it in range(5):
        lr_xyz = gp.update_learning_rate(it)
        for grp in opt.param_groups:
            if grp["name"] == "xyz":
                grp["lr"] = lr_xyz
        upstream = torch.randn(14 * P, generator=g).cuda() * 10.0 ** float(torch.randint(-6, 1, (1,), generator=g))
        # reference: loss = <activated(raw), upstream>, autograd through exp / sigmoid / normalize
        act = torch.cat([ref["xyz"].reshape(-1), ref["f_dc"].reshape(-1), torch.sigmoid(ref["opacity"]).reshape(-1),
                         torch.exp(ref["scaling"]).reshape(-1), torch.nn.functional.normalize(ref["rotation"]).reshape(-1)])
        opt.zero_grad()
        (act * upstream).sum().backward()
        opt.step()
        gp.adam_step(upstream)
        for name, ours in (("xyz", gp._xyz), ("f_dc", gp._features_dc), ("opacity", gp._opacity), ("scaling", gp._scaling),
                           ("rotation", gp._rotation)):
            torch.testing.assert_close(ours, ref[name].detach(), rtol=5e-6, atol=1e-7, msg=lambda m: f"step {it} {name}: {m}")


def test_densification_stats_match_reference_ops():
    from garmentdreamer_b200.gaussians import GaussianParams
    P, B = 20000, 4
    r = {k: v.cuda() for k, v in _raw(P, 5).items()}
    gp = GaussianParams(**r)
    g = torch.Generator().manual_seed(1)
    acc, den, mx = torch.zeros(P, 1).cuda(), torch.zeros(P, 1).cuda(), torch.zeros(P).cuda()
    for it in range(3):
        grads = [torch.randn(P, 3, generator=g).cuda() for _ in range(B)]
        radii = torch.randint(-1, 40, (B, P), generator=g, dtype=torch.int32).cuda().clamp_min(0)
        # GaussianDreamer.py:189-199,263-275
        rr = radii[0]
        for b in range(1, B):
            rr = torch.max(radii[b], rr)
        vis = rr > 0.0
        gsum = torch.zeros_like(grads[0])
        for x in grads:
            gsum = gsum + x
        mx[vis] = torch.max(mx[vis], rr[vis].float())
        acc[vis] += torch.norm(gsum[vis, :2], dim=-1, keepdim=True)
        den[vis] += 1
        gp.add_densification_stats(gsum, radii)
    assert torch.equal(gp.denom, den) and torch.equal(gp.max_radii2D, mx)
    torch.testing.assert_close(gp.xyz_gradient_accum, acc, rtol=1e-6, atol=1e-7)


def test_batched_cameras_match_reference_construction():
    """f3: one kernel for the view batch vs the CPU restatement of Camera.__init__ (synthetic.camera_from_c2w,
    which follows cameras.py:50-53 incl. the LU double inversion); fp32 4x4 algebra: 2e-6."""
    import math
    from garmentdreamer_b200 import cameras
    from garmentdreamer_b200.synthetic import camera_from_c2w, pose_spherical
    g = torch.Generator().manual_seed(4)
    B, H, W = 8, 512, 384
    c2ws, fovys = [], []
    for i in range(B):
        az, el, d = float(torch.rand(1, generator=g) * 360 - 180), float(torch.rand(1, generator=g) * 90 - 20), float(torch.rand(1, generator=g) * 2.5 + 1.5)
        pose = pose_spherical(az + 90.0, -el, d)
        m = torch.linalg.inv(pose)
        R = -torch.transpose(m[:3, :3], 0, 1)
        R[:, 0] = -R[:, 0]
        c2ws.append(torch.cat([torch.cat([R, (-m[:3, 3])[:, None]], 1), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0))
        fovys.append(math.radians(float(torch.rand(1, generator=g) * 30 + 40)))
    views, packed = cameras.cameras_from_c2w(torch.stack(c2ws).cuda(), fovys, H, W)
    for b in range(B):
        ref = camera_from_c2w(c2ws[b], fovys[b], H, W)
        torch.testing.assert_close(packed[b, 0:16].cpu().view(4, 4), ref.viewmatrix, rtol=2e-6, atol=2e-6)
        torch.testing.assert_close(packed[b, 16:32].cpu().view(4, 4), ref.projmatrix, rtol=2e-6, atol=2e-6)
        torch.testing.assert_close(packed[b, 32:35].cpu(), ref.campos, rtol=2e-6, atol=2e-6)
        assert abs(views[b].tanfovx - ref.tanfovx) < 1e-12 and abs(views[b].tanfovy - ref.tanfovy) < 1e-12
