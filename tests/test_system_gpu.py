"""GPU: the pieces either side of the rasteriser inside the full iteration and the
iteration harness itself (garmentdreamer_b200/system.py) against what the reference executes -- torch
autograd through `opacity = depths / (depths.max() + 1e-5)`, `(opacity**2 + 0.01).sqrt().mean()`
(TS/systems/GaussianDreamer.py:215,253-255), torch.optim.Adam(eps=1e-15) over the activations of
GS/scene/gaussian_model.py:95-115 -- driven by the SAME rasteriser (the drop-in autograd module)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("ties", [False, True])
def test_sparsity_loss_and_gradient_match_autograd(ties):
    from garmentdreamer_b200.system import SparsityLoss
    g = torch.Generator().manual_seed(0)
    depth = (torch.rand(3, 1, 70, 50, generator=g) * 4.0).cuda()
    depth[depth < 1.0] = 0.0                       # background pixels
    if ties:
        depth[0, 0, 3, 4] = depth[2, 0, 60, 7] = 5.0   # two arg-max pixels: autograd splits the max() term evenly
    sp = SparsityLoss(depth.device)
    dmax = depth.max().reshape(1).clone()
    for lam, n_total in ((1.0, depth.numel()), (0.3, 2 * depth.numel())):
        gd = sp.grad(depth, dmax, lam, n_total)
        x = depth.clone().requires_grad_(True)
        op = x / (x.max() + 1e-5)
        loss = lam * (op ** 2 + 0.01).sqrt().sum() / n_total
        loss.backward()
        assert rel(sp.loss, loss.reshape(1)) < 1e-6
        assert rel(gd, x.grad) < 1e-5, (ties, lam)


def test_depth_max_comes_out_of_the_forward_compositor():
    from garmentdreamer_b200 import raster
    from garmentdreamer_b200.synthetic import garment, sample_cameras
    from garmentdreamer_b200.system import SparsityLoss
    dev = torch.device("cuda:0")
    g = {k: v.to(dev) for k, v in garment(20000, 0).items()}
    views = [raster.View(c.viewmatrix.to(dev), c.projmatrix.to(dev), c.campos.to(dev), c.tanfovx, c.tanfovy)
             for c in sample_cameras(3, 200, 136)]
    col, dep, alp, rad, st = raster.forward_views(g["xyz"], g["opacity"], views, 136, 200, torch.ones(3, device=dev),
                                                  shs=g["shs"], scales=g["scales"], rotations=g["rotations"])
    assert float(SparsityLoss(dev).depth_max_of(st)) == float(dep.max())


def test_training_step_matches_autograd_through_the_dropin_rasterizer():
    """One full iteration (no guidance network: a seeded upstream colour gradient stands in for dL_sds/dcolour)
    against torch autograd on the reference's formulation, rasterised per view by the drop-in
    `diff_gaussian_rasterization` module like GS/gaussian_renderer/__init__.py:18-103. Both sides see the SAME
    activated parameters and cameras (the ones the step built on the GPU): with a random upstream gradient and the
    max() term of the sparsity loss landing on ONE pixel, a 1e-7 difference in an input moves individual gradients by
    1e-3, which would hide a real bug. The chain rule through the activations + Adam against torch.optim.Adam is
    tests/test_params_gpu.py; here the updated raw parameters are checked to be what that kernel makes of these
    gradients."""
    import diff_gaussian_rasterization as dgr
    from garmentdreamer_b200.gaussians import GaussianParams
    from garmentdreamer_b200.synthetic import garment, raw_params, sample_batch, sample_cameras
    from garmentdreamer_b200.system import GaussianDreamerB200
    dev = torch.device("cuda:0")
    P, B, S = 6000, 3, 128
    raw = {k: v.to(dev) for k, v in raw_params(garment(P, 0)).items()}
    gp = GaussianParams(raw["xyz"], raw["f_dc"], raw["opacity"], raw["scaling"], raw["rotation"], spatial_lr_scale=4.0)
    gp.training_setup()
    system = GaussianDreamerB200(gp, None)
    batch = sample_batch(B, S, S)
    dcol = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(7)).to(dev) * 1e-3
    batch["dL_dcolor"] = dcol
    out = system.training_step(batch)
    cams = sample_cameras(B, S, S)
    for c, v in zip(cams, system._fw[0]):     # GPU-built cameras == the CPU construction of the reference
        assert torch.allclose(v.viewmatrix.cpu().view(4, 4), c.viewmatrix, atol=1e-5)
        assert torch.allclose(v.projmatrix.cpu().view(4, 4), c.projmatrix, atol=1e-4)
    act = [t.clone().requires_grad_(True) for t in GaussianParams.unpack(system.packed, P)]   # xyz, shs, opacity, scales, rot
    ref_act = (raw["xyz"], raw["f_dc"].view(P, 1, 3), torch.sigmoid(raw["opacity"]), torch.exp(raw["scaling"]), F.normalize(raw["rotation"]))
    for a_, r_ in zip(act, ref_act):          # gd_params_activate == the reference activations
        assert rel(a_, r_.view(a_.shape)) < 1e-6
    bg = torch.ones(3, device=dev)
    images, depths, vsp = [], [], []
    for c in system._fw[0]:
        settings = dgr.GaussianRasterizationSettings(
            image_height=S, image_width=S, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
            viewmatrix=c.viewmatrix.view(4, 4), projmatrix=c.projmatrix.view(4, 4), sh_degree=0, campos=c.campos,
            prefiltered=False, debug=False)
        m2d = torch.zeros_like(act[0], requires_grad=True)
        img, radii, depth, alpha = dgr.GaussianRasterizer(settings)(
            means3D=act[0], means2D=m2d, shs=act[1], colors_precomp=None, opacities=act[2], scales=act[3], rotations=act[4],
            cov3D_precomp=None)
        images.append(img); depths.append(depth); vsp.append(m2d)
    images, depths = torch.stack(images), torch.stack(depths)
    opacity = depths / (depths.max() + 1e-5)                                       # GaussianDreamer.py:215
    loss_sparsity = (opacity ** 2 + 0.01).sqrt().mean()                            # :253
    ((images * dcol).sum() + loss_sparsity).backward()
    assert rel(out["loss_sparsity"], loss_sparsity.reshape(1)) < 1e-5
    assert rel(system.grad[14 * P:].view(P, 3), sum(v.grad for v in vsp)) < 1e-5   # viewspace_point_tensor grads (:268-270)
    for ours, a_ in zip(GaussianParams.unpack(system.grad[:14 * P], P), act):
        assert rel(ours, a_.grad.view(ours.shape)) < 1e-5
    assert float(gp.denom.sum()) > 0 and float(gp.max_radii2D.max()) > 0
    # the optimiser step the iteration took == gd_params_adam on a fresh copy fed these gradients
    gp2 = GaussianParams(raw["xyz"], raw["f_dc"], raw["opacity"], raw["scaling"], raw["rotation"], spatial_lr_scale=4.0)
    gp2.training_setup()
    gp2.update_learning_rate(0)
    gp2.adam_step(system.grad[:14 * P].clone())
    for k in ("_xyz", "_features_dc", "_opacity", "_scaling", "_rotation"):
        assert torch.equal(getattr(gp, k), getattr(gp2, k)) and not torch.equal(getattr(gp, k), raw[{"_features_dc": "f_dc"}.get(k, k[1:])].view(getattr(gp, k).shape)), k


def test_changing_P_between_steps_resizes_every_buffer():
    """Densification changes P every 100 steps (GaussianDreamer.py:281-283): arenas, packed buffers and the
    all-reduce buffer are re-allocated; the step after a resize equals a fresh system's step bit for bit."""
    from garmentdreamer_b200.gaussians import GaussianParams
    from garmentdreamer_b200.synthetic import garment, raw_params, sample_batch
    from garmentdreamer_b200.system import GaussianDreamerB200
    dev = torch.device("cuda:0")
    B, S = 2, 128
    batch = sample_batch(B, S, S)
    batch["dL_dcolor"] = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(1)).to(dev) * 1e-3

    def make(P):
        raw = {k: v.to(dev) for k, v in raw_params(garment(P, 0)).items()}
        gp = GaussianParams(raw["xyz"], raw["f_dc"], raw["opacity"], raw["scaling"], raw["rotation"])
        gp.training_setup()
        return gp
    a = GaussianDreamerB200(make(3000), None)
    a.training_step(batch)
    a.gaussian = make(7001)              # "densified": more Gaussians, P not a multiple of anything
    a.training_step(batch)
    assert a.grad.numel() == 17 * 7001 and a.packed.numel() == 14 * 7001
    b = GaussianDreamerB200(make(7001), None)
    b.true_global_step = 1
    b.training_step(batch)
    assert torch.equal(a.grad, b.grad) and torch.equal(a.gaussian._xyz, b.gaussian._xyz)
    a.gaussian = make(1200)              # "pruned"
    a.training_step(batch)
    assert a.grad.numel() == 17 * 1200 and torch.isfinite(a.grad).all()


def test_peer_exchange_world1_matches_plain_adam():
    """parallel.PeerExchange (gd_peer_allreduce + gd_params_adam_peers) on a one-rank group: flags, slices, padding (P % 4 != 0)
    and the fused statistics + Adam kernel against gd_densify_stats + gd_params_adam. (N > 1: tools/peer_exchange_check.py,
    profiles/r02_peer_exchange_n*.json.)"""
    import os
    import torch.distributed as dist
    from garmentdreamer_b200 import parallel
    from garmentdreamer_b200.gaussians import GaussianParams
    from garmentdreamer_b200.synthetic import garment, raw_params
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        P = 7001
        dev = torch.device("cuda", 0)
        try:
            px = parallel.PeerExchange(P, dev)
        except Exception as e:   # noqa: BLE001
            pytest.skip(f"symmetric memory unavailable here: {e}")
        raw = {k: v.to(dev) for k, v in raw_params(garment(P, 0)).items()}
        mk = lambda: GaussianParams(raw["xyz"], raw["f_dc"], raw["opacity"], raw["scaling"], raw["rotation"], spatial_lr_scale=4.0)
        ga, gb = mk(), mk()
        ga.training_setup(); gb.training_setup()
        for step in range(3):
            g = torch.Generator(device=dev).manual_seed(step)
            grad = torch.randn(17 * P, generator=g, device=dev) * 1e-3
            radii = torch.randint(0, 40, (P,), generator=g, device=dev, dtype=torch.int32)
            px.grad.copy_(grad); px.radii.copy_(radii)
            ga.update_learning_rate(step); gb.update_learning_rate(step)
            ga.add_densification_stats(grad[14 * P:].view(P, 3), radii.view(1, P))
            ga.adam_step(grad[:14 * P])
            px.allreduce()
            gb.adam_step_peers(px)
            torch.cuda.synchronize()
            assert torch.equal(px.red_grad, grad) and torch.equal(px.red_radii, radii)
            for a, b in ((ga._xyz, gb._xyz), (ga._features_dc, gb._features_dc), (ga._opacity, gb._opacity), (ga._scaling, gb._scaling),
                         (ga._rotation, gb._rotation), (ga.exp_avg, gb.exp_avg), (ga.exp_avg_sq, gb.exp_avg_sq),
                         (ga.xyz_gradient_accum, gb.xyz_gradient_accum), (ga.denom, gb.denom), (ga.max_radii2D, gb.max_radii2D)):
                assert torch.equal(a, b)
    finally:
        if created:
            dist.destroy_process_group()


def test_peer_exchange_two_ranks_matches_nccl():
    """N = 2 on one box: tools/peer_exchange_check.py under torchrun (gradient SUM / radii MAX / fused statistics + Adam through
    peer memory against NCCL all-reduce + the plain kernels, replicas bit-identical). Skipped on a single-GPU box; its N = 2 and
    N = 8 outputs are committed under profiles/r02_peer_exchange_n{2,8}.json."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "tools", "peer_exchange_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["world"] == 2 and res["P100000_p2p"]["max_rel_err_vs_nccl"] < 1e-6
