"""CPU: the drop-in package keeps the reference's Python surface
(DGR/diff_gaussian_rasterization/__init__.py:21-42,160-223) and refuses to run without CUDA."""
import inspect

import pytest
import torch

import diff_gaussian_rasterization as dgr


def _settings(dev="cpu"):
    return dgr.GaussianRasterizationSettings(
        image_height=32, image_width=32, tanfovx=0.5, tanfovy=0.5, bg=torch.ones(3, device=dev),
        scale_modifier=1.0, viewmatrix=torch.eye(4, device=dev), projmatrix=torch.eye(4, device=dev),
        sh_degree=0, campos=torch.zeros(3, device=dev), prefiltered=False, debug=False)


def test_surface():
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
        "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    assert list(inspect.signature(dgr.rasterize_gaussians).parameters) == [
        "means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations",
        "cov3Ds_precomp", "raster_settings"]
    assert list(inspect.signature(dgr.GaussianRasterizer.forward).parameters) == [
        "self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations",
        "cov3D_precomp"]
    for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert hasattr(dgr._C, fn)


def test_argument_validation_messages():
    r = dgr.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m, torch.ones(4, 1), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m, torch.ones(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=m, scales=m,
          rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.ones(4, 1), shs=torch.zeros(4, 1, 3), scales=m)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.ones(4, 1), shs=torch.zeros(4, 1, 3), scales=m, rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    r = dgr.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        r(m, m, torch.ones(4, 1), shs=torch.zeros(4, 1, 3), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        r.markVisible(m)


def test_synthetic_workload_is_deterministic():
    from garmentdreamer_b200.synthetic import garment, sample_cameras
    a, b = garment(1000, 0), garment(1000, 0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.allclose(a["rotations"].norm(dim=1), torch.ones(1000), atol=1e-6)
    c1, c2 = sample_cameras(4, 64, 64), sample_cameras(4, 64, 64)
    assert all(torch.equal(x.viewmatrix, y.viewmatrix) and torch.equal(x.projmatrix, y.projmatrix)
               for x, y in zip(c1, c2))
    for c in c1:
        assert -22.0 <= c.elevation_deg <= 70.0 and 1.5 <= c.distance <= 4.0
        # camera centre is `distance` away from the origin and the view matrix is rigid
        assert abs(float(c.campos.norm()) - c.distance) < 1e-4
        R = c.viewmatrix[:3, :3]
        assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-5)
