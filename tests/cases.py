"""Seeded rasteriser test cases shared by the golden generator and the parity tests."""
import math

import numpy as np
import torch

from garmentdreamer_b200.synthetic import camera_from_c2w, garment, sample_cameras


def _look_at_camera(dist, H, W, fovy_deg=50.0, az=30.0, el=10.0):
    from garmentdreamer_b200.synthetic import pose_spherical
    pose = pose_spherical(az + 90.0, -el, dist)
    m = torch.linalg.inv(pose)
    R = -torch.transpose(m[:3, :3], 0, 1)
    R[:, 0] = -R[:, 0]
    T = -m[:3, 3]
    c2w = torch.cat([torch.cat([R, T[:, None]], 1), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0)
    return camera_from_c2w(c2w, fovy_deg * math.pi / 180, H, W)


def make_case(name):
    """Returns dict(inputs) with numpy arrays: everything one forward+backward of one view needs."""
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    c = {"name": name, "sh_degree": 0, "scale_modifier": 1.0}
    if name == "garment_small":
        P, W, H = 1500, 96, 96
        g = garment(P, 0); cam = sample_cameras(4, H, W)[1]
    elif name == "garment_ragged":  # image size not a multiple of the 16-pixel tile
        P, W, H = 1200, 100, 70
        g = garment(P, 1); cam = sample_cameras(4, H, W)[2]
    elif name == "close_big_splats":  # camera inside the near-cull range of some Gaussians
        P, W, H = 800, 80, 80
        g = garment(P, 2); g["scales"] = g["scales"] * 4.0
        cam = _look_at_camera(0.75, H, W)
    elif name == "sh3":
        P, W, H = 900, 64, 64
        g = garment(P, 3); cam = sample_cameras(4, H, W)[0]
        g["shs"] = torch.from_numpy(rng.normal(0, 0.6, (P, 16, 3)).astype(np.float32))
        c["sh_degree"] = 3
    elif name == "precomp":  # colors_precomp + cov3D_precomp, scale_modifier != 1 (unused)
        P, W, H = 700, 64, 48
        g = garment(P, 4); cam = sample_cameras(4, H, W)[3]
    elif name == "all_culled":  # everything behind the camera -> num_rendered == 0
        P, W, H = 300, 48, 48
        g = garment(P, 5); cam = _look_at_camera(3.0, H, W)
        g["xyz"] = g["xyz"] + torch.tensor([100.0, 0.0, 0.0])
    elif name == "dense_tile":  # > 2 x 4096 instances in each of the 4 central tiles (chunked-merge sort path)
        P, W, H = 9000, 64, 64
        g = garment(P, 6); cam = _look_at_camera(2.5, H, W)
        g["xyz"] = torch.from_numpy(rng.normal(0, 0.012, (P, 3)).astype(np.float32))
        g["opacity"] = torch.from_numpy(rng.uniform(0.01, 0.08, (P, 1)).astype(np.float32))
    elif name == "c1":  # BASELINE config 1 geometry: 10k Gaussians, 1 camera 256^2
        P, W, H = 10000, 256, 256
        g = garment(P, 0); cam = sample_cameras(4, H, W)[1]
    else:
        raise KeyError(name)
    c.update(P=P, W=W, H=H)
    c["means3D"] = g["xyz"].numpy(); c["opacities"] = g["opacity"].numpy()
    c["scales"] = g["scales"].numpy(); c["rotations"] = g["rotations"].numpy()
    c["shs"] = g["shs"].numpy()
    c["colors_precomp"] = None; c["cov3D_precomp"] = None
    if name == "precomp":
        c["colors_precomp"] = rng.uniform(0, 1, (P, 3)).astype(np.float32)
        s, q = c["scales"], c["rotations"]
        cov = np.zeros((P, 6), np.float32)
        for i in range(P):
            r, x, y, z = q[i]
            Rm = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                           [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                           [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]], np.float64)
            Mm = Rm @ np.diag(s[i].astype(np.float64))
            S = Mm @ Mm.T
            cov[i] = [S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]]
        c["cov3D_precomp"] = cov
        c["shs"] = None; c["scales"] = None; c["rotations"] = None
    c["viewmatrix"] = cam.viewmatrix.numpy().copy(); c["projmatrix"] = cam.projmatrix.numpy().copy()
    c["campos"] = cam.campos.numpy().copy()
    c["tanfovx"], c["tanfovy"] = float(cam.tanfovx), float(cam.tanfovy)
    c["bg"] = np.array([1.0, 1.0, 1.0], np.float32) if name != "sh3" else np.array([0.1, 0.5, 0.9], np.float32)
    g7 = torch.Generator().manual_seed(7)
    c["dL_dcolor"] = torch.randn(3, H, W, generator=g7).numpy()
    c["dL_ddepth"] = torch.randn(1, H, W, generator=g7).numpy()
    c["dL_dalpha"] = torch.randn(1, H, W, generator=g7).numpy()
    return c


SMALL_CASES = ["garment_small", "garment_ragged", "close_big_splats", "sh3", "precomp", "all_culled"]


def oracle_run(c, backward=True):
    from oracle import raster_oracle as ro
    st = ro.forward(c["means3D"], c["opacities"], c["viewmatrix"], c["projmatrix"], c["campos"],
                    c["W"], c["H"], c["tanfovx"], c["tanfovy"], c["bg"], shs=c["shs"],
                    colors_precomp=c["colors_precomp"], scales=c["scales"], rotations=c["rotations"],
                    cov3D_precomp=c["cov3D_precomp"], scale_modifier=c["scale_modifier"],
                    sh_degree=c["sh_degree"])
    g = ro.backward(st, c["dL_dcolor"], c["dL_ddepth"], c["dL_dalpha"]) if backward else None
    return st, g


def to_cuda(c, dev="cuda"):
    t = {}
    for k, v in c.items():
        t[k] = torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v
    return t


def ours_run(c, backward=True, batched_copies=1):
    """Runs the product CUDA path through the C ABI (garmentdreamer_b200.raster)."""
    from garmentdreamer_b200 import raster
    t = to_cuda(c)
    view = raster.View(t["viewmatrix"], t["projmatrix"], t["campos"], c["tanfovx"], c["tanfovy"])
    views = [view] * batched_copies
    color, depth, alpha, radii, state = raster.forward_views(
        t["means3D"], t["opacities"], views, c["W"], c["H"], t["bg"], shs=t["shs"],
        colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"],
        cov3D_precomp=t["cov3D_precomp"], scale_modifier=c["scale_modifier"],
        sh_degree=c["sh_degree"])
    out = {"color": color, "depth": depth, "alpha": alpha, "radii": radii, "state": state}
    if backward:
        B = batched_copies
        rep = lambda x: x.unsqueeze(0).expand(B, *x.shape).contiguous()
        out["grads"] = raster.backward_views(
            state, t["means3D"], radii, alpha, t["bg"], rep(t["dL_dcolor"]), rep(t["dL_ddepth"]),
            rep(t["dL_dalpha"]), shs=t["shs"], colors_precomp=t["colors_precomp"],
            scales=t["scales"], rotations=t["rotations"], cov3D_precomp=t["cov3D_precomp"],
            scale_modifier=c["scale_modifier"], sum_views=False, want_aux=True)
    return out


def ref_run(c, backward=True):
    """Runs the unmodified reference CUDA core (oracle/_ref) on the same inputs."""
    from oracle.ref_cuda import RefRasterizer
    t = to_cuda(c)
    rr = RefRasterizer()
    kw = dict(shs=t["shs"], colors_precomp=t["colors_precomp"], scales=t["scales"],
              rotations=t["rotations"], cov3D_precomp=t["cov3D_precomp"],
              scale_modifier=c["scale_modifier"])
    out = rr.forward(t["means3D"], t["opacities"].reshape(-1).contiguous(), t["viewmatrix"],
                     t["projmatrix"], t["campos"], c["W"], c["H"], c["tanfovx"], c["tanfovy"],
                     t["bg"], sh_degree=c["sh_degree"], **kw)
    st = rr.state()
    g = None
    if backward and c["P"] > 0:
        g = rr.backward(t["means3D"], out["radii"], out["alpha"], t["viewmatrix"], t["projmatrix"],
                        t["campos"], c["tanfovx"], c["tanfovy"], t["bg"], t["dL_dcolor"],
                        t["dL_ddepth"], t["dL_dalpha"], **kw)
    return out, st, g
