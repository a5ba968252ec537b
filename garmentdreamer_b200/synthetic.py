"""Synthetic workload of SURVEY.md s.8(d): a garment-like tube of Gaussians and the reference's
random camera sampler. Host-side, CPU tensors only (the callers move them to the GPU).

Cameras restate Garment_3DGS/threestudio/data/uncond.py:190-408 (elevation / azimuth / distance /
fovy sampling, pose_spherical :49-54, c2w_3dgs :371-390) and the matrix construction of
Garment_3DGS/gaussiansplatting/scene/cameras.py:17-53 + utils/graphics_utils.py:59-93.
"""
import math
import random
from dataclasses import dataclass
from typing import List

import numpy as np
import torch

SH_C0 = 0.28209479177387814


def garment(P: int, seed: int = 0):
    """Returns dict of *activated* fp32 parameters: xyz[P,3], scales[P,3], rotations[P,4] (unit),
    opacity[P,1], shs[P,1,3] (sh_degree 0)."""
    rng = np.random.default_rng(seed)
    theta = rng.uniform(0.0, 2.0 * np.pi, P)
    h = rng.uniform(-0.5, 0.5, P)
    r = 0.30 + 0.08 * h
    xyz = np.stack([r * np.cos(theta), 0.6 * r * np.sin(theta), h], -1) * 1.75
    s0 = math.sqrt(2.0 * np.pi * 0.25 * 1.75 ** 2 / P)
    scales = s0 * rng.uniform(0.7, 1.5, (P, 3))
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opacity = rng.uniform(0.05, 0.95, (P, 1))
    f_dc = (rng.uniform(0.0, 1.0, (P, 1, 3)) - 0.5) / SH_C0
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return {"xyz": f(xyz), "scales": f(scales), "rotations": f(q), "opacity": f(opacity),
            "shs": f(f_dc)}


def _trans_t(t):
    return torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=torch.float32)


def _rot_phi(phi):
    return torch.tensor([[1, 0, 0, 0], [0, np.cos(phi), -np.sin(phi), 0],
                         [0, np.sin(phi), np.cos(phi), 0], [0, 0, 0, 1]], dtype=torch.float32)


def _rot_theta(th):
    return torch.tensor([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0],
                         [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], dtype=torch.float32)


def pose_spherical(theta, phi, radius):
    c2w = _trans_t(radius)
    c2w = _rot_phi(phi / 180.0 * np.pi) @ c2w
    c2w = _rot_theta(theta / 180.0 * np.pi) @ c2w
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ c2w


@dataclass
class CameraSample:
    viewmatrix: torch.Tensor   # [4,4] world_view_transform (transposed w2c)
    projmatrix: torch.Tensor   # [4,4] full_proj_transform
    campos: torch.Tensor       # [3]
    tanfovx: float
    tanfovy: float
    elevation_deg: float
    azimuth_deg: float
    distance: float
    fovy: float
    c2w: torch.Tensor = None   # [4,4] camera-to-world (batch['c2w_3dgs'] of the reference data module)


def _projection(znear, zfar, fovx, fovy):
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def camera_from_c2w(c2w: torch.Tensor, fovy: float, height: int, width: int, **meta) -> CameraSample:
    focal = height / (2 * math.tan(fovy / 2))
    fovx = 2 * math.atan(width / (2 * focal))
    R, T = c2w[:3, :3].float(), c2w[:3, 3].float()
    Rt = torch.zeros(4, 4)
    Rt[:3, :3] = R.transpose(0, 1)
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    Rt = torch.linalg.inv(torch.linalg.inv(Rt)).float()  # graphics_utils.py:59-70 (no recentring)
    wvt = Rt.transpose(0, 1).contiguous()
    proj = _projection(0.01, 100.0, fovx, fovy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).float().contiguous()
    campos = wvt.inverse()[3, :3].float().contiguous()
    return CameraSample(wvt, full, campos, math.tan(fovx * 0.5), math.tan(fovy * 0.5),
                        meta.get("elevation_deg", 0.0), meta.get("azimuth_deg", 0.0),
                        meta.get("distance", 0.0), fovy, c2w.float().contiguous())


def sample_cameras(B: int, height: int, width: int, seed: int = 123,
                   elevation_range=(-22.0, 70.0), azimuth_range=(-180.0, 180.0),
                   distance_range=(1.5, 4.0), fovy_range=(40.0, 70.0)) -> List[CameraSample]:
    """One training batch of B cameras with the reference sampler's distribution and RNG order."""
    g = torch.Generator().manual_seed(seed)
    rnd = random.Random(seed)
    if rnd.random() < 0.5:
        elevation_deg = torch.rand(B, generator=g) * (elevation_range[1] - elevation_range[0]) + elevation_range[0]
    else:
        pr = [(elevation_range[0] + 90.0) / 180.0, (elevation_range[1] + 90.0) / 180.0]
        elevation = torch.asin(2 * (torch.rand(B, generator=g) * (pr[1] - pr[0]) + pr[0]) - 1.0)
        elevation_deg = elevation / math.pi * 180.0
    azimuth_deg = (torch.rand(B, generator=g) + torch.arange(B)) / B * (
        azimuth_range[1] - azimuth_range[0]) + azimuth_range[0]  # batch_uniform_azimuth
    dist = torch.rand(B, generator=g) * (distance_range[1] - distance_range[0]) + distance_range[0]
    fovy_deg = torch.rand(B, generator=g) * (fovy_range[1] - fovy_range[0]) + fovy_range[0]
    fovy = fovy_deg * math.pi / 180
    cams = []
    for i in range(B):
        pose = pose_spherical(float(azimuth_deg[i]) + 180.0 - 90, -float(elevation_deg[i]), float(dist[i]))
        m = torch.linalg.inv(pose)
        R = -torch.transpose(m[:3, :3], 0, 1)
        R[:, 0] = -R[:, 0]
        T = -m[:3, 3]
        c2w = torch.cat([torch.cat([R, T[:, None]], 1), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0)
        cams.append(camera_from_c2w(c2w, float(fovy[i]), height, width,
                                    elevation_deg=float(elevation_deg[i]),
                                    azimuth_deg=float(azimuth_deg[i]), distance=float(dist[i])))
    return cams


def sample_batch(B: int, height: int, width: int, seed: int = 123, lo: int = 0, hi: int = None):
    """The training batch dict of the reference data module (threestudio/data/uncond.py:190-408) for the
    cameras [lo, hi) of a B-view batch: c2w_3dgs [b,4,4], fovy [b] (radians), elevation / azimuth (degrees),
    camera_distances, height, width -- host tensors, as a DataLoader worker would hand them over."""
    cams = sample_cameras(B, height, width, seed)[lo:hi]
    f = lambda xs: torch.tensor(xs, dtype=torch.float32)
    return {"c2w_3dgs": torch.stack([c.c2w for c in cams]), "fovy": f([c.fovy for c in cams]),
            "elevation": f([c.elevation_deg for c in cams]), "azimuth": f([c.azimuth_deg for c in cams]),
            "camera_distances": f([c.distance for c in cams]), "height": height, "width": width}


def raw_params(g):
    """Inverse activations of garment(): raw xyz, f_dc, opacity (logit), scaling (log), rotation as stored by the
    reference GaussianModel (scene/gaussian_model.py:41-50)."""
    op = g["opacity"].clamp(1e-6, 1 - 1e-6)
    return {"xyz": g["xyz"], "f_dc": g["shs"], "opacity": torch.log(op / (1 - op)), "scaling": torch.log(g["scales"]),
            "rotation": g["rotations"]}
