"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the per-iteration parameter path of the reference
GaussianModel: the activations (Garment_3DGS/gaussiansplatting/scene/gaussian_model.py:95-115:
identity / identity / sigmoid / exp / F.normalize), the chain rule autograd applies through them,
one torch.optim.Adam(eps=1e-15) step over the parameter groups of :156-165, and
add_densification_stats (:415-419) with the max_radii2D update of
threestudio/systems/GaussianDreamer.py:269-275.

Pinned on the CPU against torch itself (autograd + torch.optim.Adam) by
tests/test_oracle_params_cpu.py; the CUDA kernels gd_params_activate / gd_params_adam /
gd_densify_stats are compared against torch on the GPU (tests/test_params_gpu.py)."""
import numpy as np


def activate(xyz, f_dc, opacity, scaling, rotation):
    """Packed [14P] float32: xyz | f_dc | sigmoid(opacity) | exp(scaling) | normalize(rotation)."""
    n = np.maximum(np.sqrt((rotation.astype(np.float32) ** 2).sum(1, keepdims=True)), 1e-12)
    parts = [xyz, f_dc.reshape(-1, 3), 1.0 / (1.0 + np.exp(-opacity.astype(np.float32))), np.exp(scaling.astype(np.float32)), rotation / n]
    return np.concatenate([p.astype(np.float32).reshape(-1) for p in parts])


def raw_gradients(opacity, scaling, rotation, packed_grad):
    """Chain rule from dL/d(activated) [14P] to the raw parameters (returns 5 arrays)."""
    P = opacity.shape[0]
    g = packed_grad.astype(np.float32)
    g_xyz, g_dc = g[0:3 * P].reshape(P, 3), g[3 * P:6 * P].reshape(P, 1, 3)
    g_op, g_sc, g_rot = g[6 * P:7 * P].reshape(P, 1), g[7 * P:10 * P].reshape(P, 3), g[10 * P:14 * P].reshape(P, 4)
    sg = 1.0 / (1.0 + np.exp(-opacity.astype(np.float32)))
    nrm = np.sqrt((rotation.astype(np.float32) ** 2).sum(1, keepdims=True))
    n = np.maximum(nrm, 1e-12)
    qh = rotation / n
    dot = np.where(nrm > 1e-12, (qh * g_rot).sum(1, keepdims=True), 0.0)
    return g_xyz, g_dc, g_op * sg * (1.0 - sg), g_sc * np.exp(scaling.astype(np.float32)), (g_rot - qh * dot) / n


def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-15):
    """torch.optim.Adam single-tensor update (no weight decay / amsgrad); returns (p, m, v)."""
    p, g, m, v = (a.astype(np.float32) for a in (p, g, m, v))
    m = m + np.float32(1.0 - beta1) * (g - m)
    v = v * np.float32(beta2) + np.float32(1.0 - beta2) * g * g
    bc1 = np.float32(1.0 - beta1 ** step)
    bc2_sqrt = np.float32(np.sqrt(1.0 - beta2 ** step))
    return p - np.float32(lr) / bc1 * (m / (np.sqrt(v) / bc2_sqrt + np.float32(eps))), m, v


def densify_stats(dmeans2D_sum, radii, xyz_gradient_accum, denom, max_radii2D):
    """radii int [B,P]; visibility = max radius over the views > 0. Returns the three updated arrays."""
    r = radii.max(0)
    vis = r > 0
    acc, den, mx = xyz_gradient_accum.copy(), denom.copy(), max_radii2D.copy()
    acc[vis, 0] += np.sqrt((dmeans2D_sum[vis, :2].astype(np.float32) ** 2).sum(1))
    den[vis, 0] += 1.0
    mx[vis] = np.maximum(mx[vis], r[vis].astype(np.float32))
    return acc, den, mx
