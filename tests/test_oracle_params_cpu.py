"""CPU: the numpy restatement of the parameter path (oracle/params_oracle.py) against torch itself --
the activations, autograd through them, torch.optim.Adam(eps=1e-15) and the densification statistics
of the reference (gaussian_model.py:95-115,156-165,415-419)."""
import numpy as np
import torch

from oracle import params_oracle as po


def _raw(P, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(xyz=torch.randn(P, 3, generator=g), f_dc=torch.randn(P, 1, 3, generator=g), opacity=torch.randn(P, 1, generator=g) * 2,
                scaling=torch.randn(P, 3, generator=g) - 3, rotation=torch.randn(P, 4, generator=g))


def test_activate_and_adam_match_torch():
    P = 777
    raw = _raw(P, 0)
    ref = {k: torch.nn.Parameter(v.clone()) for k, v in raw.items()}
    lrs = dict(xyz=8.75e-5, f_dc=0.0125, opacity=0.01, scaling=0.005, rotation=0.001)
    opt = torch.optim.Adam([{"params": [ref[k]], "lr": lrs[k]} for k in lrs], lr=0.0, eps=1e-15)
    cur = {k: v.numpy().copy() for k, v in raw.items()}
    m = {k: np.zeros_like(v) for k, v in cur.items()}
    v2 = {k: np.zeros_like(v) for k, v in cur.items()}
    g = torch.Generator().manual_seed(9)
    for step in range(1, 5):
        act_t = torch.cat([ref["xyz"].reshape(-1), ref["f_dc"].reshape(-1), torch.sigmoid(ref["opacity"]).reshape(-1),
                           torch.exp(ref["scaling"]).reshape(-1), torch.nn.functional.normalize(ref["rotation"]).reshape(-1)])
        act = po.activate(cur["xyz"], cur["f_dc"], cur["opacity"], cur["scaling"], cur["rotation"])
        np.testing.assert_allclose(act, act_t.detach().numpy(), rtol=2e-6, atol=1e-7)
        up = torch.randn(14 * P, generator=g) * 10.0 ** float(torch.randint(-5, 1, (1,), generator=g))
        opt.zero_grad()
        (act_t * up).sum().backward()
        grads = po.raw_gradients(cur["opacity"], cur["scaling"], cur["rotation"], up.numpy())
        for k, gr in zip(("xyz", "f_dc", "opacity", "scaling", "rotation"), grads):
            np.testing.assert_allclose(gr.reshape(ref[k].shape), ref[k].grad.numpy(), rtol=2e-5, atol=1e-9)
        opt.step()
        for k, gr in zip(("xyz", "f_dc", "opacity", "scaling", "rotation"), grads):
            cur[k], m[k], v2[k] = po.adam_step(cur[k], gr.reshape(cur[k].shape), m[k], v2[k], lrs[k], step)
            np.testing.assert_allclose(cur[k], ref[k].detach().numpy(), rtol=1e-5, atol=1e-7, err_msg=f"step {step} {k}")


def test_densify_stats_match_reference_ops():
    P, B = 4000, 4
    g = torch.Generator().manual_seed(1)
    acc, den, mx = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)
    a2, d2, m2 = acc.numpy().copy(), den.numpy().copy(), mx.numpy().copy()
    for _ in range(3):
        grads = [torch.randn(P, 3, generator=g) for _ in range(B)]
        radii = torch.randint(-1, 40, (B, P), generator=g, dtype=torch.int32).clamp_min(0)
        rr = radii[0]
        for b in range(1, B):
            rr = torch.max(radii[b], rr)
        vis = rr > 0.0
        gsum = sum(grads[1:], grads[0])
        mx[vis] = torch.max(mx[vis], rr[vis].float())
        acc[vis] += torch.norm(gsum[vis, :2], dim=-1, keepdim=True)
        den[vis] += 1
        a2, d2, m2 = po.densify_stats(gsum.numpy(), radii.numpy(), a2, d2, m2)
    assert np.array_equal(d2, den.numpy()) and np.array_equal(m2, mx.numpy())
    np.testing.assert_allclose(a2, acc.numpy(), rtol=1e-6, atol=1e-7)
