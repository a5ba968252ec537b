"""CPU: host-side guidance logic mirrors the reference (prompt direction selection, schedules)."""
import torch

from garmentdreamer_b200.guidance import C, PromptProcessorOutput, ddim_alphas_cumprod, shift_azimuth_deg


def _pu():
    mk = lambda n, v: torch.full((n, 77, 8), float(v))
    vd = torch.stack([torch.full((77, 8), float(i)) for i in range(4)])
    return PromptProcessorOutput(mk(1, 9), mk(1, -9), vd, -vd)


def test_direction_selection_overwrite_order():
    # base.py:62-66 -- side default, front |az|<45, back |az|>135, overhead elev>60 wins last
    elev = torch.tensor([0.0, 0.0, 0.0, 70.0, 0.0])
    azim = torch.tensor([90.0, 10.0, 170.0, 10.0, -140.0])
    emb = _pu().get_text_embeddings(elev, azim, torch.ones(5), True)
    assert emb.shape == (10, 77, 8)
    assert [int(emb[i, 0, 0]) for i in range(5)] == [0, 1, 2, 3, 2]       # cond first ...
    assert [int(emb[5 + i, 0, 0]) for i in range(5)] == [0, -1, -2, -3, -2]  # ... then uncond
    emb = _pu().get_text_embeddings(elev, azim, torch.ones(5), False)
    assert float(emb[0, 0, 0]) == 9.0 and float(emb[5, 0, 0]) == -9.0


def test_perp_neg_shapes():
    pu = _pu()
    emb, w = pu.get_text_embeddings_perp_neg(torch.tensor([0.0, 70.0]), torch.tensor([30.0, 0.0]), torch.ones(2))
    assert emb.shape == (8, 77, 8) and w.shape == (2, 2)
    assert float(w[1].abs().sum()) == 0.0  # overhead view: dummy negatives with zero weight


def test_schedules():
    assert float(shift_azimuth_deg(torch.tensor(190.0))) == -170.0
    assert C([0, 1.5, 2.0, 1000], 0, 500) == 1.75 and C(0.98, 0, 7) == 0.98
    a = ddim_alphas_cumprod()
    assert a.shape == (1000,) and abs(float(a[0]) - (1 - 0.00085)) < 1e-6 and float(a[-1]) < 0.01
    assert bool((a[1:] < a[:-1]).all())
