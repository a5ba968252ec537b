"""GPU: the whole B200 UNet forward and compute_grad_sds against the PyTorch restatement
(oracle/unet_ref.py, "parity unpinned": diffusers is not available offline).

Tolerance. north_star asks for 1e-3 relative fp16 on the SDS gradient. Two fp16 evaluations of
this 860 M-parameter network (different summation orders, roundings at different places) differ
from the fp32 result by more than that, so the test (a) measures the error of the PyTorch-eager
fp16 restatement -- what the reference itself executes -- against fp32, and (b) requires our
error against fp32 to be no worse than max(1e-3, 1.5 x that)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def nets():
    from oracle import unet_ref
    from garmentdreamer_b200.unet import UNetB200
    sd = unet_ref.make_state_dict(0)
    sd32 = {k: v.cuda() for k, v in sd.items()}
    sd16 = {k: v.half() for k, v in sd32.items()}
    net = UNetB200(sd16, "cuda", use_cuda_graph=False)
    return unet_ref, sd32, sd16, net


def _inputs(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 4, 64, 64, generator=g).cuda()
    t = torch.randint(20, 981, (B,), generator=g).cuda()
    ctx = torch.randn(B, 77, 1024, generator=g).cuda()
    return x, t, ctx


def test_unet_forward_matches_restatement(nets):
    unet_ref, sd32, sd16, net = nets
    x, t, ctx = _inputs(2)
    with torch.no_grad():
        ref32 = unet_ref.unet_forward(sd32, x, t, ctx).float()
        ref16 = unet_ref.unet_forward(sd16, x.half(), t.half(), ctx.half()).float()
        ours = net(x.half(), t.half(), encoder_hidden_states=ctx.half()).sample.float()
    assert ours.shape == (2, 4, 64, 64) and torch.isfinite(ours).all()
    e_torch16, e_ours = rel(ref16, ref32), rel(ours, ref32)
    print(f"rel err vs fp32: torch fp16 {e_torch16:.3e}, ours {e_ours:.3e}, ours vs torch fp16 {rel(ours, ref16):.3e}")
    assert e_ours < max(1e-3, 1.5 * e_torch16)


def test_unet_cuda_graph_and_odd_batch(nets):
    unet_ref, sd32, sd16, net = nets
    from garmentdreamer_b200.unet import UNetB200
    x, t, ctx = _inputs(3, seed=5)   # odd batch: 8x8 conv tiles hold two images, last one half empty
    with torch.no_grad():
        eager = net(x.half(), t.half(), encoder_hidden_states=ctx.half()).sample
        gnet = UNetB200.__new__(UNetB200)
        gnet.__dict__.update(net.__dict__)
        gnet.use_cuda_graph, gnet._graphs = True, {}
        g1 = gnet(x.half(), t.half(), encoder_hidden_states=ctx.half()).sample.clone()
        g2 = gnet(x.half(), t.half(), encoder_hidden_states=ctx.half()).sample.clone()
        ref32 = unet_ref.unet_forward(sd32, x, t, ctx).float()
    assert torch.equal(g1, g2) and torch.equal(g1, eager)   # deterministic, graph == eager
    assert rel(eager.float(), ref32) < 5e-3


def test_compute_grad_sds_matches_restatement(nets):
    unet_ref, sd32, sd16, net = nets
    from garmentdreamer_b200.guidance import PromptProcessorOutput, StableDiffusionGuidance
    B = 2
    g = torch.Generator().manual_seed(3)
    bank = lambda n: torch.randn(n, 77, 1024, generator=g).cuda()
    pu = PromptProcessorOutput(bank(1), bank(1), bank(4), bank(4))
    lat = torch.randn(B, 4, 64, 64, generator=g).cuda()
    t = torch.tensor([37, 640], device="cuda")
    elev = torch.tensor([10.0, 75.0], device="cuda")
    azim = torch.tensor([20.0, -170.0], device="cuda")
    dist = torch.tensor([2.0, 3.0], device="cuda")
    guide = StableDiffusionGuidance(net, "cuda", generator=torch.Generator(device="cuda").manual_seed(11))
    grad, aux = guide.compute_grad_sds(lat, t, pu, elev, azim, dist)
    assert set(aux) == {"use_perp_neg", "neg_guidance_weights", "text_embeddings", "t_orig", "latents_noisy", "noise_pred"}
    # restatement of stable_diffusion_guidance.py:229-265 in fp32 with the same noise
    noise = torch.randn(lat.shape, generator=torch.Generator(device="cuda").manual_seed(11), device="cuda")
    emb = pu.get_text_embeddings(elev, azim, dist, True)
    assert torch.equal(emb[0], pu.text_embeddings_vd[1]) and torch.equal(emb[1], pu.text_embeddings_vd[3])
    a = unet_ref.alphas_cumprod().cuda()[t]
    noisy = a.sqrt().view(-1, 1, 1, 1) * lat + (1 - a).sqrt().view(-1, 1, 1, 1) * noise
    assert torch.allclose(aux["latents_noisy"], noisy, atol=1e-6)
    with torch.no_grad():
        eps = unet_ref.unet_forward(sd32, torch.cat([noisy] * 2), torch.cat([t] * 2), emb).float()
    e_text, e_unc = eps.chunk(2)
    ref = (1 - a).view(-1, 1, 1, 1) * (e_text + 100.0 * (e_text - e_unc) - noise)
    e = rel(grad, ref)
    print(f"SDS grad rel err vs fp32 restatement: {e:.3e}")
    # guidance scale 100 amplifies the fp16 error of (e_text - e_uncond); see module docstring
    with torch.no_grad():
        eps16 = unet_ref.unet_forward(sd16, torch.cat([noisy] * 2).half(), torch.cat([t] * 2).half(), emb.half()).float()
    ref16 = (1 - a).view(-1, 1, 1, 1) * (eps16[:B] + 100.0 * (eps16[:B] - eps16[B:]) - noise)
    assert e < max(1e-3, 1.5 * rel(ref16, ref))
