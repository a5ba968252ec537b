"""BASELINE config 4(b): the UNet call pattern of the normal-VSD loop of Garment_Deformer_NeTF
(netf/guidance/sd_vsd_utils.py:164-213): per step two no-grad forwards -- the pretrained UNet on batch 2
(cond + uncond, classifier-free guidance 7.5 with the UNCOND prediction as the base term) and the second
("q") UNet on batch 1 -- followed by grad = w(t) * (noise_pred - noise_pred_q). The LoRA / camera-embedding
additions of the q-UNet and its training step (netf/trainer.py:228-257) are out of scope:
the q-UNet is timed as a second SD-2.1 UNet with its own weights. Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garmentdreamer_b200.unet import UNetB200          # noqa: E402
from garmentdreamer_b200.unet_init import random_state_dict  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    unet = UNetB200(random_state_dict(0, dev, torch.float16), dev, use_cuda_graph=True)
    q_unet = UNetB200(random_state_dict(1, dev, torch.float16), dev, use_cuda_graph=True)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 4, 64, 64, generator=g).to(dev)
    emb = torch.randn(2, 77, 1024, generator=g).to(dev).half()
    alphas = torch.cumprod(1.0 - torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2, 0).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)

    def step():
        t = torch.randint(20, 981, (1,), device=dev, generator=gen)
        noise = torch.randn(lat.shape, device=dev, generator=gen)
        a = alphas[t].view(1, 1, 1, 1)
        noisy = a.sqrt() * lat + (1 - a).sqrt() * noise
        eps = unet(torch.cat([noisy] * 2).half(), torch.cat([t] * 2).half(), encoder_hidden_states=emb).sample.float()
        e_c, e_u = eps.chunk(2)
        noise_pred = e_u + 7.5 * (e_c - e_u)
        v_q = q_unet(noisy.half(), t.half(), encoder_hidden_states=emb[:1]).sample.float()
        noise_pred_q = a.sqrt() * v_q + (1 - a).sqrt() * noisy          # :199-208
        return torch.nan_to_num((1 - a) * (noise_pred - noise_pred_q))

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    flops = 804.3e9 * 3
    out = {"config": "c4b: VSD UNet call pattern, batch 2 (CFG 7.5) + batch 1, 64x64 latents, fp16 (the reference holds fp32 weights, sd_vsd_utils.py:35)",
           "ms_per_step": ms, "steps_per_s": 1e3 / ms, "tflops": flops / (ms * 1e-3) / 1e12, "algorithmic_flops": flops}
    if "--ref" in sys.argv:
        from oracle import unet_ref
        sd = {k: v.to(dev).half() for k, v in unet_ref.make_state_dict(0).items()}
        x2, t2 = torch.randn(2, 4, 64, 64, device=dev).half(), torch.tensor([500, 500], device=dev).half()
        rs = []
        with torch.no_grad():
            for it in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                unet_ref.unet_forward(sd, x2, t2, emb)
                unet_ref.unet_forward(sd, x2[:1], t2[:1], emb[:1])
                e1.record()
                torch.cuda.synchronize()
                rs.append(e0.elapsed_time(e1))
        out["eager_fp16_restatement_ms"] = sorted(rs[2:])[len(rs[2:]) // 2]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
