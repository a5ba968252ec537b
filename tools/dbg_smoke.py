import sys, torch
sys.path.insert(0, "/root/repo")
from garmentdreamer_b200 import sds_step
from garmentdreamer_b200.gaussians import GaussianParams
from garmentdreamer_b200.synthetic import garment, raw_params, sample_batch
from garmentdreamer_b200.system import GaussianDreamerB200
dev = torch.device("cuda:0")
guide = sds_step.make_bench_guidance(dev, 2, use_vae=True)
raw = {k: v.to(dev) for k, v in raw_params(garment(5000, 0)).items()}
gp = GaussianParams(raw["xyz"], raw["f_dc"], raw["opacity"], raw["scaling"], raw["rotation"], spatial_lr_scale=4.0)
gp.training_setup()
system = GaussianDreamerB200(gp, guide)
batch = sample_batch(2, 512, 512)
orig = guide.image_grad
def wrapped(color, *a, **k):
    d = orig(color, *a, **k)
    print("color finite", bool(torch.isfinite(color).all()), "min/max", float(color.min()), float(color.max()),
          "| sds grad finite", bool(torch.isfinite(guide.last_grad).all()), float(guide.last_grad.abs().max()),
          "| dcol finite", bool(torch.isfinite(d).all()), float(torch.nan_to_num(d).abs().max()))
    return d
guide.image_grad = wrapped
for i in range(3):
    out = system.training_step(batch)
    torch.cuda.synchronize()
    g = system.grad
    P = gp.P
    names = ["xyz","f_dc","op","sc","rot","m2d"]
    offs = [0,3*P,6*P,7*P,10*P,14*P,17*P]
    print(i, {n: (bool(torch.isfinite(g[offs[k]:offs[k+1]]).all()), float(torch.nan_to_num(g[offs[k]:offs[k+1]]).abs().max())) for k,n in enumerate(names)})
    print("  params finite", all(bool(torch.isfinite(t).all()) for t in (gp._xyz, gp._features_dc, gp._opacity, gp._scaling, gp._rotation)))
