"""CPU: host-side logic that needs no GPU -- learning-rate schedule and optimiser hyper-parameters of
the GaussianModel mirror, camera intrinsics, and the bench.py JSON contract (checked on the
committed line of the last measured run)."""
import json
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_expon_lr_schedule_matches_reference_formula():
    """gaussiansplatting/utils/general_utils.py:29-62 with the values of arguments/__init__.py:73-76."""
    from garmentdreamer_b200.gaussians import OptimizationParams, get_expon_lr_func
    a = OptimizationParams()
    f = get_expon_lr_func(a.position_lr_init * 1.75, a.position_lr_final * 1.75, lr_delay_mult=a.position_lr_delay_mult,
                          max_steps=a.position_lr_max_steps)
    assert math.isclose(f(0), a.position_lr_init * 1.75, rel_tol=1e-12)
    assert math.isclose(f(a.position_lr_max_steps), a.position_lr_final * 1.75, rel_tol=1e-12)
    assert math.isclose(f(10 ** 9), a.position_lr_final * 1.75, rel_tol=1e-12)          # clipped
    mid = f(a.position_lr_max_steps // 2)
    assert math.isclose(mid, 1.75 * math.sqrt(a.position_lr_init * a.position_lr_final), rel_tol=1e-9)   # log-linear
    assert f(-1) == 0.0 and get_expon_lr_func(0.0, 0.0)(5) == 0.0
    g = get_expon_lr_func(1e-2, 1e-3, lr_delay_steps=100, lr_delay_mult=0.1, max_steps=1000)
    assert math.isclose(g(0), 1e-3, rel_tol=1e-9) and g(50) < 1e-2 * np.exp(np.log(0.1) * 0.05)
    assert (a.feature_lr, a.opacity_lr, a.scaling_lr, a.rotation_lr, a.percent_dense) == (0.0125, 0.01, 0.005, 0.001, 0.01)


def test_camera_intrinsics_match_reference_formulas():
    """cameras.py:24 (FoVx from FoVy through the focal length) and graphics_utils.py:98-101."""
    from garmentdreamer_b200.cameras import focal2fov, fov2focal
    fovy = math.radians(50.0)
    for h, w in ((512, 512), (512, 384), (1024, 768)):
        fovx = focal2fov(fov2focal(fovy, h), w)
        assert math.isclose(math.tan(fovx / 2), w * math.tan(fovy / 2) / h, rel_tol=1e-12)
    assert math.isclose(fov2focal(math.pi / 2, 100), 50.0, rel_tol=1e-12)


def test_bench_line_contract_on_committed_run():
    """The JSON line of the last measured run (profiles/) carries every key the contract names."""
    path = os.path.join(ROOT, "profiles", "r02_b2_bench.json")
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["scaling"] == "weak" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and math.isclose(r["frac"], r["achieved"] / r["peak"], rel_tol=1e-9)
    assert r["raster_bwd"]["bound"] == "hbm" and r["vae"]["bound"] == "tensor"
    assert any(k["traffic"] for k in r["dominant_kernels"])
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert math.isclose(d["value"], d["n_gpus"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3), rel_tol=1e-6)
    # round 2: the step is the reference's whole iteration, with a same-run GPU reference block
    assert set(d["phase_ms"]) == {"raster_fwd", "sparsity", "guidance", "raster_bwd", "allreduce", "adam"}
    assert set(d["gpu_reference"]) >= {"raster_fwd_ms", "raster_bwd_ms", "unet_fp16_eager_ms", "vae_fp16_eager_autograd_ms", "ms_per_step"}
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02_b2_bench_reference.json")))
    assert ref["impl"] == "reference" and ref["extrapolated"] is True and ref["config"]["workload"] == d["config"]["workload"]
    assert ref["cpu_baseline"]["kind"] == "port" and ref["e2e"]["h2d_bytes_per_step"] == 0
