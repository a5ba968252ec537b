"""CPU: stage-boundary formats: cameras.json entries follow
TS/systems/GaussianDreamer.py:353-364 and survive the consumer's parsing
(Garment_Deformer_NeTF/deformer/core/view.py:55-95); RGBA PNGs follow threestudio/utils/saving.py:331-354."""
import json
import math

import numpy as np
import torch

from garmentdreamer_b200 import export
from garmentdreamer_b200.synthetic import sample_batch


def test_cameras_json_fields_and_consumer_parsing(tmp_path):
    batch = sample_batch(3, 1024, 1024)
    infos = [export.camera_info(batch["c2w_3dgs"][i], i, 1024, 1024, batch["fovy"][i]) for i in range(3)]
    export.save_cameras_json(str(tmp_path / "cameras.json"), infos)
    back = json.load(open(tmp_path / "cameras.json"))
    assert [set(e) for e in back] == [{"id", "img_name", "width", "height", "position", "rotation", "fy", "fx"}] * 3
    for i, e in enumerate(sorted(back, key=lambda x: x["id"])):
        c2w = batch["c2w_3dgs"][i].double().numpy()
        assert e["img_name"] == str(i) and e["width"] == 1024 and e["height"] == 1024
        assert np.allclose(e["position"], c2w[:3, 3]) and np.allclose(e["rotation"], -c2w[:3, :3])
        fy = 1024 / (2 * math.tan(float(batch["fovy"][i]) / 2))
        assert math.isclose(e["fy"], fy, rel_tol=1e-9) and math.isclose(e["fx"], fy, rel_tol=1e-6)   # square image: fx == fy
        # the consumer's transformation (view.py:62-85) yields a proper rigid world-to-camera matrix
        position, rotation = np.array(e["position"]), np.array(e["rotation"])
        rotation[:, 0] *= -1
        position[1] = -position[1]
        rotation[1, 0] = -rotation[1, 0]
        rotation[1, 2] = -rotation[1, 2]
        rotation[:, 1] = np.cross(rotation[:, 2], rotation[:, 0])
        rotation[:, 1] /= np.linalg.norm(rotation[:, 1])
        rotation[:, 2] *= -1
        C2W = np.eye(4); C2W[:3, :3] = rotation; C2W[:3, 3] = position
        R = np.linalg.inv(C2W)[:3, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-5) and abs(abs(np.linalg.det(R)) - 1) < 1e-5


def test_rgba_png_round_trip_and_rounding(tmp_path):
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(37, 53, 3, generator=g) * 1.2 - 0.1          # out-of-range values are clipped
    alpha = torch.rand(37, 53, generator=g)
    mask = alpha >= 0.8
    p = export.save_image_rgba(str(tmp_path / "gs_rendered_rgba" / "7.png"), rgb, mask)
    img = export.load_image_rgba(p)
    assert img.shape == (37, 53, 4) and img.dtype == np.uint8
    assert np.array_equal(img[..., :3], np.rint(rgb.clamp(0, 1).numpy() * 255.0).astype(np.uint8))
    assert set(np.unique(img[..., 3]).tolist()) <= {0, 255} and np.array_equal(img[..., 3] == 255, mask.numpy())
    try:   # a standard PNG: any decoder reads it
        from PIL import Image
        assert np.array_equal(np.array(Image.open(p)), img)
    except ImportError:
        pass
