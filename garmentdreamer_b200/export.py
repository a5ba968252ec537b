"""On-disk formats at the stage boundary, host side only.

What the 3DGS stage hands to Normal_estimator_Metric3D / Garment_Deformer_NeTF besides last_3dgs.ply
(gaussians.save_ply): one RGBA PNG per test view and cameras.json, written by the reference's
``test_step`` / ``on_test_epoch_end`` (Garment_3DGS/threestudio/systems/GaussianDreamer.py:334-417,
threestudio/utils/saving.py:331-354) and read back by Garment_Deformer_NeTF/deformer/core/view.py:55-95.

  * cameras.json: a list of {"id", "img_name", "width", "height", "position", "rotation", "fy", "fx"} with
    position = c2w[:3,3], rotation = -c2w[:3,:3] (the reference negates the rotation in place, :358-361),
    fy = fov2focal(fovy, height), fx = fov2focal(focal2fov(fy, width), width);
  * gs_rendered_rgba/<index>.png: 8-bit RGBA, rgb = round(clip(comp_rgb, 0, 1) * 255) rendered on a white
    background, alpha = 255 where the accumulated alpha >= alpha_threshold (0.8) else 0.

The PNG codec is a minimal zlib one (the image has no cv2); files are standard PNGs.
"""
import json
import math
import os
import struct
import zlib

import numpy as np
import torch


def fov2focal(fov, pixels):
    return pixels / (2 * math.tan(fov / 2))


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def camera_info(c2w, index, width, height, fovy):
    """One cameras.json entry (GaussianDreamer.py:353-364). c2w: [4,4] (batch['c2w'][0])."""
    C2W = np.array(torch.as_tensor(c2w).detach().cpu().numpy(), dtype=np.float64)
    pos = C2W[:3, 3]
    rot = -C2W[:3, :3]
    fovy = float(fovy)
    fy = fov2focal(fovy, height)
    fx = fov2focal(focal2fov(fov2focal(fovy, height), width), width)
    return {"id": int(index), "img_name": str(int(index)), "width": int(width), "height": int(height),
            "position": pos.tolist(), "rotation": [r.tolist() for r in rot], "fy": float(fy), "fx": float(fx)}


def save_cameras_json(path, camera_info_list):
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "w") as f:
        json.dump(camera_info_list, f)


def _png_chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def save_image_rgba(path, rgb, mask):
    """rgb [H,W,3] float in [0,1] (clipped), mask [H,W] in {0,1} (bool or float, max <= 1) -> 8-bit RGBA PNG."""
    rgb = torch.as_tensor(rgb).detach().float().cpu().numpy()
    mask = torch.as_tensor(mask).detach().float().cpu().numpy()
    assert mask.max() <= 1.0 and rgb.ndim == 3 and rgb.shape[2] == 3 and mask.shape == rgb.shape[:2]
    img = np.concatenate([rgb, mask[..., None]], -1).clip(0.0, 1.0) * 255.0
    img = np.rint(img).astype(np.uint8)                      # cv2.imwrite saturate_cast<uchar>(float): round to nearest
    H, W = img.shape[:2]
    raw = np.concatenate([np.zeros((H, 1), np.uint8), img.reshape(H, W * 4)], 1).tobytes()   # filter type 0 per scanline
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(_png_chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 6, 0, 0, 0)))
        f.write(_png_chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(_png_chunk(b"IEND", b""))
    return path


def load_image_rgba(path):
    """Reads back an 8-bit RGBA PNG written with filter type 0 (save_image_rgba) -> uint8 [H,W,4]."""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, W, H = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            W, H, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert depth == 8 and ctype == 6
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(H, 1 + 4 * W)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(H, W, 4).copy()


def test_step(system, batch, alpha_threshold=0.8, save_dir=None):
    """Mirror of GaussianDreamer.test_step (:338-411): one view on a white background, mask = alpha >= threshold,
    the cameras.json entry; writes gs_rendered_rgba/<index>.png under save_dir when given.
    batch: c2w_3dgs [1,4,4] (render pose), c2w [1,4,4] (exported pose), fovy [1], index [1], height, width."""
    bg = torch.tensor([1.0, 1.0, 1.0], dtype=torch.float32, device=system.dev)
    out = system.forward(batch, bg)
    alpha = out["alphas"].squeeze()
    rgb = out["comp_rgb"].squeeze()
    mask = alpha >= alpha_threshold
    idx = int(torch.as_tensor(batch["index"]).reshape(-1)[0])
    info = camera_info(torch.as_tensor(batch["c2w"])[0], idx, batch["width"], batch["height"], float(torch.as_tensor(batch["fovy"]).reshape(-1)[0]))
    if save_dir is not None:
        save_image_rgba(os.path.join(save_dir, "gs_rendered_rgba", f"{idx}.png"), rgb, mask)
    return {"rgb": rgb, "mask": mask, "camera_info": info}
