"""CPU: last_3dgs.ply round trip in the reference's layout (gaussian_model.py:187-262): header
property order, float32 little-endian table, channel-major f_rest, raw (pre-activation) values."""
import numpy as np

from garmentdreamer_b200.gaussians import load_ply, ply_attribute_names, save_ply


def test_ply_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    P = 1234
    xyz, dc = rng.normal(size=(P, 3)).astype(np.float32), rng.normal(size=(P, 1, 3)).astype(np.float32)
    rest = rng.normal(size=(P, 3, 3)).astype(np.float32)   # sh_degree 1: 3 extra coefficients x 3 channels
    op, sc, rot = rng.normal(size=(P, 1)).astype(np.float32), rng.normal(size=(P, 3)).astype(np.float32), rng.normal(size=(P, 4)).astype(np.float32)
    path = str(tmp_path / "save" / "last_3dgs.ply")
    save_ply(path, xyz, dc, op, sc, rot, f_rest=rest)
    raw = open(path, "rb").read()
    head = raw[:raw.index(b"end_header\n")].decode().split("\n")
    assert head[0] == "ply" and head[1] == "format binary_little_endian 1.0" and head[2] == f"element vertex {P}"
    assert [h.split()[2] for h in head[3:] if h] == ply_attribute_names(9)
    assert ply_attribute_names(0) == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity",
                                      "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    body = np.frombuffer(raw, "<f4", offset=raw.index(b"end_header\n") + 11).reshape(P, -1)
    assert body.shape[1] == 26 and np.array_equal(body[:, 3:6], np.zeros((P, 3), np.float32))     # normals are zeros
    assert np.array_equal(body[:, 9:18].reshape(P, 3, 3), rest.transpose(0, 2, 1))                # channel-major f_rest
    back = load_ply(path)
    for k, v in dict(xyz=xyz, f_dc=dc, f_rest=rest, opacity=op, scaling=sc, rotation=rot).items():
        assert back[k].dtype == np.float32 and np.array_equal(back[k], v), k


def test_ply_degree0_and_empty(tmp_path):
    path = str(tmp_path / "a.ply")
    z = lambda *s: np.zeros(s, np.float32)
    save_ply(path, z(0, 3), z(0, 1, 3), z(0, 1), z(0, 3), z(0, 4))
    back = load_ply(path)
    assert back["xyz"].shape == (0, 3) and back["f_rest"].shape == (0, 0, 3) and back["rotation"].shape == (0, 4)
