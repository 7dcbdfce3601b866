"""PLY reader / writer (gaussianip_b200/ply.py): layout of GaussianModel.save_ply (gaussian_model.py:185-214)
and the two loaders (gaussian_model.py:223-264, gs_renderer.py:525-602).  `plyfile` is not in the image, so the
file layout is checked byte for byte against its documented format instead of against the library."""
import struct
import types

import numpy as np
import pytest
import torch

from gaussianip_b200 import ply


def make_model(P=37, deg=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    K = (deg + 1) ** 2
    m = types.SimpleNamespace(max_sh_degree=deg, active_sh_degree=0)
    m._xyz = torch.randn(P, 3, generator=g)
    m._features_dc = torch.randn(P, 1, 3, generator=g)
    m._features_rest = torch.randn(P, K - 1, 3, generator=g)
    m._opacity = torch.randn(P, 1, generator=g)
    m._scaling = torch.randn(P, 3, generator=g)
    m._rotation = torch.randn(P, 4, generator=g)
    return m


def test_header_and_record_layout(tmp_path):
    m = make_model(P=5, deg=1)
    path = str(tmp_path / "sub" / "point_cloud.ply")          # directory is created, as mkdir_p does
    ply.save_ply(m, path)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().strip().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 5"]
    names = [l.split()[2] for l in lines[3:]]
    assert all(l.startswith("property float ") for l in lines[3:])
    assert names == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + \
        [f"f_rest_{i}" for i in range(9)] + ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert len(body) == 5 * len(names) * 4
    rec0 = struct.unpack("<" + "f" * len(names), body[:len(names) * 4])
    assert rec0[0:3] == tuple(m._xyz[0].tolist()) and rec0[3:6] == (0.0, 0.0, 0.0)
    assert rec0[6:9] == tuple(m._features_dc[0, 0].tolist())
    # f_rest is channel-major: all coefficients of R, then G, then B (transpose(1, 2).flatten)
    assert rec0[9:12] == tuple(m._features_rest[0, :, 0].tolist())
    assert rec0[18] == m._opacity[0, 0].item() and rec0[19:22] == tuple(m._scaling[0].tolist())
    assert rec0[22:26] == tuple(m._rotation[0].tolist())


@pytest.mark.parametrize("deg", [0, 1, 3])
def test_round_trip_is_exact(tmp_path, deg):
    m = make_model(P=101, deg=deg, seed=deg)
    path = str(tmp_path / "a.ply")
    ply.save_ply(m, path)
    r = types.SimpleNamespace(max_sh_degree=deg, active_sh_degree=0)
    ply.load_ply(r, path, device="cpu")
    for k in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation"):
        got, want = getattr(r, k), getattr(m, k)
        assert isinstance(got, torch.nn.Parameter) and got.requires_grad and got.is_contiguous()
        assert got.shape == want.shape and torch.equal(got.detach(), want), k
    assert r.active_sh_degree == deg
    with pytest.raises(AssertionError):
        ply.load_ply(types.SimpleNamespace(max_sh_degree=deg + 1), path, device="cpu")


def test_axis_swap_loader(tmp_path):
    m = make_model(P=9, deg=0)
    path = str(tmp_path / "a.ply")
    ply.save_ply(m, path)
    r = types.SimpleNamespace(max_sh_degree=0)
    ply.load_ply(r, path, device="cpu", swap_axes=True)
    assert torch.equal(r._xyz.detach(), m._xyz[:, [0, 2, 1]])
    assert torch.equal(r._scaling.detach(), m._scaling[:, [0, 2, 1]])
    want = m._rotation[:, [0, 1, 3, 2]].clone()
    want[:, 0] *= -1
    assert torch.equal(r._rotation.detach(), want)


def test_reads_ascii_big_endian_and_unsorted_columns(tmp_path):
    names = ["x", "y", "z", "opacity", "f_dc_0", "f_dc_1", "f_dc_2", "scale_2", "scale_0", "scale_1",
             "rot_0", "rot_1", "rot_2", "rot_3"]
    vals = np.arange(2 * len(names), dtype=np.float64).reshape(2, len(names)) * 0.5
    p1 = tmp_path / "ascii.ply"
    p1.write_text("ply\nformat ascii 1.0\ncomment made by a test\nelement vertex 2\n" +
                  "".join(f"property double {n}\n" for n in names) + "element face 0\nproperty list uchar int vertex_indices\n"
                  "end_header\n" + "\n".join(" ".join(repr(float(v)) for v in row) for row in vals) + "\n")
    p2 = tmp_path / "be.ply"
    with open(p2, "wb") as f:
        f.write(("ply\nformat binary_big_endian 1.0\nelement vertex 2\n" + "".join(f"property float {n}\n" for n in names)
                 + "end_header\n").encode())
        f.write(vals.astype(">f4").tobytes())
    for p in (p1, p2):
        a = ply.load_arrays(str(p), max_sh_degree=0)
        assert np.array_equal(a["xyz"], vals[:, 0:3]) and np.array_equal(a["opacities"][:, 0], vals[:, 3])
        assert np.array_equal(a["scales"], vals[:, [8, 9, 7]])      # sorted by index, not file order
        assert np.array_equal(a["rots"], vals[:, 10:14])
    with pytest.raises(ValueError):
        (tmp_path / "bad.ply").write_bytes(b"plx\n")
        ply.read_vertices(str(tmp_path / "bad.ply"))
    trunc = open(p2, "rb").read()[:-3]
    (tmp_path / "trunc.ply").write_bytes(trunc)
    with pytest.raises(ValueError):
        ply.read_vertices(str(tmp_path / "trunc.ply"))
