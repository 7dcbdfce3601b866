"""Gaussian-splat PLY reader / writer — the on-disk format between GaussianIP's stages.

Mirrors ``GaussianModel.save_ply`` / ``load_ply`` (gaussiansplatting/scene/gaussian_model.py:185-264) and the
animation-path loader with its axis swap (gs_renderer.py:525-602), without the ``plyfile`` dependency:

* attribute order ``x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*`` (construct_list_of_attributes :185-197),
  all ``float`` (``f4``), normals zero, SH coefficients stored channel-major (``transpose(1, 2).flatten``);
* file layout as plyfile writes ``PlyData([PlyElement.describe(elements, 'vertex')])``: ASCII header
  ``ply / format binary_little_endian 1.0 / element vertex N / property float <name> ... / end_header`` followed by
  N packed little-endian records.

Host code (numpy): file IO is not GPU work.  The reader accepts binary little/big endian and ascii files whose
vertex element holds scalar properties (what 3DGS tools write)."""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np
import torch
from torch import nn

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def attribute_names(n_dc: int, n_rest: int, n_scale: int = 3, n_rot: int = 4) -> List[str]:
    """gaussian_model.py:185-197."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def pack_attributes(xyz, features_dc, features_rest, opacity, scaling, rotation) -> Tuple[np.ndarray, List[str]]:
    """[P, n_attr] float32 table in the file's column order (gaussian_model.py:202-214)."""
    t = lambda a: a.detach().cpu() if torch.is_tensor(a) else torch.as_tensor(a)
    xyz = t(xyz).float().numpy()
    f_dc = t(features_dc).float().transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    f_rest = t(features_rest).float().transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    table = np.concatenate((xyz, np.zeros_like(xyz), f_dc, f_rest, t(opacity).float().numpy().reshape(len(xyz), -1),
                            t(scaling).float().numpy(), t(rotation).float().numpy()), axis=1).astype("<f4")
    names = attribute_names(f_dc.shape[1], f_rest.shape[1], t(scaling).shape[1], t(rotation).shape[1])
    assert table.shape[1] == len(names)
    return table, names


def write_table(path: str, table: np.ndarray, names: List[str]) -> None:
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)       # mkdir_p, gaussian_model.py:200
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {table.shape[0]}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(np.ascontiguousarray(table, dtype="<f4").tobytes())


def save_ply(model, path: str) -> None:
    """Drop-in for GaussianModel.save_ply (gaussian_model.py:199-214)."""
    table, names = pack_attributes(model._xyz, model._features_dc, model._features_rest, model._opacity,
                                   model._scaling, model._rotation)
    write_table(path, table, names)


def read_vertices(path: str) -> Dict[str, np.ndarray]:
    """name -> column of the first element (plydata.elements[0][name])."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, props, count, in_first, n_elements = None, [], None, False, 0
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: header without end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                n_elements += 1
                in_first = n_elements == 1
                if in_first:
                    count = int(tok[2])
            elif tok[0] == "property" and in_first:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties in the first element are not supported")
                if tok[1] not in _PLY_TYPES:
                    raise ValueError(f"{path}: unknown property type {tok[1]}")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt is None or count is None:
            raise ValueError(f"{path}: incomplete header")
        if fmt == "ascii":
            rows = [f.readline().split() for _ in range(count)]
            if any(len(r) < len(props) for r in rows):
                raise ValueError(f"{path}: truncated vertex data")
            return {name: np.array([r[i] for r in rows], dtype=np.float64).astype(ty)
                    for i, (name, ty) in enumerate(props)}
        order = {"binary_little_endian": "<", "binary_big_endian": ">"}.get(fmt)
        if order is None:
            raise ValueError(f"{path}: unknown format {fmt}")
        dtype = np.dtype([(name, order + ty) for name, ty in props])
        raw = f.read(count * dtype.itemsize)
        if len(raw) != count * dtype.itemsize:
            raise ValueError(f"{path}: truncated vertex data")
        rec = np.frombuffer(raw, dtype=dtype, count=count)
        return {name: np.asarray(rec[name]) for name, _ in props}


def _sorted_cols(cols: Dict[str, np.ndarray], prefix: str, sort: bool) -> List[str]:
    names = [n for n in cols if n.startswith(prefix)]
    return sorted(names, key=lambda x: int(x.split("_")[-1])) if sort else names


def load_arrays(path: str, max_sh_degree: int, swap_axes: bool = False) -> Dict[str, np.ndarray]:
    """The numpy stage of load_ply (gaussian_model.py:223-255); swap_axes applies the animation loader's
    'coordinate shift' (gs_renderer.py:576-581: y<->z of xyz and scales, rot_2<->rot_3, rot_0 negated)."""
    cols = read_vertices(path)
    xyz = np.stack((cols["x"], cols["y"], cols["z"]), axis=1).astype(np.float64)
    opacities = np.asarray(cols["opacity"], dtype=np.float64)[..., np.newaxis]
    features_dc = np.zeros((xyz.shape[0], 3, 1))
    for c in range(3):
        features_dc[:, c, 0] = cols[f"f_dc_{c}"]
    extra = _sorted_cols(cols, "f_rest_", sort=not swap_axes)      # the animation loader keeps file order
    assert len(extra) == 3 * (max_sh_degree + 1) ** 2 - 3, "PLY SH degree does not match max_sh_degree"
    features_extra = np.zeros((xyz.shape[0], len(extra)))
    for i, n in enumerate(extra):
        features_extra[:, i] = cols[n]
    features_extra = features_extra.reshape((xyz.shape[0], 3, (max_sh_degree + 1) ** 2 - 1))
    scale_names = _sorted_cols(cols, "scale_", sort=not swap_axes)
    scales = np.stack([cols[n] for n in scale_names], axis=1).astype(np.float64)
    rot_names = _sorted_cols(cols, "rot", sort=not swap_axes)
    rots = np.stack([cols[n] for n in rot_names], axis=1).astype(np.float64)
    if swap_axes:
        xyz[:, [1, 2]] = xyz[:, [2, 1]]
        scales[:, [1, 2]] = scales[:, [2, 1]]
        rots[:, [2, 3]] = rots[:, [3, 2]]
        rots[:, [0]] *= -1
    return {"xyz": xyz, "features_dc": features_dc, "features_extra": features_extra, "opacities": opacities,
            "scales": scales, "rots": rots}


def load_ply(model, path: str, device="cuda", swap_axes: bool = False) -> None:
    """Drop-in for GaussianModel.load_ply (gaussian_model.py:223-264): fills the six parameters and sets
    active_sh_degree = max_sh_degree."""
    a = load_arrays(path, model.max_sh_degree, swap_axes)
    par = lambda x: nn.Parameter(x.requires_grad_(True))
    tt = lambda x: torch.tensor(x, dtype=torch.float, device=device)
    model._xyz = par(tt(a["xyz"]))
    model._features_dc = par(tt(a["features_dc"]).transpose(1, 2).contiguous())
    model._features_rest = par(tt(a["features_extra"]).transpose(1, 2).contiguous())
    model._opacity = par(tt(a["opacities"]))
    model._scaling = par(tt(a["scales"]))
    model._rotation = par(tt(a["rots"]))
    model.active_sh_degree = model.max_sh_degree
