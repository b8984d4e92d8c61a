"""Gaussian-set export helpers: the flat attribute arrays of `load_ply` (visualize.py:146-179)
and a binary PLY writer/reader in the standard 3DGS attribute order (x y z nx ny nz f_dc_* f_rest_* opacity scale_*
rot_*).  Host-side I/O, numpy only (the reference goes through `plyfile`, and its `path is not None` branch references
an undefined `el`; the `path=None` branch -- the one visualize.py uses -- is what `flat_attributes` reproduces)."""
from __future__ import annotations

import os

import numpy as np
import torch


def flat_attributes(gs_dic: dict, bb: int):
    """(xyz[P,3], f_dc[P,3], f_rest[P,45], opacities[P,1], scale[P,3], rotation[P,4]) of scene `bb`, as
    visualize.py:165-179 returns them: features_dc transposed to channel-major and flattened, f_rest all zeros with the
    degree-3 width (15 coefficients x 3), everything detached, on the set's device."""
    xyz = gs_dic["xyz"][bb].detach()
    dc = gs_dic["features_dc"][bb].detach()
    f_dc = dc.transpose(1, 2).flatten(start_dim=1).contiguous()
    f_rest = torch.zeros_like(dc).expand([-1, (3 + 1) ** 2 - 1, -1]).transpose(1, 2).flatten(start_dim=1).contiguous()
    return (xyz, f_dc, f_rest, gs_dic["opacity"][bb].detach(), gs_dic["scaling"][bb].detach(),
            gs_dic["rotation"][bb].detach())


def attribute_names(n_dc: int = 3, n_rest: int = 45, n_scale: int = 3, n_rot: int = 4) -> list[str]:
    """construct_list_of_attributes (visualize.py:147-160)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)] + [f"f_rest_{i}" for i in range(n_rest)] + ["opacity"]
    return names + [f"scale_{i}" for i in range(n_scale)] + [f"rot_{i}" for i in range(n_rot)]


def save_ply(gs_dic: dict, bb: int, path: str) -> int:
    """Write scene `bb` as binary little-endian PLY (one float32 property per attribute, zero normals).  Returns P."""
    xyz, f_dc, f_rest, opac, scale, rot = [t.cpu().numpy().astype(np.float32) for t in flat_attributes(gs_dic, bb)]
    names = attribute_names(f_dc.shape[1], f_rest.shape[1], scale.shape[1], rot.shape[1])
    table = np.concatenate([xyz, np.zeros_like(xyz), f_dc, f_rest, opac.reshape(len(xyz), -1), scale, rot], axis=1)
    assert table.shape[1] == len(names)
    if os.path.dirname(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(xyz)
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(np.ascontiguousarray(table, dtype="<f4").tobytes())
    return len(xyz)


def read_ply(path: str) -> dict:
    """Read back a file written by `save_ply` (float32 vertex properties only): {name: array[P]}."""
    with open(path, "rb") as f:
        names, count = [], 0
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                count = int(line.split()[-1])
            elif line.startswith("property"):
                kind, name = line.split()[1:3]
                if kind not in ("float", "float32"):
                    raise ValueError(f"unsupported property type {kind}")
                names.append(name)
            elif line.startswith("format") and "binary_little_endian" not in line:
                raise ValueError("only binary_little_endian PLY files")
            elif line == "end_header":
                break
        table = np.frombuffer(f.read(count * len(names) * 4), dtype="<f4").reshape(count, len(names))
    return {n: table[:, i] for i, n in enumerate(names)}
