"""Seeded rasterizer test cases (flat argument dicts) shared by the GPU parity tests, the golden
generator and the CPU-oracle tests."""
from __future__ import annotations

import math

import torch

from f3d_gaus_b200 import cameras, synthetic


def make_case(pc: dict, world_view, full_proj, campos, *, W: int, H: int, fov_deg: float, sh_degree: int = 1,
              bg=(0.0, 0.0, 0.0), kernel_size: float = 0.0, scale_modifier: float = 1.0, use_colors: bool = False,
              device="cpu") -> dict:
    d = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
    tanfov = math.tan(fov_deg * math.pi / 360)
    c = {
        "W": int(W), "H": int(H), "D": int(sh_degree), "tanfovx": tanfov, "tanfovy": tanfov,
        "kernel_size": float(kernel_size), "scale_modifier": float(scale_modifier),
        "bg": d(torch.tensor(bg)), "means3D": d(pc["xyz"][0]), "opacities": d(pc["opacity"][0]),
        "scales": d(pc["scaling"][0]), "rotations": d(pc["rotation"][0]),
        "viewmatrix": d(world_view.reshape(4, 4)), "projmatrix": d(full_proj.reshape(4, 4)),
        "campos": d(campos.reshape(3)),
    }
    if use_colors:
        g = torch.Generator().manual_seed(1234)
        c["colors_precomp"] = d(torch.rand(pc["xyz"].shape[1], 3, generator=g))
        c["D"] = 0
    else:
        c["shs"] = d(torch.cat([pc["features_dc"][0], pc["features_rest"][0]], dim=1))
    return c


def f3d_case(seed: int, S: int, res: int, view: int | None, device="cpu", **kw) -> dict:
    """f3d-like cloud of S*S Gaussians rendered at res x res from the canonical camera
    (view=None) or novel orbit view `view` of 8."""
    pc = synthetic.f3d_like(seed, S)
    cams = cameras.canonical_camera() if view is None else cameras.orbit_cameras(8)
    k = 0 if view is None else view
    return make_case(pc, cams.world_view[k], cams.full_proj[k], cams.centers[k], W=res, H=res, fov_deg=13.164,
                     device=device, **kw)


def unit_case(seed: int, P: int, W: int, H: int, sh_degree: int = 1, device="cpu", **kw) -> dict:
    pc = synthetic.unit_cloud(seed, P, sh_degree=sh_degree)
    wv, proj, campos = synthetic.perspective_camera(60.0)
    return make_case(pc, wv, wv @ proj, campos, W=W, H=H, fov_deg=60.0, sh_degree=sh_degree, device=device, **kw)


def grad_seed(c: dict, seed: int = 7) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(9, c["H"], c["W"], generator=g).to(c["means3D"].device)


def case_to(c: dict, device) -> dict:
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in c.items()}


GOLDEN_CASES = {
    # name: builder (CPU tensors)
    "f3d_s48_r128_view2": lambda: f3d_case(0, 48, 128, 2),
    "unit_p1200_120x88_sh3": lambda: unit_case(1, 1200, 120, 88, sh_degree=3, bg=(0.2, 0.5, 0.9)),
    "f3d_s32_r96_colors_ks": lambda: f3d_case(2, 32, 96, None, use_colors=True, kernel_size=0.1,
                                              scale_modifier=1.3, bg=(1.0, 0.5, 0.25)),
}
