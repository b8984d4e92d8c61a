"""Seeded synthetic Gaussian clouds shaped like what the F3D-Gaus predictor emits
(SURVEY.md 8d): one Gaussian per pixel of an SxS source view, un-projected along the
predictor's ray grid (src/gaussian_predictor.py:657-670, y inverted) at depths in
[z_near, z_far], scale ~ 0.01, SH degree 1.  A second, well-conditioned "unit-cloud" separates
logic errors from the float32 conditioning of the GOF quadric.

Everything is generated on the CPU with a torch.Generator so that the same seed gives the same
cloud everywhere; callers move the dict to the GPU.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

CFG_256 = {"model": {"fov": 13.164, "training_resolution": 256, "max_sh_degree": 1, "radius": 7.667,
                     "look_at": 7.667},
           "dataset_params": {"z_near": 6.667, "z_far": 8.667}}


def cfg_for(resolution: int) -> dict:
    cfg = {k: dict(v) for k, v in CFG_256.items()}
    cfg["model"]["training_resolution"] = int(resolution)
    return cfg


def f3d_like(seed: int, S: int = 256, *, fov_deg: float = 13.164, z_near: float = 6.667, z_far: float = 8.667,
             log_scale_mean: float = math.log(0.01), log_scale_std: float = 0.3, view_to_world: torch.Tensor | None = None,
             quat: torch.Tensor | None = None) -> dict:
    """P = S*S Gaussians.  Keys/shapes follow the predictor's output dict with a batch dim of 1:
    xyz[1,P,3] opacity[1,P,1] scaling[1,P,3] rotation[1,P,4] features_dc[1,P,1,3] features_rest[1,P,3,3].
    If `view_to_world` ([4,4], row-vector convention) is given the cloud is expressed in the world
    frame through that source view (the re-predicted sets of the cycle-aggregative loop)."""
    g = torch.Generator().manual_seed(int(seed))
    P = S * S
    f_S = S / (2 * math.tan(math.radians(fov_deg) / 2))
    i = torch.arange(S, dtype=torch.float32)
    x = (i + 0.5 - S / 2) / f_S
    y = -(i + 0.5 - S / 2) / f_S
    ray = torch.stack([x[None, :].expand(S, S), y[:, None].expand(S, S), torch.ones(S, S)], dim=-1)
    u = torch.rand(1, 1, 16, 16, generator=g)
    u = F.interpolate(u, size=(S, S), mode="bilinear", align_corners=False)[0, 0].clamp(0, 1)
    depth = z_near + (z_far - z_near) * u
    xyz = (ray * depth[..., None]).reshape(P, 3)
    scaling = torch.exp(log_scale_mean + log_scale_std * torch.randn(P, 3, generator=g))
    rotation = F.normalize(torch.randn(P, 4, generator=g), dim=-1)
    opacity = torch.sigmoid(2.0 * torch.randn(P, 1, generator=g))
    features_dc = (torch.rand(P, 1, 3, generator=g) - 0.5) / 0.28209479177387814
    features_rest = 0.1 * torch.randn(P, 3, 3, generator=g)
    if view_to_world is not None:
        v2w = view_to_world.float()
        xyz = torch.cat([xyz, torch.ones(P, 1)], dim=1) @ v2w
        xyz = xyz[:, :3].contiguous()
        if quat is not None:   # left-multiply rotations by the source camera's quaternion
            rotation = quat_multiply(quat[None, :].expand(P, 4), rotation)
    out = {"xyz": xyz, "opacity": opacity, "scaling": scaling, "rotation": rotation, "features_dc": features_dc,
           "features_rest": features_rest}
    return {k: v.unsqueeze(0).contiguous() for k, v in out.items()}


def unit_cloud(seed: int, P: int = 4096, *, sh_degree: int = 1) -> dict:
    """Well-conditioned cloud: xyz ~ U([-1,1]^2 x [2,4]), scales exp(N(log 0.05, 0.5^2)); render at fov 60."""
    g = torch.Generator().manual_seed(int(seed))
    xyz = torch.rand(P, 3, generator=g)
    xyz = torch.stack([xyz[:, 0] * 2 - 1, xyz[:, 1] * 2 - 1, xyz[:, 2] * 2 + 2], dim=-1)
    M = (sh_degree + 1) ** 2
    out = {
        "xyz": xyz,
        "scaling": torch.exp(math.log(0.05) + 0.5 * torch.randn(P, 3, generator=g)),
        "rotation": F.normalize(torch.randn(P, 4, generator=g), dim=-1),
        "opacity": torch.sigmoid(2.0 * torch.randn(P, 1, generator=g)),
        "features_dc": (torch.rand(P, 1, 3, generator=g) - 0.5) / 0.28209479177387814,
        "features_rest": 0.1 * torch.randn(P, M - 1, 3, generator=g),
    }
    return {k: v.unsqueeze(0).contiguous() for k, v in out.items()}


def quat_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product a*b for (r,x,y,z) quaternions (src/gaussian_predictor.py:839-855)."""
    ar, ax, ay, az = a.unbind(-1)
    br, bx, by, bz = b.unbind(-1)
    return torch.stack([ar * br - ax * bx - ay * by - az * bz,
                        ar * bx + ax * br + ay * bz - az * by,
                        ar * by - ax * bz + ay * br + az * bx,
                        ar * bz + ax * by - ay * bx + az * br], dim=-1)


def concat_sets(sets: list[dict]) -> dict:
    """multi-view union (src/gaussian_predictor.py:796-800 / visualize.py:336-340): concat on dim 1."""
    return {k: torch.cat([s[k] for s in sets], dim=1).contiguous() for k in sets[0]}


def to_device(pc: dict, device) -> dict:
    return {k: v.to(device) for k, v in pc.items()}


def perspective_camera(fov_deg: float, z_near: float = 0.1, z_far: float = 100.0):
    """Identity view + the reference-style projection at `fov_deg` (for the unit cloud)."""
    from .cameras import projection_matrix
    wv = torch.eye(4)
    return wv, projection_matrix(z_near, z_far, fov_deg), torch.zeros(3)
