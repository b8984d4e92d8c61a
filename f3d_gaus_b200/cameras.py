"""Camera conventions of F3D-Gaus, restated (all matrices in the reference's row-vector
convention, i.e. p_view = [p, 1] @ world_view_transform, stored row-major):

  orbit sampling        src/utils.py:64-91            (sample_front_circle_gs, trajectory 'front_circle')
  cam2world             src/camera.py:17-32,65-92     (spherical2cartesian, compute_cam2world_matrix)
  projection            src/dataio_gs_test_256_demo.py:237-260 (getProjectionMatrix; note P[2,2]=(n+f)/(f-n))
  canonical re-basing   src/dataio_gs_test_256_demo.py:300-352 (update_camera_pose), visualize.py:241-273

With `update_pose: true` the world frame is the first (canonical) camera's frame, so the
canonical world_view_transform is the identity and its camera centre is the origin.
"""
from __future__ import annotations

import math
from typing import NamedTuple

import numpy as np
import torch


class Cameras(NamedTuple):
    world_view: torch.Tensor      # [V,4,4]
    view_to_world: torch.Tensor   # [V,4,4]
    full_proj: torch.Tensor       # [V,4,4]
    centers: torch.Tensor         # [V,3]


def projection_matrix(z_near: float, z_far: float, fov_deg: float) -> torch.Tensor:
    """Row-vector-convention projection (already transposed, as the dataset stores it)."""
    t = math.tan(fov_deg * 2 * np.pi / 360 / 2)
    top, right = t * z_near, t * z_near
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * z_near / (2 * right)
    P[1, 1] = 2.0 * z_near / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = (z_near + z_far) / (z_far - z_near)
    P[2, 3] = -(z_far * z_near) / (z_far - z_near)
    return P.transpose(0, 1).contiguous()


def _unit(v: torch.Tensor) -> torch.Tensor:
    return v / torch.norm(v, dim=-1, keepdim=True)


def look_at_pose(yaw: torch.Tensor, pitch: torch.Tensor, radius: float, look_at: torch.Tensor) -> torch.Tensor:
    """[V,4,4] pose matrix M = T(origin) @ [-left | up | -forward] (compute_cam2world_matrix)."""
    x = -radius * torch.sin(yaw) * torch.cos(pitch) + look_at[0]
    y = -radius * torch.sin(pitch) + look_at[1]
    z = -radius * torch.cos(pitch) * torch.cos(yaw) + look_at[2]
    origins = torch.stack([x, y, z], dim=-1)
    forward = _unit(_unit(look_at[None, :] - origins))
    up0 = torch.tensor([0.0, 1.0, 0.0]).expand_as(forward)
    left = _unit(torch.cross(up0, forward, dim=-1))
    up = _unit(torch.cross(forward, left, dim=-1))
    V = yaw.shape[0]
    rot = torch.eye(4).unsqueeze(0).repeat(V, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, up, -forward), dim=-1)
    trans = torch.eye(4).unsqueeze(0).repeat(V, 1, 1)
    trans[:, :3, 3] = origins
    return trans @ rot


def orbit_angles(num_frames: int, yaw_diff: float = 0.25, pitch_diff: float = 0.15):
    steps = torch.linspace(0, 1, num_frames)
    yaw = 0.0 - yaw_diff * torch.sin(steps * 2 * np.pi)
    pitch = 0.0 + pitch_diff * torch.cos(steps * 2 * np.pi)
    return yaw, pitch


def make_cameras(yaw: torch.Tensor, pitch: torch.Tensor, *, fov_deg: float = 13.164, radius: float = 7.667,
                 look_at_z: float = 7.667, z_near: float = 6.667, z_far: float = 8.667,
                 rebase: bool = True) -> Cameras:
    """Cameras at (yaw, pitch) on the sphere around the look-at point, re-based so that the
    (yaw=0, pitch=0) camera is the world frame (visualize.py:241-273)."""
    look_at = torch.tensor([0.0, 0.0, look_at_z])
    M = look_at_pose(yaw, pitch, radius, look_at)
    cam2w = torch.inverse(M)                 # the reference's (confusingly named) `cam2w`
    Rt = torch.inverse(cam2w).contiguous()
    world_view = Rt.transpose(1, 2).contiguous()
    view_to_world = cam2w.transpose(1, 2).contiguous()
    P = projection_matrix(z_near, z_far, fov_deg)
    full_proj = world_view.bmm(P.unsqueeze(0).expand(world_view.shape[0], -1, -1))
    if rebase:
        M0 = look_at_pose(torch.zeros(1), torch.zeros(1), radius, look_at)
        wv0 = torch.inverse(torch.inverse(M0)).transpose(1, 2)[0]
        F = wv0.inverse()
        Finv = F.inverse()
        world_view = torch.stack([F @ w for w in world_view])
        view_to_world = torch.stack([v @ Finv for v in view_to_world])
        full_proj = torch.stack([F @ f for f in full_proj])
    centers = torch.stack([w.inverse()[3, :3] for w in world_view])
    return Cameras(world_view.contiguous(), view_to_world.contiguous(), full_proj.contiguous(), centers.contiguous())


def canonical_camera(**kw) -> Cameras:
    return make_cameras(torch.zeros(1), torch.zeros(1), **kw)


def orbit_cameras(num_frames: int = 8, yaw_diff: float = 0.25, pitch_diff: float = 0.15, **kw) -> Cameras:
    yaw, pitch = orbit_angles(num_frames, yaw_diff, pitch_diff)
    return make_cameras(yaw, pitch, **kw)


def matrix_to_quaternion(M: torch.Tensor) -> torch.Tensor:
    """(r,x,y,z) of a rotation matrix (standard, positive-r branch first)."""
    m = M.double()
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = torch.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = torch.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = torch.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = torch.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    return torch.stack([torch.as_tensor(v) for v in q]).float()
