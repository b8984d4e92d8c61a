"""Torch restatement of the reference's L2 post-processing -- TEST INFRASTRUCTURE (the oracle of the fused epilogue
kernel, csrc/epilogue.cu).  Follows /root/reference/src/gaussian_renderer/__init__.py:

  depths_to_points   :881-896   c2w = (wvt^T)^-1; intrinsics [[fx,0,W/2],[0,fy,H/2],[0,0,1]] with fx = W/(2 tan(FoVx/2));
                                rays_d = [x, y, 1] @ K^-T @ c2w[:3,:3]^T; points = depth * rays_d + c2w[:3,3]
  depth_to_normal    :898-909   dx = P[2:,1:-1] - P[:-2,1:-1]; dy = P[1:-1,2:] - P[1:-1,:-2];
                                out[1:-1,1:-1] = normalize(cross(dx, dy)); border stays 0
  post-processing    :1043-1053 normal_world = c2w[:3,:3] @ normalize(out_color[3:6], dim=0); depth_normal.permute(2,0,1)

Device- and dtype-agnostic (the reference hard-codes 'cuda' and float32), so the same code gives the float64 value the
GPU results are measured against.  When oracle/_ref/pyref holds the reference's own file (staged by oracle/Makefile where
/root/reference exists), tests also call ITS functions directly -- see test_gpu_epilogue.py.
"""
import math

import torch


def depths_to_points(world_view_transform, W, H, FoVx, FoVy, depthmap):
    dt, dev = depthmap.dtype, depthmap.device
    c2w = (world_view_transform.to(dt).T).inverse()
    fx = W / (2 * math.tan(FoVx / 2.))
    fy = H / (2 * math.tan(FoVy / 2.))
    intrins = torch.tensor([[fx, 0., W / 2.], [0., fy, H / 2.], [0., 0., 1.0]], dtype=dt, device=dev)
    grid_x, grid_y = torch.meshgrid(torch.arange(W, device=dev, dtype=dt), torch.arange(H, device=dev, dtype=dt), indexing='xy')
    points = torch.stack([grid_x, grid_y, torch.ones_like(grid_x)], dim=-1).reshape(-1, 3)
    rays_d = points @ intrins.inverse().T @ c2w[:3, :3].T
    rays_o = c2w[:3, 3]
    return depthmap.reshape(-1, 1) * rays_d + rays_o


def depth_to_normal(world_view_transform, W, H, FoVx, FoVy, depth):
    points = depths_to_points(world_view_transform, W, H, FoVx, FoVy, depth).reshape(*depth.shape[1:], 3)
    output = torch.zeros_like(points)
    dx = points[2:, 1:-1] - points[:-2, 1:-1]
    dy = points[1:-1, 2:] - points[1:-1, :-2]
    output[1:-1, 1:-1, :] = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    return output


def postprocess(rendered_image, world_view_transform, W, H, FovX, FovY):
    """(normal_world[3,H,W], depth_normal[3,H,W]) of out_color[9,H,W], in rendered_image's dtype."""
    dt = rendered_image.dtype
    wvt = world_view_transform.squeeze().to(dt)
    render_normal = torch.nn.functional.normalize(rendered_image[3:6], p=2, dim=0)
    c2w = (wvt.T).inverse()
    normal_world = (c2w[:3, :3] @ render_normal.reshape(3, -1)).reshape(3, *render_normal.shape[1:])
    depth_normal = depth_to_normal(wvt, W, H, FovX, FovY, rendered_image[6:7])
    return normal_world, depth_normal.permute(2, 0, 1)
