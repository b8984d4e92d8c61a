"""The cycle-aggregative render loop of F3D-Gaus as a library function.

Reference: inline script code in visualize.py:281-340 --
  1. predict a Gaussian set per scene from the source image                      (:282-283, the model)
  2. render K novel views of every scene from it                                 (:293-314)
  3. feed every rendered (rgb clamped to [0,1], alpha, depth) back to the predictor with that view's
     view_to_world transform and quaternion                                      (:326-334)
  4. concatenate all K+1 Gaussian sets on dim 1                                  (:336-340)
The merged set ([bs, (1+K)*H*W, .]) is what the final orbit is rendered from (:387-402).

What is ours: step 2 is ONE batched pass per scene (gaussian_renderer.render_views ->
gof_forward_batch) instead of K*bs blocking calls, nothing is copied to the host (the reference does
three `.cpu()` per frame and copies them back, :304-306,328-332), and the concat writes into a
preallocated buffer.  The predictor itself (Unet_GS_gtunet) is the reference's network and is passed
in as a callable; `unproject_predictor` is a deterministic stand-in with the same output contract
(one Gaussian per pixel, un-projected along the predictor's ray grid) for tests and benchmarks,
since the checkpoint is not shipped.
"""
from __future__ import annotations

import math
from typing import Callable

import torch

from . import cameras as _cameras

PC_KEYS = ("xyz", "opacity", "scaling", "rotation", "features_dc", "features_rest")
SH_C0 = 0.28209479177387814


def view_quaternions(view_to_world: torch.Tensor) -> torch.Tensor:
    """source_cv2wT_quat of visualize.py:275-277: quaternion (r,x,y,z) of view_to_world[:3,:3]^T per view."""
    return torch.stack([_cameras.matrix_to_quaternion(v[:3, :3].transpose(0, 1).contiguous().cpu())
                        for v in view_to_world.reshape(-1, 4, 4)]).to(view_to_world.device)


def from_unet_gs(model, background, squre_clip: bool = True) -> Callable:
    """Adapter for the reference network: `model` is an Unet_GS_gtunet (src/unet_gs.py:48-101), called the
    way visualize.py:334 calls it."""
    def predict(novel_img, view_to_world, quat, depth):
        _, _, out = model(novel_img, background, view_to_world, quat, return_3d_features=True, render=False,
                          squre_clip=squre_clip, unet_depth=depth)
        return out
    return predict


def from_network(network, cfg: dict, device, squre_clip: float = 10000.0) -> Callable:
    """Predictor = the reference's UNet alone (`gaussian_predictor.network_with_offset` / `network_wo_offset`, called as
    src/gaussian_predictor.py:929-945) followed by the fused output head (predictor_head.PredictorHead, one kernel
    instead of :954-1008).  `network(x[B*V,C_in,H,W], film_camera_emb=None, N_views_xa=...) -> [B*V,C,H,W]`."""
    from .predictor_head import PredictorHead
    head = PredictorHead(cfg, device)
    xa = bool(cfg["model"].get("cross_view_attention", True))

    def predict(novel_img, view_to_world, quat, depth):
        B, V = novel_img.shape[:2]
        x = novel_img.reshape(B * V, *novel_img.shape[2:])
        const_offset = None
        if head.origin_distances:                                   # :915-917
            const_offset, x = x[:, 3:], x[:, :3]
        raw = network(x, film_camera_emb=None, N_views_xa=V if xa else 1)
        out = head(raw, depth, view_to_world, quat, B, V, squre_clip=squre_clip, const_offset=const_offset)
        return {k: out[k] for k in PC_KEYS}
    return predict


def unproject_predictor(cfg: dict, scale: float = 0.01) -> Callable:
    """Stand-in predictor with the reference's output contract (src/gaussian_predictor.py:857-881,954-1002):
    one Gaussian per pixel at `ray_dir * depth` (y inverted, :657-670), moved to the world frame with the
    row-vector view_to_world (:961-966), rotation = the view's quaternion, SH degree 1 with features_dc
    from the rgb and zero features_rest, opacity from alpha.  No learned weights."""
    fov = cfg["model"]["fov"]

    def predict(novel_img, view_to_world, quat, depth):
        B, _, _, H, W = novel_img.shape
        dev = novel_img.device
        f = W / (2 * math.tan(math.radians(fov) / 2))
        i = torch.arange(W, dtype=torch.float32, device=dev)
        x = (i + 0.5 - W / 2) / f
        y = -(torch.arange(H, dtype=torch.float32, device=dev) + 0.5 - H / 2) / f
        ray = torch.stack([x[None, :].expand(H, W), y[:, None].expand(H, W), torch.ones(H, W, device=dev)], dim=-1)
        d = depth.reshape(B, H, W, 1)
        mid = 0.5 * (cfg["dataset_params"]["z_near"] + cfg["dataset_params"]["z_far"])
        d = torch.where(d > 0, d, torch.full_like(d, mid))          # uncovered pixels: mid depth
        pos = (ray[None] * d).reshape(B, H * W, 3)
        pos = torch.cat([pos, torch.ones(B, H * W, 1, device=dev)], dim=-1).bmm(view_to_world.reshape(B, 4, 4))[..., :3]
        rgb = novel_img[:, 0, 0:3].permute(0, 2, 3, 1).reshape(B, H * W, 1, 3)
        alpha = novel_img[:, 0, 3:4].permute(0, 2, 3, 1).reshape(B, H * W, 1)
        return {
            "xyz": pos.contiguous(),
            "opacity": (0.9 * alpha).clamp(0.0, 0.99).contiguous(),
            "scaling": (scale * d.reshape(B, H * W, 1) / mid).expand(B, H * W, 3).contiguous(),
            "rotation": quat.reshape(B, 1, 4).expand(B, H * W, 4).contiguous(),
            "features_dc": ((rgb - 0.5) / SH_C0).contiguous(),
            "features_rest": torch.zeros(B, H * W, 3, 3, device=dev),
        }
    return predict


def render_scene_views(pc: dict, cams, cfg: dict, background: torch.Tensor, workspace=None, render_fn=None):
    """Render all views of every scene of `pc` ([B, P, .] per key).  Returns rgb[B,V,3,H,W] (unclamped),
    depth[B,V,1,H,W], alpha[B,V,1,H,W] on the device.  One batched rasterizer pass per scene."""
    if render_fn is None:
        from .gaussian_renderer import render_views as render_fn
    B = pc["xyz"].shape[0]
    V = cams.world_view.shape[0]
    H = W = int(cfg["model"]["training_resolution"])
    dev = pc["xyz"].device
    rgb = torch.empty((B, V, 3, H, W), dtype=torch.float32, device=dev)
    depth = torch.empty((B, V, 1, H, W), dtype=torch.float32, device=dev)
    alpha = torch.empty((B, V, 1, H, W), dtype=torch.float32, device=dev)
    for b in range(B):
        while True:
            o = render_fn(pc, b, cams.world_view, cams.full_proj, cams.centers, background, cfg, workspace=workspace,
                          epilogue=False)
            rgb[b].copy_(o["render"])
            depth[b].copy_(o["rendered_depth"])
            alpha[b].copy_(o["rendered_alpha"])
            if workspace is None or workspace.finish() is not None:
                break                                   # else: the binning blob was grown, render this scene again
    return rgb, depth, alpha


def cycle_aggregate(pc: dict, predict: Callable, cams, cfg: dict, background: torch.Tensor, workspace=None,
                    render_fn=None):
    """visualize.py:288-340.  `pc`: the source view's Gaussian set ([B, P0, .] per key); `cams`: the K
    aggregation views (cameras.Cameras on pc's device); `predict(novel_img[B,1,4,H,W], view_to_world[B,1,4,4],
    quat[B,1,4], depth[B,1,H,W]) -> dict` with the same keys.  Returns (merged set [B, P0 + K*H*W, .],
    {"rgb", "depth", "alpha"} of the K rendered views)."""
    B = pc["xyz"].shape[0]
    K = cams.world_view.shape[0]
    rgb, depth, alpha = render_scene_views(pc, cams, cfg, background, workspace, render_fn)
    rgb = rgb.clamp(0, 1)                               # visualize.py:311
    quats = view_quaternions(cams.view_to_world)        # [K,4]
    keys = [k for k in pc.keys()]
    sets = [pc]
    for k in range(K):
        novel_img = torch.cat([rgb[:, k:k + 1], alpha[:, k:k + 1]], dim=2)                  # [B,1,4,H,W]
        v2w = cams.view_to_world[k:k + 1].unsqueeze(0).expand(B, -1, -1, -1)                # [B,1,4,4]
        q = quats[k:k + 1].unsqueeze(0).expand(B, -1, -1)                                   # [B,1,4]
        out = predict(novel_img, v2w, q, depth[:, k])
        sets.append({key: out[key] for key in keys})
    # one allocation per key instead of K growing torch.cat calls (visualize.py:336-340)
    merged = {}
    for key in keys:
        total = sum(s[key].shape[1] for s in sets)
        buf = torch.empty((B, total) + tuple(pc[key].shape[2:]), dtype=pc[key].dtype, device=pc[key].device)
        at = 0
        for s in sets:
            n = s[key].shape[1]
            buf[:, at:at + n].copy_(s[key])
            at += n
        merged[key] = buf
    return merged, {"rgb": rgb, "depth": depth, "alpha": alpha}
