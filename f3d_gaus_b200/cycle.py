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


_QUAT_CACHE: dict = {}


def view_quaternions(view_to_world: torch.Tensor) -> torch.Tensor:
    """source_cv2wT_quat of visualize.py:275-277: quaternion (r,x,y,z) of view_to_world[:3,:3]^T per view.  The reference
    computes it once per camera set, outside its render loops; here the result is cached per camera tensor (same storage,
    shape and in-place version), so that the loop does not read the matrices back to the host on every scene."""
    key = (view_to_world.data_ptr(), view_to_world._version, tuple(view_to_world.shape), view_to_world.device)
    hit = _QUAT_CACHE.get(key)
    if hit is not None and hit[0] is view_to_world:
        return hit[1]
    q = torch.stack([_cameras.matrix_to_quaternion(v[:3, :3].transpose(0, 1).contiguous().cpu())
                     for v in view_to_world.reshape(-1, 4, 4)]).to(view_to_world.device)
    if len(_QUAT_CACHE) > 64:
        _QUAT_CACHE.clear()
    _QUAT_CACHE[key] = (view_to_world, q)          # the camera tensor is kept alive: its address stays its own
    return q


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

    grids: dict = {}

    def ray_grid(H, W, dev):
        key = (H, W, dev)
        ray = grids.get(key)
        if ray is None:
            f = W / (2 * math.tan(math.radians(fov) / 2))
            x = (torch.arange(W, dtype=torch.float32, device=dev) + 0.5 - W / 2) / f
            y = -(torch.arange(H, dtype=torch.float32, device=dev) + 0.5 - H / 2) / f
            ray = grids[key] = torch.stack([x[None, :].expand(H, W), y[:, None].expand(H, W),
                                            torch.ones(H, W, device=dev)], dim=-1).contiguous()
        return ray

    def predict(novel_img, view_to_world, quat, depth):
        B, _, _, H, W = novel_img.shape
        dev = novel_img.device
        ray = ray_grid(H, W, dev)
        d = depth.reshape(B, H, W, 1)
        mid = 0.5 * (cfg["dataset_params"]["z_near"] + cfg["dataset_params"]["z_far"])
        d = torch.where(d > 0, d, mid)                              # uncovered pixels: mid depth
        pos = (ray[None] * d).reshape(B, H * W, 3)
        v2w = view_to_world.reshape(B, 4, 4)
        pos = pos.bmm(v2w[:, :3, :3]) + v2w[:, 3:4, :3]             # [pos, 1] @ view_to_world, row-vector convention
        rgb = novel_img[:, 0, 0:3].permute(0, 2, 3, 1).reshape(B, H * W, 1, 3)
        alpha = novel_img[:, 0, 3:4].permute(0, 2, 3, 1).reshape(B, H * W, 1)
        return {
            "xyz": pos.contiguous(),
            "opacity": (0.9 * alpha).clamp(0.0, 0.99).contiguous(),
            "scaling": (d.reshape(B, H * W, 1) * (scale / mid)).expand(B, H * W, 3).contiguous(),
            "rotation": quat.reshape(B, 1, 4).expand(B, H * W, 4).contiguous(),
            "features_dc": ((rgb - 0.5) / SH_C0).contiguous(),
            "features_rest": torch.zeros(B, H * W, 3, 3, device=dev),
        }
    return predict


def render_scene_views(pc: dict, cams, cfg: dict, background: torch.Tensor, workspace=None, render_fn=None,
                       check_overflow: bool = True):
    """Render all views of every scene of `pc` ([B, P, .] per key).  Returns rgb[B,V,3,H,W] (unclamped),
    depth[B,V,1,H,W], alpha[B,V,1,H,W] on the device.  One batched rasterizer pass per scene.

    With a workspace the pass is sync-free and an undersized binning blob shows only in `workspace.finish()`.
    `check_overflow=True` reads it here (one host synchronisation per scene) and renders the scene again after the
    workspace has grown; with False nothing synchronises and the CALLER checks `workspace.finish()` at its own
    synchronisation point -- a None there means the frames (NaN-poisoned) and everything derived from them are to be
    computed again."""
    if render_fn is None:
        from .gaussian_renderer import render_views as render_fn
    B = pc["xyz"].shape[0]
    per_scene = []
    for b in range(B):
        while True:
            o = render_fn(pc, b, cams.world_view, cams.full_proj, cams.centers, background, cfg, workspace=workspace,
                          epilogue=False)
            # (with several scenes the workspace's mailbox only holds the last one: always check then)
            if workspace is None or not (check_overflow or B > 1) or workspace.finish() is not None:
                break                                   # else: the binning blob was grown, render this scene again
        per_scene.append(o)
    if B == 1:
        o = per_scene[0]
        return o["render"].unsqueeze(0), o["rendered_depth"].unsqueeze(0), o["rendered_alpha"].unsqueeze(0)
    return (torch.stack([o["render"] for o in per_scene]), torch.stack([o["rendered_depth"] for o in per_scene]),
            torch.stack([o["rendered_alpha"] for o in per_scene]))


def cycle_aggregate(pc: dict, predict: Callable, cams, cfg: dict, background: torch.Tensor, workspace=None,
                    render_fn=None, check_overflow: bool = True):
    """visualize.py:288-340.  `pc`: the source view's Gaussian set ([B, P0, .] per key); `cams`: the K
    aggregation views (cameras.Cameras on pc's device); `predict(novel_img[B,1,4,H,W], view_to_world[B,1,4,4],
    quat[B,1,4], depth[B,1,H,W]) -> dict` with the same keys.  Returns (merged set [B, P0 + K*H*W, .],
    {"rgb", "depth", "alpha"} of the K rendered views).  `check_overflow`: see render_scene_views."""
    B = pc["xyz"].shape[0]
    K = cams.world_view.shape[0]
    rgb, depth, alpha = render_scene_views(pc, cams, cfg, background, workspace, render_fn, check_overflow)
    rgb = rgb.clamp(0, 1)                               # visualize.py:311
    quats = view_quaternions(cams.view_to_world)        # [K,4]
    keys = [k for k in pc.keys()]
    sets = [pc]
    novel = torch.cat([rgb, alpha], dim=2)              # [B,K,4,H,W]: every view's network input in one kernel (:329-332)
    for k in range(K):
        v2w = cams.view_to_world[k:k + 1].unsqueeze(0).expand(B, -1, -1, -1)                # [B,1,4,4]
        q = quats[k:k + 1].unsqueeze(0).expand(B, -1, -1)                                   # [B,1,4]
        out = predict(novel[:, k:k + 1], v2w, q, depth[:, k])
        sets.append({key: out[key] for key in keys})
    # one concatenation per key instead of K growing torch.cat calls (visualize.py:336-340)
    merged = {key: torch.cat([s[key] for s in sets], dim=1) for key in keys}
    return merged, {"rgb": rgb, "depth": depth, "alpha": alpha}
