"""The predictor's output head: network output -> Gaussian set, one CUDA kernel (gof_predictor_head).

Mirrors the post-network half of `GaussianSplatPredictor_gtunet` (src/gaussian_predictor.py:596-1008): same cfg keys,
same `ray_dirs` buffer, same call arguments after the UNet (`source_cameras_view_to_world`, `source_cv2wT_quat`,
`squre_clip`, `unet_depth`), same output dict (`xyz, opacity, scaling, rotation, features_dc, unet_depth,
features_rest`, each `[B, N_views * H * W, ...]`, contiguous).  The reference spends ~25 elementwise / permute / bmm /
cat launches and a `make_contiguous` pass here (:954-1008); this is one launch that reads the NCHW planes once and
writes the point lists the rasterizer consumes.

Drop-in use inside the reference's `forward` (the UNet call stays the reference's):

    raw = self.network_with_offset(x, film_camera_emb=None, N_views_xa=N_views_xa)       # [B*V, C, H, W]
    return self.head(raw, unet_depth, source_cameras_view_to_world, source_cv2wT_quat, B, N_views, squre_clip)

CUDA only: there is no CPU path (the numpy restatement in oracle/head_oracle.py is test infrastructure).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _lib


def ray_tables(cfg: dict) -> tuple[torch.Tensor, torch.Tensor]:
    """x row [W] and y column [H] of the module's `ray_dirs` buffer (init_ray_dirs, :657-681), built with the same torch
    CPU ops so the values are the reference buffer's bit for bit (CPU float32 tensors)."""
    m = cfg["model"]
    res = int(m["training_resolution"])
    x = torch.linspace(-res // 2 + 0.5, res // 2 - 0.5, res)
    y = torch.linspace(res // 2 - 0.5, -res // 2 + 0.5, res)
    if m.get("inverted_x", False):
        x = -x
    if m.get("inverted_y", False):
        y = -y
    focal = res / (2 * math.tan((m["fov"] * np.pi / 180) / 2))
    return x / focal, y / focal


class PredictorHead:
    def __init__(self, cfg: dict, device="cuda"):
        m = cfg["model"]
        self.cfg = cfg
        self.res = int(m["training_resolution"])
        self.with_offset = bool(m.get("network_with_offset", False))
        if self.with_offset == bool(m.get("network_without_offset", not self.with_offset)):
            raise ValueError("exactly one of network_with_offset / network_without_offset (gaussian_predictor.py:600-618)")
        self.sh_degree = int(m["max_sh_degree"])
        if self.sh_degree not in (0, 1):
            raise ValueError("Only accepting degree 1")            # the reference's assertion, :993
        self.isotropic = bool(m.get("isotropic", False))
        self.origin_distances = bool(m.get("origin_distances", False))
        self.channels = (3 if self.with_offset else 0) + 1 + 3 + 4 + 3 + (9 if self.sh_degree > 0 else 0)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PredictorHead: CUDA device required; there is no CPU path")
        x, y = ray_tables(cfg)
        self.ray_x, self.ray_y = x.to(self.device).contiguous(), y.to(self.device).contiguous()

    @property
    def ray_dirs(self) -> torch.Tensor:
        """[1,3,H,W], the reference's registered buffer."""
        gx = self.ray_x[None, :].expand(self.res, self.res)
        gy = self.ray_y[:, None].expand(self.res, self.res)
        return torch.stack([gx, gy, torch.ones_like(gx)]).unsqueeze(0)

    def __call__(self, network_out: torch.Tensor, unet_depth: torch.Tensor, source_cameras_view_to_world: torch.Tensor,
                 source_cv2wT_quat: torch.Tensor, B: int, N_views: int, squre_clip: float = 10000.0,
                 const_offset: torch.Tensor | None = None, sh_transform: torch.Tensor | None = None) -> dict:
        """network_out [B*V, C, H, W] (the UNet's raw output, before `.split`), unet_depth [B*V, 1, H, W],
        source_cameras_view_to_world [B, V, 4, 4] (or [B*V, 4, 4]), source_cv2wT_quat [B, V, 4]; const_offset
        [B*V, 1, H, W] when cfg.model.origin_distances (the 4th input channel, :915-917)."""
        assert source_cv2wT_quat is not None                        # :985
        BV, H, W = B * N_views, self.res, self.res
        dev = network_out.device
        if dev.type != "cuda":
            raise RuntimeError("PredictorHead: tensors must be on a CUDA device; there is no CPU path")
        if tuple(network_out.shape) != (BV, self.channels, H, W):
            raise RuntimeError(f"network_out must be [{BV},{self.channels},{H},{W}], got {tuple(network_out.shape)}")
        if self.origin_distances != (const_offset is not None):
            raise RuntimeError("const_offset must be given exactly when cfg.model.origin_distances is set")
        f = lambda t: t.to(dtype=torch.float32).contiguous()
        net, depth = f(network_out), f(unet_depth).reshape(BV, 1, H, W)
        v2w = f(source_cameras_view_to_world).reshape(BV, 16)
        quat = f(source_cv2wT_quat).reshape(BV, 4)
        co = f(const_offset).reshape(BV, 1, H, W) if const_offset is not None else None
        sht = f(sh_transform).reshape(BV, 9) if sh_transform is not None else None
        N = H * W
        e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        out = {"xyz": e(B, N_views * N, 3), "opacity": e(B, N_views * N, 1), "scaling": e(B, N_views * N, 3),
               "rotation": e(B, N_views * N, 4), "features_dc": e(B, N_views * N, 1, 3),
               "unet_depth": depth.reshape(B, N_views * N, 1),             # flatten_vector of a 1-channel map: same memory
               "features_rest": e(B, N_views * N, 3, 3) if self.sh_degree > 0 else
               torch.zeros((B, N_views * N, 0, 3), dtype=torch.float32, device=dev)}
        prm = _lib.GofHeadParams(BV, H, W, self.channels, int(self.with_offset), self.sh_degree, int(self.isotropic),
                                 float(squre_clip))
        ptr = lambda t: t.data_ptr() if (t is not None and t.numel()) else None
        ray_x, ray_y = (self.ray_x, self.ray_y) if dev == self.ray_x.device else (self.ray_x.to(dev), self.ray_y.to(dev))
        with torch.cuda.device(dev):
            rc = _lib.lib.gof_predictor_head(ctypes.byref(prm), ptr(net), ptr(depth), ptr(co), ptr(ray_x), ptr(ray_y),
                                             ptr(v2w), ptr(quat), ptr(sht), ptr(out["xyz"]), ptr(out["opacity"]),
                                             ptr(out["scaling"]), ptr(out["rotation"]), ptr(out["features_dc"]),
                                             ptr(out["features_rest"]), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "gof_predictor_head")
        return out
