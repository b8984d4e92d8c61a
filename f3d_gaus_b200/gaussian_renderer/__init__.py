"""Drop-in for the live part of the reference's `src/gaussian_renderer/__init__.py`:

  render_predicted_more_v2_gof   (:915-1067)  dict-of-tensors -> rasterizer -> output dict
  depths_to_points / depth_to_normal (:881-909)
  render_predicted_more_v2_gof_in (:1070-1228) point integration for mesh extraction (rasterizer.integrate)
  HostFrameSink                   render_views + pipelined D2H of rgb/depth/alpha (the loops' `.cpu()` return path)
  SceneStreamer                   the multi-scene loop as a 2-slot pipeline: H2D of scene k+1 | render k | frames of k-1 to host
  render_views                    all V views of a scene in one batched pass (what the reference's render loops do frame by frame)
  render                          the vanilla signature (src/gaussian-splatting/gaussian_renderer/__init__.py:18-100)

Same arguments, same output keys.  Differences are internal: the rasterizer is libgof_b200,
the per-call `subpixel_offset` allocation is dropped (no kernel reads it), and the ~20 small torch
kernels of the post-processing (normalise, 4x4 inverse, back-projection, cross product) are one fused
epilogue kernel (gof_render_epilogue) with a hand-written backward for the training path.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from .. import _lib
from ..staging import PinnedScene  # noqa: F401  (public here; defined in a module that does not load the CUDA library)
from ..diff_gof_rasterization import (BatchWorkspace, GaussianRasterizationSettings_GOF, GaussianRasterizer_GOF,
                                     _absent, _on_device, rasterize_gaussians, rasterize_views)

_EMPTY_OFFSET: dict = {}


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def _epilogue_forward(raster, vm, V, W, H, FovX, FovY, want_normal=True, want_depth_normal=True):
    dev = raster.device
    normal_world = torch.empty((V, 3, H, W), dtype=torch.float32, device=dev) if want_normal else None
    depth_normal = torch.empty((V, 3, H, W), dtype=torch.float32, device=dev) if want_depth_normal else None
    with _on_device(dev):
        rc = _lib.lib.gof_render_epilogue_batch(raster.data_ptr(), vm.data_ptr(), V, W, H, ctypes.c_float(FovX),
                                                ctypes.c_float(FovY),
                                                normal_world.data_ptr() if want_normal else None,
                                                depth_normal.data_ptr() if want_depth_normal else None,
                                                _lib.raw_stream(dev))
    _lib.check(rc, "gof_render_epilogue_batch")
    return normal_world, depth_normal


class _FusedEpilogue(torch.autograd.Function):
    """(normal_world, depth_normal) = epilogue(out_color) with a hand-written backward (csrc/epilogue.cu): the
    training path of render_predicted_more_v2_gof without the reference's ~20 torch kernels and 4x4 inverse() per
    frame (src/gaussian_renderer/__init__.py:881-909,1043-1053).  out_color: [V,9,H,W], vm: [V,16]."""

    @staticmethod
    def forward(ctx, raster, vm, W, H, FovX, FovY):
        raster = raster.contiguous()
        ctx.save_for_backward(raster, vm)
        ctx.dims = (int(raster.shape[0]), W, H, FovX, FovY)
        return _epilogue_forward(raster, vm, raster.shape[0], W, H, FovX, FovY)

    @staticmethod
    def backward(ctx, g_normal_world, g_depth_normal):
        raster, vm = ctx.saved_tensors
        V, W, H, FovX, FovY = ctx.dims
        dev = raster.device
        grad = torch.empty_like(raster)           # every element is written by the kernel
        keep = [t.contiguous() if t is not None else None for t in (g_normal_world, g_depth_normal)]
        with torch.cuda.device(dev):
            rc = _lib.lib.gof_render_epilogue_backward_batch(
                raster.data_ptr(), vm.data_ptr(), V, W, H, ctypes.c_float(FovX), ctypes.c_float(FovY),
                keep[0].data_ptr() if keep[0] is not None else None,
                keep[1].data_ptr() if keep[1] is not None else None, grad.data_ptr(), _lib.raw_stream(dev))
        _lib.check(rc, "gof_render_epilogue_backward_batch")
        return grad, None, None, None, None, None


def fused_epilogue(rendered_image, world_view_transform, W, H, FovX, FovY):
    """normal_world[3,H,W], depth_normal[3,H,W] from out_color[9,H,W] in one kernel; differentiable w.r.t.
    `rendered_image` (the normal channels 3..5 and the median depth 6) when it carries a graph."""
    if torch.is_grad_enabled() and rendered_image.requires_grad:
        vm = world_view_transform.reshape(1, 16).contiguous()
        normal_world, depth_normal = _FusedEpilogue.apply(rendered_image.unsqueeze(0), vm, W, H, FovX, FovY)
        return normal_world[0], depth_normal[0]
    # inference: the single-frame entry point on the tensors as they are (no views, no reshapes on the per-frame path)
    dev = rendered_image.device
    img = rendered_image if rendered_image.is_contiguous() else rendered_image.contiguous()
    vm = world_view_transform if world_view_transform.is_contiguous() else world_view_transform.contiguous()
    if vm.numel() != 16 or vm.dtype is not torch.float32 or img.dtype is not torch.float32 or vm.device != dev:
        raise RuntimeError("fused_epilogue: expected a float32 [9,H,W] frame and a 4x4 float32 view matrix on the same device")
    normal_world = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    depth_normal = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    with _on_device(dev):
        rc = _lib.lib.gof_render_epilogue(img.data_ptr(), vm.data_ptr(), W, H, FovX, FovY, normal_world.data_ptr(),
                                          depth_normal.data_ptr(), _lib.raw_stream(dev))
    _lib.check(rc, "gof_render_epilogue")
    return normal_world, depth_normal


def depth_to_normal(world_view_transform, image_width, image_height, FoVx, FoVy, depth):
    """Reference helper name (:898-909): normals [H,W,3] from finite differences of the back-projected depth
    [1,H,W], zero on the border.  Served by the fused epilogue kernel (no 4x4 inverse, no meshgrid)."""
    H, W = int(image_height), int(image_width)
    raster = torch.zeros((1, _lib.OUTPUT_CHANNELS, H, W), dtype=torch.float32, device=depth.device)
    raster[0, 6] = depth.reshape(H, W)
    vm = world_view_transform.reshape(1, 16).contiguous()
    if torch.is_grad_enabled() and depth.requires_grad:
        _, dn = _FusedEpilogue.apply(raster, vm, W, H, FoVx, FoVy)
    else:
        _, dn = _epilogue_forward(raster, vm, 1, W, H, FoVx, FoVy, want_normal=False)
    return dn[0].permute(1, 2, 0)


def depths_to_points(world_view_transform, image_width, image_height, FoVx, FoVy, depthmap):
    """Reference helper name (:881-896): world points [H*W,3] of a depth map [1,H,W] (pixel (x,y) -> camera ray
    ((x-W/2)/fx, (y-H/2)/fy, 1) * depth, moved to the world frame)."""
    W, H = int(image_width), int(image_height)
    dev = depthmap.device
    view_to_world = world_view_transform.reshape(4, 4).inverse()          # row-vector convention: p_w = [p_v,1] @ V2W
    cx = (torch.arange(W, device=dev, dtype=torch.float32) - W / 2.) / (W / (2 * math.tan(FoVx / 2.)))
    cy = (torch.arange(H, device=dev, dtype=torch.float32) - H / 2.) / (H / (2 * math.tan(FoVy / 2.)))
    d = depthmap.reshape(H, W)
    cam = torch.stack([cx[None, :] * d, cy[:, None] * d, d], dim=-1).reshape(-1, 3)
    return cam @ view_to_world[:3, :3] + view_to_world[3, :3]


_SH_CACHE: dict = {}


def _packed_sh(features_dc: torch.Tensor, features_rest: torch.Tensor) -> torch.Tensor:
    """[P,1,3] + [P,M-1,3] -> contiguous [P,M,3] (the reference's per-call torch.cat, :1006).  The render loops call the
    renderer once per view with the SAME scene tensors, so under no_grad the packed copy is cached for as long as the
    two inputs are unmodified (same storage, shape and in-place version counter); with autograd the cat is part of the
    graph and is done every time."""
    if torch.is_grad_enabled() and (features_dc.requires_grad or features_rest.requires_grad):
        return torch.cat([features_dc, features_rest], dim=1).contiguous()
    key = (features_dc.data_ptr(), features_rest.data_ptr(), features_dc._version, features_rest._version,
           tuple(features_dc.shape), tuple(features_rest.shape), features_dc.device)
    hit = _SH_CACHE.get("last")
    if hit is not None and hit[0] == key:
        return hit[1]
    shs = torch.cat([features_dc, features_rest], dim=1).contiguous()
    _SH_CACHE["last"] = (key, shs, features_dc, features_rest)      # the inputs are kept alive: their addresses stay theirs
    return shs


def _subpixel_offset(H, W, device):
    # The reference allocates zeros(H,W,2) per call (:954); no kernel reads it, so share one.
    key = (H, W, str(device))
    t = _EMPTY_OFFSET.get(key)
    if t is None:
        t = _EMPTY_OFFSET[key] = torch.zeros((H, W, 2), dtype=torch.float32, device=device)
    return t


def render_predicted_more_v2_gof(pc: dict, bs, world_view_transform, full_proj_transform, camera_center,
                                 bg_color: torch.Tensor, cfg, kernel_size=0.0, scaling_modifier=1.0,
                                 override_color=None, subpixel_offset=None):
    """Render scene `bs` of the predicted Gaussian dict `pc` (reference :915-1067)."""
    xyz = pc["xyz"][bs]
    device = xyz.device
    if torch.is_grad_enabled():
        screenspace_points = torch.zeros_like(xyz, dtype=pc["xyz"].dtype, requires_grad=True, device=device) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    else:
        screenspace_points = torch.zeros_like(xyz)          # no graph to retain a gradient in

    fov = cfg['model']['fov']
    tanfovx = math.tan(fov * np.pi / 360)
    tanfovy = math.tan(fov * np.pi / 360)
    FovX = fov * np.pi / 180
    FovY = fov * np.pi / 180
    image_height = int(cfg['model']['training_resolution'])
    image_width = int(cfg['model']['training_resolution'])

    raster_settings = GaussianRasterizationSettings_GOF(
        image_height=image_height, image_width=image_width, tanfovx=tanfovx, tanfovy=tanfovy,
        kernel_size=kernel_size, subpixel_offset=_subpixel_offset(image_height, image_width, device),
        bg=bg_color, scale_modifier=scaling_modifier, viewmatrix=world_view_transform,
        projmatrix=full_proj_transform, sh_degree=cfg['model']['max_sh_degree'], campos=camera_center,
        prefiltered=False, debug=False)
    opacity = pc["opacity"][bs]
    scales = pc["scaling"][bs]
    rotations = pc["rotation"][bs]
    absent = _absent()
    if override_color is None:
        shs = _packed_sh(pc["features_dc"][bs], pc["features_rest"][bs])
        colors_precomp = absent
    else:
        shs = absent
        colors_precomp = pc["rgbs"][bs]
    # GaussianRasterizer_GOF(raster_settings)(...) of the reference (:1023-1041) without building an nn.Module per frame:
    # its forward is this one call (the argument combination is fixed here, so there is nothing to validate)
    rendered_image, radii = rasterize_gaussians(xyz, screenspace_points, shs, colors_precomp, opacity, scales, rotations,
                                                absent, absent, raster_settings)

    normal_world, depth_normal = fused_epilogue(rendered_image, world_view_transform, image_width, image_height,
                                                FovX, FovY)

    return {"render": rendered_image[:3, :, :],
            "rendered_normal": normal_world,
            "rendered_depth": rendered_image[6:7, :, :],
            "depth_normal": depth_normal,
            "rendered_alpha": rendered_image[7:8, :, :],
            "distortion_map": rendered_image[8:9, :, :],
            "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0,
            "radii": radii}


def render_predicted_more_v2_gof_in(points3D, pc: dict, bs, world_view_transform, full_proj_transform, camera_center,
                                    bg_color: torch.Tensor, cfg, kernel_size=0.0, scaling_modifier=1.0,
                                    override_color=None, subpixel_offset=None):
    """Point integration for mesh extraction (reference :1070-1228): the dict of
    `render_predicted_more_v2_gof` plus `alpha_integrated[PN]` and `color_integrated[PN,3]`."""
    xyz = pc["xyz"][bs]
    device = xyz.device
    screenspace_points = torch.zeros_like(xyz, dtype=pc["xyz"].dtype, requires_grad=True, device=device) + 0
    fov = cfg['model']['fov']
    tanfov = math.tan(fov * np.pi / 360)
    Fov = fov * np.pi / 180
    H = W = int(cfg['model']['training_resolution'])
    raster_settings = GaussianRasterizationSettings_GOF(
        image_height=H, image_width=W, tanfovx=tanfov, tanfovy=tanfov, kernel_size=kernel_size,
        subpixel_offset=_subpixel_offset(H, W, device), bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=world_view_transform, projmatrix=full_proj_transform, sh_degree=cfg['model']['max_sh_degree'],
        campos=camera_center, prefiltered=False, debug=False)
    rasterizer = GaussianRasterizer_GOF(raster_settings=raster_settings)
    if override_color is None:
        shs, colors_precomp = torch.cat([pc["features_dc"][bs], pc["features_rest"][bs]], dim=1).contiguous(), None
    else:
        shs, colors_precomp = None, pc["rgbs"][bs]
    rendered_image, alpha_integrated, color_integrated, radii = rasterizer.integrate(
        points3D=points3D, means3D=xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
        opacities=pc["opacity"][bs], scales=pc["scaling"][bs], rotations=pc["rotation"][bs], cov3D_precomp=None,
        view2gaussian_precomp=None)
    normal_world, depth_normal = fused_epilogue(rendered_image, world_view_transform, W, H, Fov, Fov)
    return {"render": rendered_image[:3], "rendered_normal": normal_world, "rendered_depth": rendered_image[6:7],
            "depth_normal": depth_normal, "rendered_alpha": rendered_image[7:8], "distortion_map": rendered_image[8:9],
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "alpha_integrated": alpha_integrated, "color_integrated": color_integrated, "radii": radii}


def render_views(pc: dict, bs, world_view_transforms, full_proj_transforms, camera_centers, bg_color: torch.Tensor, cfg,
                 kernel_size=0.0, scaling_modifier=1.0, override_color=None, workspace: BatchWorkspace | None = None,
                 epilogue: bool = True, out_color: torch.Tensor | None = None, sink: torch.Tensor | None = None):
    """All V views of scene `bs` in one pass of the pipeline: the batched form of the reference's
    `for th in range(num_frames): render_predicted_more_v2_gof(pc, bs, wvt[th:th+1], ...)` loops
    (visualize.py:293-306,387-402).  Inference only (no autograd graph).

    world_view_transforms / full_proj_transforms: [V,4,4] or [V,1,4,4]; camera_centers: [V,3] or [V,1,3].
    Returns the keys of `render_predicted_more_v2_gof` with a leading view dimension
    (render[V,3,H,W], rendered_depth[V,1,H,W], rendered_alpha[V,1,H,W], distortion_map[V,1,H,W],
    radii[V,P], visibility_filter[V,P]; rendered_normal / depth_normal [V,3,H,W] when `epilogue`),
    plus `raster` = the full [V,9,H,W] rasterizer output.  Frame v is bit-identical to the per-view call.
    With a `workspace` the call does not synchronise the host (see BatchWorkspace).  `sink`: see `rasterize_views`."""
    xyz = pc["xyz"][bs]
    device = xyz.device
    fov = cfg['model']['fov']
    tanfov = math.tan(fov * np.pi / 360)
    Fov = fov * np.pi / 180
    H = W = int(cfg['model']['training_resolution'])
    V = int(world_view_transforms.reshape(-1, 16).shape[0])
    if override_color is None:
        shs = torch.cat([pc["features_dc"][bs], pc["features_rest"][bs]], dim=1).contiguous()
        colors = None
    else:
        shs, colors = None, pc["rgbs"][bs]
    with torch.no_grad():
        R, raster, radii, _, _, _ = rasterize_views(
            bg_color, xyz, colors, pc["opacity"][bs], pc["scaling"][bs], pc["rotation"][bs], scaling_modifier,
            world_view_transforms, full_proj_transforms, tanfov, tanfov, kernel_size, H, W, shs,
            cfg['model']['max_sh_degree'], camera_centers, workspace=workspace, out_color=out_color, sink=sink)
        out = {"raster": raster, "render": raster[:, 0:3], "rendered_depth": raster[:, 6:7],
               "rendered_alpha": raster[:, 7:8], "distortion_map": raster[:, 8:9], "radii": radii,
               "visibility_filter": radii > 0, "num_rendered": R}
        if epilogue:
            vm = world_view_transforms.reshape(V, 16).contiguous()
            out["rendered_normal"], out["depth_normal"] = _epilogue_forward(raster, vm, V, W, H, Fov, Fov)
    return out


class HostFrameSink:
    """Return path of the render loops to HOST memory (the reference does `.cpu()` on rgb, depth and alpha of every
    frame, visualize.py:304-306).  `host` is a pinned [V,5,H,W] buffer (rgb, median depth, alpha).

    zero_copy (default): the blend kernel itself stores the five channels into `host` (gof_set_frame_sink) -- posted
    PCIe writes that drain while the remaining tiles blend, so the read-back adds little time after the kernel.
    `host` is then channels_last in memory ([V,H,W,5] storage, the layout image writers want and the one that gives
    320-byte PCIe bursts per tile row) unless channels_last=False.
    Otherwise the V views are rendered in `chunks` batched passes and each pass's block is packed on the device and
    copied by the DMA engine on a side stream while the next pass renders (tools/e2e_breakdown.py compares them).
    `finish()` waits for the frames and returns the per-view num_rendered (None => a binning blob overflowed and was
    grown: call `render` again)."""

    def __init__(self, V: int, H: int, W: int, device, chunks: int = 1, zero_copy: bool = True,
                 channels_last: bool = True):
        self.V, self.H, self.W, self.device = V, H, W, torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.chunks = max(1, min(chunks, V))
        self.per = -(-V // self.chunks)
        self.zero_copy = bool(zero_copy)
        if self.zero_copy and channels_last:
            self.host = torch.empty((V, H, W, 5), dtype=torch.float32).pin_memory().permute(0, 3, 1, 2)
        else:
            self.host = torch.empty((V, 5, H, W), dtype=torch.float32).pin_memory()
        self.staging = None if self.zero_copy else torch.empty((V, 5, H, W), dtype=torch.float32, device=self.device)
        self.copy_stream = None if self.zero_copy else torch.cuda.Stream(device=self.device)
        self.workspaces = [BatchWorkspace(self.device) for _ in range(self.chunks)]
        self.done = torch.cuda.Event()
        self.last_raster = None

    def render(self, pc, bs, world_view_transforms, full_proj_transforms, camera_centers, bg_color, cfg, **kw):
        main = torch.cuda.current_stream(self.device)
        wv = world_view_transforms.reshape(self.V, 4, 4)
        fp = full_proj_transforms.reshape(self.V, 4, 4)
        cc = camera_centers.reshape(self.V, 3)
        for ci in range(self.chunks):
            lo, hi = ci * self.per, min(self.V, (ci + 1) * self.per)
            if lo >= hi:
                break
            o = render_views(pc, bs, wv[lo:hi], fp[lo:hi], cc[lo:hi], bg_color, cfg, workspace=self.workspaces[ci],
                             epilogue=False, sink=self.host[lo:hi] if self.zero_copy else self.staging[lo:hi], **kw)
            self.last_raster = o["raster"]            # [v,9,H,W] of the last pass (the whole batch when chunks == 1)
            if self.zero_copy:
                continue
            st = self.staging[lo:hi]                  # packed by the blend kernel (device sink)
            ready = torch.cuda.Event()
            ready.record(main)
            self.copy_stream.wait_event(ready)
            with torch.cuda.stream(self.copy_stream):
                self.host[lo:hi].copy_(st, non_blocking=True)
        if self.zero_copy:
            self.done.record(main)
        else:
            self.done.record(self.copy_stream)
            main.wait_event(self.done)      # a later synchronisation of the main stream covers the copies
        return self.host

    def finish(self):
        self.done.synchronize()
        out = []
        for ws in self.workspaces[:self.chunks]:
            if ws.key is None:
                continue
            r = ws.finish()
            if r is None:
                return None
            out += r
        return out


class SceneStreamer:
    """Pipelined multi-scene render loop to HOST memory -- the streaming form of the reference's
    `for batch: for scene: for view: render(...).cpu()` loops (visualize.py:221,293-306,387-402).

    `slots` scenes are in flight: while scene k renders, the H2D copy of scene k+1 runs on a copy stream and the
    frames of scene k-1 drain to pinned host memory (stored there by the blend kernel itself, `zero_copy`, or stored
    by it into a packed device block that the DMA engine copies on a third stream -- the better choice when many GPUs
    share one host: DMA bursts use the host's write path more efficiently than the kernel's 320-byte posted writes).  Nothing synchronises the host except `collect()`, which
    waits on ONE event: that of the oldest outstanding scene.  The binning-overflow check of a scene rides on the same
    event (BatchWorkspace.finish_async), so it does not stall the scenes behind it; an overflowed scene is re-rendered
    inside `collect()` after its workspace has grown.

        st = SceneStreamer(V, H, W, device, cams, bg, cfg)
        for k, scene in enumerate(scenes):                 # scene: staging.PinnedScene (one pinned slab)
            if st.pending == st.slots:
                frames, R = st.collect()                   # [V,5,H,W] pinned host view of scene k - slots: consume / copy it
            st.submit(scene)
        while st.pending: frames, R = st.collect()

    The host view returned by `collect()` is reused by the submit after the next `slots - 1` ones."""

    def __init__(self, V: int, H: int, W: int, device, world_view_transforms, full_proj_transforms, camera_centers,
                 bg_color, cfg, slots: int = 2, zero_copy: bool = True, channels_last: bool = True):
        self.V, self.H, self.W = V, H, W
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.cams = (world_view_transforms.reshape(V, 4, 4), full_proj_transforms.reshape(V, 4, 4), camera_centers.reshape(V, 3))
        self.bg, self.cfg = bg_color, cfg
        self.slots = max(1, int(slots))
        self.zero_copy = bool(zero_copy)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.d2h_stream = None if self.zero_copy else torch.cuda.Stream(device=self.device)
        self.slot = []
        for _ in range(self.slots):
            if self.zero_copy and channels_last:
                host = torch.empty((V, H, W, 5), dtype=torch.float32).pin_memory().permute(0, 3, 1, 2)
            else:
                host = torch.empty((V, 5, H, W), dtype=torch.float32).pin_memory()
            self.slot.append({"host": host, "slab": None, "ws": BatchWorkspace(self.device), "scene": None, "pc": None,
                              "staging": None if self.zero_copy else torch.empty((V, 5, H, W), dtype=torch.float32, device=self.device),
                              "raster": torch.empty((V, 9, H, W), dtype=torch.float32, device=self.device),
                              "h2d": torch.cuda.Event(), "done": torch.cuda.Event(), "packed": torch.cuda.Event()})
        self.head = 0          # next slot to submit into
        self.pending = 0       # scenes submitted and not yet collected

    def _render(self, s):
        wv, fp, cc = self.cams
        # the blend kernel packs rgb / depth / alpha itself: into the pinned host buffer (zero_copy) or into a device
        # staging block that the DMA engine then moves on its own stream -- no pack kernels either way
        render_views(s["pc"], 0, wv, fp, cc, self.bg, self.cfg, workspace=s["ws"], epilogue=False, out_color=s["raster"],
                     sink=s["host"] if self.zero_copy else s["staging"])
        s["ws"].finish_async()                     # the mailbox copy rides behind the render on the compute stream
        main = torch.cuda.current_stream(self.device)
        if self.zero_copy:
            s["done"].record(main)
            return
        s["packed"].record(main)
        self.d2h_stream.wait_event(s["packed"])
        with torch.cuda.stream(self.d2h_stream):
            s["host"].copy_(s["staging"], non_blocking=True)
            s["done"].record(self.d2h_stream)

    def submit(self, scene):
        """Enqueue scene (a staging.PinnedScene): H2D on the copy stream, render on the current stream.  Never blocks
        the host; raises if all slots are in flight (call collect() first)."""
        if self.pending == self.slots:
            raise RuntimeError("SceneStreamer: all slots in flight; collect() before submitting more")
        s = self.slot[self.head]
        self.head = (self.head + 1) % self.slots
        self.pending += 1
        nbytes = scene.host_slab.numel()
        if s["slab"] is None or s["slab"].numel() < nbytes:
            s["slab"] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            s["pc"] = scene.upload(s["slab"])
            s["h2d"].record(self.copy_stream)
        s["scene"] = scene
        main.wait_event(s["h2d"])
        self._render(s)

    def collect(self):
        """Wait for the OLDEST outstanding scene; returns (frames [V,5,H,W] in pinned host memory, per-view num_rendered)."""
        if self.pending == 0:
            raise RuntimeError("SceneStreamer: nothing to collect")
        s = self.slot[(self.head - self.pending) % self.slots]
        s["done"].synchronize()
        R = s["ws"].finish_poll()
        while R is None:                           # binning blob too small: it has been grown, render this scene again
            self._render(s)
            s["done"].synchronize()
            R = s["ws"].finish_poll()
        self.pending -= 1
        return s["host"], R


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
           kernel_size=0.0):
    """The vanilla `render()` signature (src/gaussian-splatting/gaussian_renderer/__init__.py:18-100)
    on top of the GOF rasterizer.  `viewpoint_camera` provides FoVx, FoVy, image_height, image_width,
    world_view_transform, full_proj_transform, camera_center; `pc` provides get_xyz, get_opacity,
    get_scaling, get_rotation, get_features, active_sh_degree (and get_covariance when
    pipe.compute_cov3D_python)."""
    xyz = pc.get_xyz
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    H, W = int(viewpoint_camera.image_height), int(viewpoint_camera.image_width)
    raster_settings = GaussianRasterizationSettings_GOF(
        image_height=H, image_width=W, tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), kernel_size=kernel_size,
        subpixel_offset=_subpixel_offset(H, W, xyz.device), bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False,
        debug=bool(getattr(pipe, "debug", False)))
    rasterizer = GaussianRasterizer_GOF(raster_settings=raster_settings)
    shs, colors_precomp = (pc.get_features, None) if override_color is None else (None, override_color)
    rendered_image, radii = rasterizer(means3D=xyz, means2D=screenspace_points, shs=shs,
                                       colors_precomp=colors_precomp, opacities=pc.get_opacity,
                                       scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None,
                                       view2gaussian_precomp=None)
    return {"render": rendered_image[:3], "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii}
