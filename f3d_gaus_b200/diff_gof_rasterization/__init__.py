"""Drop-in replacement for the reference's `diff_gof_rasterization` package
(RAST/diff_gof_rasterization/__init__.py), backed by libgof_b200.so through ctypes.

Exports the reference's names -- `GaussianRasterizationSettings_GOF`, `GaussianRasterizer_GOF`
(:168,:185), `rasterize_gaussians` (:21), the private `_C` module surface (ext.cpp:15-19) --
plus the un-suffixed `GaussianRasterizationSettings` / `GaussianRasterizer` aliases the
vanilla renderer signature uses.  Install it under the reference's import name with
`f3d_gaus_b200.install_drop_in()`.

No CPU path exists: tensors must live on a CUDA device and the CUDA library must load.
"""
from __future__ import annotations

import ctypes
from typing import NamedTuple

import torch
import torch.nn as nn

from .. import _lib

__all__ = [
    "GaussianRasterizationSettings_GOF", "GaussianRasterizer_GOF", "GaussianRasterizationSettings",
    "GaussianRasterizer", "rasterize_gaussians", "NumRendered", "rasterize_views", "rasterize_views_autograd", "BatchWorkspace",
]


def _dev_ptr(t: torch.Tensor | None, device: torch.device, keep: list, dtype=torch.float32):
    """Device pointer of `t` (made contiguous), or None for the reference's "absent" sentinel
    (a tensor with numel()==0, RAST/diff_gof_rasterization/__init__.py:211-225)."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype is not dtype or not t.is_cuda or t.get_device() != device.index:
        if t.device != device:
            raise RuntimeError(f"expected a tensor on {device}, got one on {t.device}")
        raise RuntimeError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():          # (the hot path hands over contiguous tensors: nothing to copy, nothing to keep alive)
        t = t.contiguous()
        keep.append(t)
    return t.data_ptr()


class _on_device:
    """`with torch.cuda.device(d)` for the hot path: nothing to do (and nothing to undo) when `d` is already current."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is None or cur == self.idx:
            self.prev = -1
        else:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


_SIZES: dict = {}


def _state_sizes(P: int, W: int, H: int, V: int = 1):
    """(geom_bytes, img_bytes) of the opaque state blobs -- pure host arithmetic in the library, cached."""
    key = (P, W, H, V)
    got = _SIZES.get(key)
    if got is None:
        gsz, isz, bsz = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(_lib.lib.gof_state_sizes_batch(P, W, H, V, 0, ctypes.byref(gsz), ctypes.byref(isz), ctypes.byref(bsz)),
                   "gof_state_sizes")
        got = _SIZES[key] = (gsz.value, isz.value)
    return got


class _BinningAllocator:
    """The binning-blob allocation callback (the reference's resizeFunctional, rasterize_points.cu:28-34): one
    persistent ctypes callback; `device` is set before each call, `blob` holds the tensor it allocated."""

    def __init__(self):
        self.device = None
        self.blob = None
        self.callback = _lib.ALLOC_FN(self._alloc)

    def _alloc(self, _user, nbytes):
        self.blob = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self.blob.data_ptr()

    def arm(self, device):
        self.device, self.blob = device, None
        return self.callback

    def take(self, device):
        blob, self.blob = self.blob, None
        return blob if blob is not None else torch.empty(0, dtype=torch.uint8, device=device)


_ALLOC = _BinningAllocator()


class NumRendered(int):
    """`num_rendered` as the reference returns it (a Python int)."""


class _CModule:
    """The four functions of the reference's pybind module `_C` (ext.cpp:15-19)."""

    # -- forward ------------------------------------------------------------------------------
    @staticmethod
    def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                            cov3D_precomp, view2gaussian_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                            kernel_size, subpixel_offset, image_height, image_width, sh, degree, campos,
                            prefiltered, debug):
        if means3D.dim() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise RuntimeError("diff_gof_rasterization (B200): tensors must be on a CUDA device; there is no CPU path")
        device = means3D.device
        P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
        M = int(sh.size(1)) if (sh is not None and sh.numel() != 0 and sh.size(0) != 0) else 0
        keep: list = []
        with _on_device(device):
            stream = _lib.raw_stream(device)
            byte_opts = dict(dtype=torch.uint8, device=device)
            if P == 0:
                out_color = torch.zeros((_lib.OUTPUT_CHANNELS, H, W), dtype=torch.float32, device=device)
                radii = torch.zeros((0,), dtype=torch.int32, device=device)
                empty = torch.empty(0, **byte_opts)
                return NumRendered(0), out_color, radii, empty, empty.clone(), empty.clone()

            out_color = torch.empty((_lib.OUTPUT_CHANNELS, H, W), dtype=torch.float32, device=device)
            radii = torch.empty((P,), dtype=torch.int32, device=device)
            gbytes, ibytes = _state_sizes(P, W, H)
            geom = torch.empty(gbytes, **byte_opts)
            img = torch.empty(ibytes, **byte_opts)
            # (inside an autograd.Function's forward grad mode is off, but the inputs still say whether they require grad)
            prm = _lib.GofParams(P, int(degree), M, W, H, float(tan_fovx), float(tan_fovy), float(kernel_size),
                                 float(scale_modifier), int(bool(prefiltered)), int(bool(debug)),
                                 _lib.default_flags(means3D, colors, opacity, scales, rotations, sh, cov3D_precomp,
                                                    view2gaussian_precomp))
            inp = _lib.GofInputs(
                _dev_ptr(background, device, keep), _dev_ptr(means3D, device, keep), _dev_ptr(sh, device, keep),
                _dev_ptr(colors, device, keep), _dev_ptr(opacity, device, keep), _dev_ptr(scales, device, keep),
                _dev_ptr(rotations, device, keep), _dev_ptr(cov3D_precomp, device, keep),
                _dev_ptr(view2gaussian_precomp, device, keep), _dev_ptr(viewmatrix, device, keep),
                _dev_ptr(projmatrix, device, keep), _dev_ptr(campos, device, keep))
            R = ctypes.c_int32(0)
            bin_out = ctypes.c_void_p()
            rc = _lib.lib.gof_forward(_lib.context(device.index), ctypes.byref(prm), ctypes.byref(inp),
                                      geom.data_ptr(), gbytes, img.data_ptr(), ibytes,
                                      None, 0, _ALLOC.arm(device), None, out_color.data_ptr(), radii.data_ptr(),
                                      ctypes.byref(R), ctypes.byref(bin_out), stream)
            _lib.check(rc, "rasterize_gaussians")
            binning = _ALLOC.take(device)
        return NumRendered(R.value), out_color, radii, geom, binning, img

    # -- backward -----------------------------------------------------------------------------
    @staticmethod
    def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                     cov3D_precomp, view2gaussian_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                     kernel_size, subpixel_offset, dL_dout_color, sh, degree, campos, geomBuffer, R,
                                     binningBuffer, imageBuffer, debug):
        device = means3D.device
        P = int(means3D.size(0))
        H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
        M = int(sh.size(1)) if (sh is not None and sh.numel() != 0 and sh.size(0) != 0) else 0
        opts = dict(dtype=torch.float32, device=device)
        if P == 0:
            z = lambda *shape: torch.zeros(shape, **opts)
            return z(0, 3), z(0, 3), z(0, 1), z(0, 3), z(0, 6), z(0, M, 3), z(0, 3), z(0, 4), z(0, 10)
        # every element is written by the library: no zero fill needed
        e = lambda *shape: torch.empty(shape, **opts)
        dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D = e(P, 3), e(P, 3), e(P, 1), e(P, 3)
        dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dv2g = e(P, 6), e(P, M, 3), e(P, 3), e(P, 4), e(P, 10)
        keep: list = []
        with torch.cuda.device(device):
            stream = _lib.raw_stream(device)
            prm = _lib.GofParams(P, int(degree), M, W, H, float(tan_fovx), float(tan_fovy), float(kernel_size),
                                 float(scale_modifier), 0, int(bool(debug)), _lib.default_flags())
            inp = _lib.GofInputs(
                _dev_ptr(background, device, keep), _dev_ptr(means3D, device, keep), _dev_ptr(sh, device, keep),
                _dev_ptr(colors, device, keep), None, _dev_ptr(scales, device, keep),
                _dev_ptr(rotations, device, keep), _dev_ptr(cov3D_precomp, device, keep),
                _dev_ptr(view2gaussian_precomp, device, keep), _dev_ptr(viewmatrix, device, keep),
                _dev_ptr(projmatrix, device, keep), _dev_ptr(campos, device, keep))
            grads = _lib.GofGrads(dL_dmeans2D.data_ptr(), dL_dcolors.data_ptr(), dL_dopacity.data_ptr(),
                                  dL_dmeans3D.data_ptr(), dL_dcov3D.data_ptr(),
                                  dL_dsh.data_ptr() if M > 0 else None, dL_dscales.data_ptr(),
                                  dL_drotations.data_ptr(), dL_dv2g.data_ptr())
            rc = _lib.lib.gof_backward(
                _lib.context(device.index), ctypes.byref(prm), ctypes.byref(inp), int(R),
                _dev_ptr(radii, device, keep, torch.int32),
                geomBuffer.data_ptr(), binningBuffer.data_ptr() if binningBuffer.numel() else None,
                binningBuffer.numel(), imageBuffer.data_ptr(), _dev_ptr(dL_dout_color, device, keep), ctypes.byref(grads), stream)
            _lib.check(rc, "rasterize_gaussians_backward")
        return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations,
                dL_dv2g)

    # -- markVisible --------------------------------------------------------------------------
    @staticmethod
    def mark_visible(means3D, viewmatrix, projmatrix):
        if not means3D.is_cuda:
            raise RuntimeError("mark_visible: tensors must be on a CUDA device")
        device = means3D.device
        P = int(means3D.size(0))
        present = torch.zeros((P,), dtype=torch.bool, device=device)
        if P:
            keep: list = []
            with torch.cuda.device(device):
                rc = _lib.lib.gof_mark_visible(P, _dev_ptr(means3D, device, keep), _dev_ptr(viewmatrix, device, keep),
                                               _dev_ptr(projmatrix, device, keep), present.data_ptr(),
                                               _lib.raw_stream(device))
            _lib.check(rc, "mark_visible")
        return present

    # -- integrate ----------------------------------------------------------------------------
    @staticmethod
    def integrate_gaussians_to_points(background, points3D, means3D, colors, opacity, scales, rotations, scale_modifier,
                                      cov3D_precomp, view2gaussian_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                      kernel_size, subpixel_offset, image_height, image_width, sh, degree, campos,
                                      prefiltered, debug):
        """IntegrateGaussiansToPointsCUDA (rasterize_points.cu:234-343): returns (num_rendered, out_color[9,H,W],
        out_alpha_integrated[PN], out_color_integrated[PN,3], radii[P], geomBuffer, binningBuffer, imgBuffer)."""
        if means3D.dim() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if points3D.dim() != 2 or points3D.size(1) != 3:
            raise RuntimeError("points3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise RuntimeError("diff_gof_rasterization (B200): tensors must be on a CUDA device; there is no CPU path")
        device = means3D.device
        P, PN, H, W = int(means3D.size(0)), int(points3D.size(0)), int(image_height), int(image_width)
        M = int(sh.size(1)) if (sh is not None and sh.numel() != 0 and sh.size(0) != 0) else 0
        f32 = dict(dtype=torch.float32, device=device)
        byte_opts = dict(dtype=torch.uint8, device=device)
        out_color = torch.zeros((_lib.OUTPUT_CHANNELS, H, W), **f32)
        radii = torch.zeros((P,), dtype=torch.int32, device=device)
        out_alpha = torch.ones((PN,), **f32)
        out_rgb = torch.zeros((PN, 3), **f32)
        empty = torch.empty(0, **byte_opts)
        if P == 0 or PN == 0:      # rasterize_points.cu:291: nothing is launched
            return NumRendered(0), out_color, out_alpha, out_rgb, radii, empty, empty.clone(), empty.clone()
        keep: list = []
        with torch.cuda.device(device):
            stream = _lib.raw_stream(device)
            gsz, isz, bsz = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
            _lib.check(_lib.lib.gof_state_sizes(P, W, H, 0, ctypes.byref(gsz), ctypes.byref(isz), ctypes.byref(bsz)),
                       "gof_state_sizes")
            geom, img = torch.empty(gsz.value, **byte_opts), torch.empty(isz.value, **byte_opts)
            holder = {}

            def _alloc(_user, nbytes):
                holder["binning"] = torch.empty(int(nbytes), **byte_opts)
                return holder["binning"].data_ptr()

            prm = _lib.GofParams(P, int(degree), M, W, H, float(tan_fovx), float(tan_fovy), float(kernel_size),
                                 float(scale_modifier), int(bool(prefiltered)), int(bool(debug)), _lib.default_flags())
            inp = _lib.GofInputs(
                _dev_ptr(background, device, keep), _dev_ptr(means3D, device, keep), _dev_ptr(sh, device, keep),
                _dev_ptr(colors, device, keep), _dev_ptr(opacity, device, keep), _dev_ptr(scales, device, keep),
                _dev_ptr(rotations, device, keep), _dev_ptr(cov3D_precomp, device, keep),
                _dev_ptr(view2gaussian_precomp, device, keep), _dev_ptr(viewmatrix, device, keep),
                _dev_ptr(projmatrix, device, keep), _dev_ptr(campos, device, keep))
            R = ctypes.c_int32(0)
            rc = _lib.lib.gof_integrate(_lib.context(device.index), ctypes.byref(prm), ctypes.byref(inp), PN,
                                        _dev_ptr(points3D, device, keep), geom.data_ptr(), geom.numel(), img.data_ptr(),
                                        img.numel(), _lib.ALLOC_FN(_alloc), None, out_color.data_ptr(), radii.data_ptr(),
                                        out_alpha.data_ptr(), out_rgb.data_ptr(), ctypes.byref(R), stream)
            _lib.check(rc, "integrate_gaussians_to_points")
        return (NumRendered(R.value), out_color, out_alpha, out_rgb, radii, geom, holder.get("binning", empty), img)


_C = _CModule()


class BatchWorkspace:
    """Reusable device buffers for `rasterize_views` (the batched entry point, gof_forward_batch).

    With a workspace the forward is SYNC-FREE: the binning blob is a cached buffer (capacity = 1.5x
    the largest batch seen so far) and `num_rendered` stays on the device until `finish()` reads the
    mailbox -- call it after your own synchronisation point (e.g. after the D2H copy of the frames).
    `finish()` returns the per-view R, or None if the blob was too small for this batch.  In that case
    every output of the batch (out_color, the frame sink) holds NaN -- never a plausible image --, the
    workspace has been grown, and the CALLER re-runs the batch (cycle.render_scene_views and
    HostFrameSink users do; `rasterize_views` itself cannot, it does not synchronise)."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.geom = self.img = self.binning = None
        self.key = None
        self.capacity_hint = 0
        self.mailbox_host = None

    def _ensure(self, P, W, H, V):
        gsz, isz, bsz = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        if self.capacity_hint == 0:
            self.capacity_hint = max(4 * P * V, 1 << 16)      # first call: generous (4 tiles per Gaussian)
        _lib.check(_lib.lib.gof_state_sizes_batch(P, W, H, V, self.capacity_hint, ctypes.byref(gsz), ctypes.byref(isz),
                                                  ctypes.byref(bsz)), "gof_state_sizes_batch")
        opts = dict(dtype=torch.uint8, device=self.device)
        if self.key != (P, W, H, V) or self.geom is None:
            self.geom = torch.empty(gsz.value, **opts)
            self.img = torch.empty(isz.value, **opts)
            self.key = (P, W, H, V)
        if self.binning is None or self.binning.numel() < bsz.value:
            self.binning = torch.empty(bsz.value, **opts)

    def finish_async(self):
        """Non-blocking `finish`: enqueue the mailbox copy on the current stream (gof_num_rendered_async); read the
        result with `finish_poll()` once an event recorded behind this call has completed."""
        P, W, H, V = self.key
        if self.mailbox_host is None or self.mailbox_host.numel() != 4 + V:
            self.mailbox_host = torch.zeros(4 + V, dtype=torch.int32).pin_memory()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.gof_num_rendered_async(self.geom.data_ptr(), P, V, self.mailbox_host.data_ptr(),
                                                       _lib.raw_stream(self.device)), "gof_num_rendered_async")

    def finish_poll(self):
        """Per-view R of the batch `finish_async` was enqueued behind, or None on overflow (the workspace is grown)."""
        m = self.mailbox_host.tolist()
        R = m[4:]
        if m[1]:
            self.capacity_hint = int(1.5 * m[0]) + 1024
            self.binning = None
            return None
        if 1.25 * sum(R) > self.capacity_hint:
            self.capacity_hint = int(1.5 * sum(R)) + 1024
        return R

    def finish(self):
        P, W, H, V = self.key
        out = (ctypes.c_int32 * V)()
        with torch.cuda.device(self.device):
            rc = _lib.lib.gof_num_rendered(_lib.context(self.device.index), self.geom.data_ptr(), P, V,
                                           torch.cuda.current_stream(self.device).cuda_stream, out)
        R = [int(x) for x in out]
        if rc == _lib.GOF_EOVERFLOW:
            self.capacity_hint = int(1.5 * sum(R)) + 1024
            self.binning = None
            return None
        _lib.check(rc, "gof_num_rendered")
        if 1.25 * sum(R) > self.capacity_hint:                 # keep some head-room for the next batch
            self.capacity_hint = int(1.5 * sum(R)) + 1024
        return R


def rasterize_views(background, means3D, colors, opacity, scales, rotations, scale_modifier, viewmatrices, projmatrices,
                    tan_fovx, tan_fovy, kernel_size, image_height, image_width, sh, degree, campos, prefiltered=False,
                    debug=False, workspace: BatchWorkspace | None = None, out_color=None, sink=None):
    """V views of one Gaussian set in ONE pass of the pipeline (gof_forward_batch).

    viewmatrices/projmatrices: [V,4,4] (or [V,1,4,4]), campos: [V,3]; background [3] or [V,3].
    Returns (num_rendered, color[V,9,H,W], radii[V,P], geom, binning, img).  Frame v is bit-identical
    to `_C.rasterize_gaussians` with camera v.  Without a workspace num_rendered is a list of ints and
    one host synchronisation happens (as in the reference); with one, it is None until
    `workspace.finish()`.  `sink`: optional float32 [V,5,H,W] tensor (contiguous or channels_last) in device or PINNED
    host memory that the blend kernel additionally fills with rgb / median depth / alpha (gof_set_frame_sink) -- with
    pinned memory the frames reach the host while the kernel runs (valid once the stream has been synchronised)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("diff_gof_rasterization (B200): tensors must be on a CUDA device; there is no CPU path")
    device = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    V = int(viewmatrices.reshape(-1, 16).size(0))
    M = int(sh.size(1)) if (sh is not None and sh.numel() != 0 and sh.size(0) != 0) else 0
    keep: list = []
    f32 = dict(dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        stream = _lib.raw_stream(device)
        if out_color is None:
            out_color = torch.empty((V, _lib.OUTPUT_CHANNELS, H, W), **f32)
        if P == 0:
            out_color.zero_()
            if sink is not None:
                sink.zero_()
            e = torch.empty(0, dtype=torch.uint8, device=device)
            return [0] * V, out_color, torch.zeros((V, 0), dtype=torch.int32, device=device), e, e.clone(), e.clone()
        radii = torch.empty((V, P), dtype=torch.int32, device=device)
        bg = background.reshape(-1)
        bg_stride = 3 if bg.numel() == 3 * V and V > 1 else 0
        prm = _lib.GofParams(P, int(degree), M, W, H, float(tan_fovx), float(tan_fovy), float(kernel_size),
                             float(scale_modifier), int(bool(prefiltered)), int(bool(debug)),
                             _lib.default_flags(means3D, colors, opacity, scales, rotations, sh))
        inp = _lib.GofInputs(
            _dev_ptr(bg, device, keep), _dev_ptr(means3D, device, keep), _dev_ptr(sh, device, keep),
            _dev_ptr(colors, device, keep), _dev_ptr(opacity, device, keep), _dev_ptr(scales, device, keep),
            _dev_ptr(rotations, device, keep), None, None, _dev_ptr(viewmatrices.reshape(V, 16), device, keep),
            _dev_ptr(projmatrices.reshape(V, 16), device, keep), _dev_ptr(campos.reshape(V, 3), device, keep))
        ctx = _lib.context(device.index)
        armed = False
        if sink is not None:
            if sink.dtype != torch.float32 or tuple(sink.shape) != (V, _lib.SINK_CHANNELS, H, W) \
                    or not (sink.is_cuda or sink.is_pinned()):
                raise RuntimeError(f"sink must be a float32 [{V},{_lib.SINK_CHANNELS},{H},{W}] tensor in device or pinned host memory")
            if sink.is_contiguous():
                layout = _lib.SINK_CHW
            elif sink.is_contiguous(memory_format=torch.channels_last):
                layout = _lib.SINK_HWC
            else:
                raise RuntimeError("sink must be contiguous (NCHW) or channels_last")
            _lib.check(_lib.lib.gof_set_frame_sink(ctx, sink.data_ptr(), sink.numel() * 4, layout), "gof_set_frame_sink")
            armed = True
        try:
            return _rasterize_views_call(ctx, prm, inp, V, bg_stride, P, W, H, device, stream, out_color, radii, workspace)
        finally:
            # the sink is one-shot and consumed by gof_forward_batch; if anything raised before the library got
            # there (allocation, argument errors), disarm it so that no later forward writes through a stale pointer
            if armed:
                _lib.lib.gof_set_frame_sink(ctx, None, 0, 0)


def _rasterize_views_call(ctx, prm, inp, V, bg_stride, P, W, H, device, stream, out_color, radii, workspace):
    Rv = (ctypes.c_int32 * V)()
    bin_out = ctypes.c_void_p()
    if workspace is None:
        gsz, isz, bsz = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(_lib.lib.gof_state_sizes_batch(P, W, H, V, 0, ctypes.byref(gsz), ctypes.byref(isz),
                                                  ctypes.byref(bsz)), "gof_state_sizes_batch")
        byte_opts = dict(dtype=torch.uint8, device=device)
        geom, img = torch.empty(gsz.value, **byte_opts), torch.empty(isz.value, **byte_opts)
        holder = {}

        def _alloc(_user, nbytes):
            holder["binning"] = torch.empty(int(nbytes), **byte_opts)
            return holder["binning"].data_ptr()

        rc = _lib.lib.gof_forward_batch(ctx, ctypes.byref(prm), ctypes.byref(inp), V, bg_stride, geom.data_ptr(),
                                        geom.numel(), img.data_ptr(), img.numel(), None, 0, _lib.ALLOC_FN(_alloc),
                                        None, out_color.data_ptr(), radii.data_ptr(), Rv, ctypes.byref(bin_out),
                                        stream)
        _lib.check(rc, "rasterize_views")
        binning = holder.get("binning", torch.empty(0, **byte_opts))
        return [int(x) for x in Rv], out_color, radii, geom, binning, img
    workspace._ensure(P, W, H, V)
    rc = _lib.lib.gof_forward_batch(ctx, ctypes.byref(prm), ctypes.byref(inp), V, bg_stride,
                                    workspace.geom.data_ptr(), workspace.geom.numel(), workspace.img.data_ptr(),
                                    workspace.img.numel(), workspace.binning.data_ptr(), workspace.binning.numel(),
                                    _lib.ALLOC_FN(), None, out_color.data_ptr(), radii.data_ptr(), Rv,
                                    ctypes.byref(bin_out), stream)
    _lib.check(rc, "rasterize_views")
    return None, out_color, radii, workspace.geom, workspace.binning, workspace.img


class _RasterizeViews(torch.autograd.Function):
    """Differentiable batched rasterisation: V views of one Gaussian set, forward through gof_forward_batch and
    backward through gof_backward_batch (gradients summed over the views inside the kernels)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cfg):
        (bg, viewmatrices, projmatrices, campos, tanfovx, tanfovy, kernel_size, scale_modifier, H, W, sh_degree) = cfg
        R, color, radii, geom, binning, img = rasterize_views(
            bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier, viewmatrices, projmatrices,
            tanfovx, tanfovy, kernel_size, H, W, sh, sh_degree, campos)
        ctx.cfg = cfg
        ctx.num_rendered = int(sum(R))
        ctx.save_for_backward(means3D, sh, colors_precomp, scales, rotations, radii, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        (bg, viewmatrices, projmatrices, campos, tanfovx, tanfovy, kernel_size, scale_modifier, H, W, sh_degree) = ctx.cfg
        means3D, sh, colors_precomp, scales, rotations, radii, geom, binning, img = ctx.saved_tensors
        device = means3D.device
        P = int(means3D.size(0))
        V = int(radii.size(0))
        M = int(sh.size(1)) if (sh is not None and sh.numel() != 0) else 0
        e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
        g2d, gcol, gop, gm3, gcov, gsh, gsc, grot, gv2g = (e(P, 3), e(P, 3), e(P, 1), e(P, 3), e(P, 6), e(P, M, 3),
                                                           e(P, 3), e(P, 4), e(P, 10))
        keep: list = []
        bgf = bg.reshape(-1)
        with torch.cuda.device(device):
            prm = _lib.GofParams(P, int(sh_degree), M, int(W), int(H), float(tanfovx), float(tanfovy), float(kernel_size),
                                 float(scale_modifier), 0, 0, _lib.default_flags())
            inp = _lib.GofInputs(
                _dev_ptr(bgf, device, keep), _dev_ptr(means3D, device, keep), _dev_ptr(sh, device, keep),
                _dev_ptr(colors_precomp, device, keep), None, _dev_ptr(scales, device, keep),
                _dev_ptr(rotations, device, keep), None, None, _dev_ptr(viewmatrices.reshape(V, 16), device, keep),
                _dev_ptr(projmatrices.reshape(V, 16), device, keep), _dev_ptr(campos.reshape(V, 3), device, keep))
            grads = _lib.GofGrads(g2d.data_ptr(), gcol.data_ptr(), gop.data_ptr(), gm3.data_ptr(), gcov.data_ptr(),
                                  gsh.data_ptr() if M > 0 else None, gsc.data_ptr(), grot.data_ptr(), gv2g.data_ptr())
            rc = _lib.lib.gof_backward_batch(
                _lib.context(device.index), ctypes.byref(prm), ctypes.byref(inp), V,
                3 if (bgf.numel() == 3 * V and V > 1) else 0, ctx.num_rendered, _dev_ptr(radii, device, keep, torch.int32),
                geom.data_ptr(), binning.data_ptr() if binning.numel() else None, binning.numel(), img.data_ptr(),
                _dev_ptr(grad_color, device, keep), ctypes.byref(grads), _lib.raw_stream(device))
            _lib.check(rc, "rasterize_views backward")
        has = lambda t: t is not None and t.numel() != 0
        return (gm3, g2d, gsh if has(sh) else None, gcol if has(colors_precomp) else None, gop.reshape(-1, 1),
                gsc if has(scales) else None, grot if has(rotations) else None, None)


def rasterize_views_autograd(means3D, means2D, opacities, *, shs=None, colors_precomp=None, scales, rotations, bg,
                             viewmatrices, projmatrices, campos, tanfovx, tanfovy, image_height, image_width, sh_degree,
                             kernel_size=0.0, scale_modifier=1.0):
    """Differentiable `rasterize_views`: returns (color[V,9,H,W], radii[V,P]); gradients w.r.t. means3D, means2D (the
    screen-space densification statistic), shs / colors_precomp, opacities, scales, rotations are the sums over the
    V views -- what autograd accumulates when the reference renders the views one call at a time."""
    if (shs is None) == (colors_precomp is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    fill = lambda t: _absent() if t is None else t
    cfg = (bg, viewmatrices, projmatrices, campos, tanfovx, tanfovy, kernel_size, scale_modifier, int(image_height),
           int(image_width), int(sh_degree))
    return _RasterizeViews.apply(means3D, means2D, fill(shs), fill(colors_precomp), opacities, scales, rotations, cfg)


def state_array_batch(name: str, P: int, W: int, H: int, V: int, R: int, geom, binning, img):
    """Test accessor on the state of a V-view batch (gof_state_get_batch)."""
    device = geom.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    spec = {
        "depths": (torch.float32, (V, P)), "means2D": (torch.float32, (V, P, 2)),
        "conic_opacity": (torch.float32, (V, P, 4)), "view2gaussian": (torch.float32, (V, P, 10)),
        "rgb": (torch.float32, (V, P, 3)), "clamped": (torch.uint8, (V, P, 3)),
        "tiles_touched": (torch.int32, (V, P)), "point_offsets": (torch.int32, (V, P)),
        "final_T": (torch.float32, (V, 4, H, W)), "n_contrib": (torch.int32, (V, 2, H, W)),
        "ranges": (torch.int32, (V * T, 2)), "point_list": (torch.int32, (R,)), "point_list_keys": (torch.int64, (R,)),
    }[name]
    out = torch.empty(spec[1], dtype=spec[0], device=device)
    with torch.cuda.device(device):
        n = _lib.lib.gof_state_get_batch(name.encode(), P, W, H, V, R, geom.data_ptr(),
                                         binning.data_ptr() if binning.numel() else None, binning.numel(), img.data_ptr(),
                                         out.data_ptr() if out.numel() else None, out.numel() * out.element_size(),
                                         _lib.raw_stream(device))
    if n < 0:
        raise RuntimeError(f"gof_state_get_batch({name}) failed: {_lib.last_error()}")
    return out


def backward_accumulators(device, V: int, P: int):
    """Test accessor (gof_backward_accumulators): [V,P,20] per-view gradient accumulators of the last backward."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((V, P, 20), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        n = _lib.lib.gof_backward_accumulators(_lib.context(device.index), out.data_ptr(), out.numel() * 4,
                                               _lib.raw_stream(device))
    if n != out.numel() * 4:
        raise RuntimeError(f"gof_backward_accumulators: expected {out.numel() * 4} bytes, got {n}: {_lib.last_error()}")
    return out


def state_array(name: str, P: int, W: int, H: int, R: int, geom, binning, img):
    """Test accessor: decode one array of the opaque forward state (gof_state_get)."""
    device = geom.device
    spec = {
        "depths": (torch.float32, (P,)), "means2D": (torch.float32, (P, 2)),
        "conic_opacity": (torch.float32, (P, 4)), "view2gaussian": (torch.float32, (P, 10)),
        "rgb": (torch.float32, (P, 3)), "clamped": (torch.uint8, (P, 3)),
        "tiles_touched": (torch.int32, (P,)), "point_offsets": (torch.int32, (P,)),
        "final_T": (torch.float32, (4, H, W)), "n_contrib": (torch.int32, (2, H, W)),
        "ranges": (torch.int32, (((W + 15) // 16) * ((H + 15) // 16), 2)),
        "point_list": (torch.int32, (R,)), "point_list_keys": (torch.int64, (R,)),
    }[name]
    out = torch.empty(spec[1], dtype=spec[0], device=device)
    with torch.cuda.device(device):
        n = _lib.lib.gof_state_get(name.encode(), P, W, H, R, geom.data_ptr(),
                                   binning.data_ptr() if binning.numel() else None, binning.numel(), img.data_ptr(),
                                   out.data_ptr() if out.numel() else None, out.numel() * out.element_size(),
                                   _lib.raw_stream(device))
    if n < 0:
        raise RuntimeError(f"gof_state_get({name}) failed: {_lib.last_error()}")
    return out


def preprocess_backward_stage(c_means3D, radii, sh, scales, rotations, viewmatrix, campos, degree, geom,
                              dL_dview2gaussian, dL_dcolors):
    """Stage accessor (gof_preprocess_backward): K10 alone on caller-provided dL/dview2gaussian and
    dL/dcolor.  Returns (dL_dmeans3D, dL_dsh, dL_dscales, dL_drotations)."""
    device = c_means3D.device
    P = int(c_means3D.size(0))
    M = int(sh.size(1)) if (sh is not None and sh.numel() != 0) else 0
    e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
    g2d, gcol, gop, gm3, gcov, gsh, gsc, grot, gv2g = (e(P, 3), e(P, 3), e(P, 1), e(P, 3), e(P, 6), e(P, M, 3),
                                                       e(P, 3), e(P, 4), e(P, 10))
    keep: list = []
    with torch.cuda.device(device):
        prm = _lib.GofParams(P, int(degree), M, 16, 16, 1.0, 1.0, 0.0, 1.0, 0, 0, 0)
        inp = _lib.GofInputs(None, _dev_ptr(c_means3D, device, keep), _dev_ptr(sh, device, keep), None, None,
                             _dev_ptr(scales, device, keep), _dev_ptr(rotations, device, keep), None, None,
                             _dev_ptr(viewmatrix, device, keep), None, _dev_ptr(campos, device, keep))
        grads = _lib.GofGrads(g2d.data_ptr(), gcol.data_ptr(), gop.data_ptr(), gm3.data_ptr(), gcov.data_ptr(),
                              gsh.data_ptr() if M > 0 else None, gsc.data_ptr(), grot.data_ptr(), gv2g.data_ptr())
        rc = _lib.lib.gof_preprocess_backward(
            _lib.context(device.index), ctypes.byref(prm), ctypes.byref(inp), _dev_ptr(radii, device, keep, torch.int32),
            geom.data_ptr(), _dev_ptr(dL_dview2gaussian, device, keep), _dev_ptr(dL_dcolors, device, keep),
            ctypes.byref(grads), _lib.raw_stream(device))
        _lib.check(rc, "gof_preprocess_backward")
    return gm3, gsh, gsc, grot


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        view2gaussian_precomp, raster_settings):
    if not torch.is_grad_enabled() and not raster_settings.debug:
        # inference (the render loops run under no_grad): same call without the autograd.Function round trip
        rs = raster_settings
        _, color, radii, _, _, _ = _C.rasterize_gaussians(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            view2gaussian_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.kernel_size,
            rs.subpixel_offset, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        return color, radii
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, view2gaussian_precomp, raster_settings)


def _cpu_copy(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


class _RasterizeGaussians(torch.autograd.Function):
    """autograd bridge (reference: RAST/diff_gof_rasterization/__init__.py:46-165)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                view2gaussian_precomp, raster_settings):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                view2gaussian_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.kernel_size,
                rs.subpixel_offset, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered,
                rs.debug)
        if rs.debug:
            snapshot = _cpu_copy(args)
            try:
                num_rendered, color, radii, geom, binning, img = _C.rasterize_gaussians(*args)
            except Exception:
                torch.save(snapshot, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            num_rendered, color, radii, geom, binning, img = _C.rasterize_gaussians(*args)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, view2gaussian_precomp,
                              radii, sh, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, view2gaussian_precomp, radii, sh, geom, binning,
         img) = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                view2gaussian_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.kernel_size,
                rs.subpixel_offset, grad_out_color, sh, rs.sh_degree, rs.campos, geom, ctx.num_rendered, binning, img,
                rs.debug)
        if rs.debug:
            snapshot = _cpu_copy(args)
            try:
                out = _C.rasterize_gaussians_backward(*args)
            except Exception:
                torch.save(snapshot, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        else:
            out = _C.rasterize_gaussians_backward(*args)
        (g_means2D, g_colors, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rot, g_v2g) = out
        # forward-argument order; the reference returns a tensor for every slot, even absent inputs
        return (g_means3D, g_means2D, g_sh, g_colors, g_opacity, g_scales, g_rot, g_cov3D, g_v2g, None)


class GaussianRasterizationSettings_GOF(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    kernel_size: float
    subpixel_offset: torch.Tensor
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


_ABSENT = torch.Tensor([])          # the reference's "not provided" sentinel (numel() == 0); never written to


def _absent() -> torch.Tensor:
    return _ABSENT


class GaussianRasterizer_GOF(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    @staticmethod
    def _validate(shs, colors_precomp, scales, rotations, cov3D_precomp):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, view2gaussian_precomp=None):
        self._validate(shs, colors_precomp, scales, rotations, cov3D_precomp)
        fill = lambda t: _absent() if t is None else t
        return rasterize_gaussians(means3D, means2D, fill(shs), fill(colors_precomp), opacities, fill(scales),
                                   fill(rotations), fill(cov3D_precomp), fill(view2gaussian_precomp),
                                   self.raster_settings)

    def integrate(self, points3D, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                  rotations=None, cov3D_precomp=None, view2gaussian_precomp=None):
        """GaussianRasterizer_GOF.integrate (reference :241-307): (color, alpha_integrated, color_integrated, radii)."""
        self._validate(shs, colors_precomp, scales, rotations, cov3D_precomp)
        rs = self.raster_settings
        fill = lambda t: _absent() if t is None else t
        args = (rs.bg, points3D, means3D, fill(colors_precomp), opacities, fill(scales), fill(rotations),
                rs.scale_modifier, fill(cov3D_precomp), fill(view2gaussian_precomp), rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, rs.kernel_size, rs.subpixel_offset, rs.image_height, rs.image_width, fill(shs),
                rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        if rs.debug:
            snapshot = _cpu_copy(args)
            try:
                out = _C.integrate_gaussians_to_points(*args)
            except Exception:
                torch.save(snapshot, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            out = _C.integrate_gaussians_to_points(*args)
        _num_rendered, color, alpha_integrated, color_integrated, radii, _g, _b, _i = out
        return color, alpha_integrated, color_integrated, radii


# Un-suffixed names (commented out in the reference at :167,:184; used by vanilla render()).
GaussianRasterizationSettings = GaussianRasterizationSettings_GOF
GaussianRasterizer = GaussianRasterizer_GOF
