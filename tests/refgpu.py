"""ctypes driver of oracle/_ref/libgof_ref.so -- the UNMODIFIED reference CUDA rasterizer
compiled for sm_100a (oracle/Makefile, oracle/ref_harness.cu).  Test infrastructure only.

`run_reference(case)` / `run_ours(case)` take the same flat `case` dict (see cases.py) and
return dicts of torch tensors with identical keys so that tests can zip over them.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libgof_ref.so")

STATE_SPECS = {
    "depths": (torch.float32, lambda P, N, T, R: (P,)),
    "means2D": (torch.float32, lambda P, N, T, R: (P, 2)),
    "conic_opacity": (torch.float32, lambda P, N, T, R: (P, 4)),
    "view2gaussian": (torch.float32, lambda P, N, T, R: (P, 10)),
    "rgb": (torch.float32, lambda P, N, T, R: (P, 3)),
    "clamped": (torch.uint8, lambda P, N, T, R: (P, 3)),
    "tiles_touched": (torch.int32, lambda P, N, T, R: (P,)),
    "point_offsets": (torch.int32, lambda P, N, T, R: (P,)),
    "final_T": (torch.float32, lambda P, N, T, R: (4, N)),
    "n_contrib": (torch.int32, lambda P, N, T, R: (2, N)),
    "ranges": (torch.int32, lambda P, N, T, R: (T, 2)),
    "point_list": (torch.int32, lambda P, N, T, R: (R,)),
    "point_list_keys": (torch.int64, lambda P, N, T, R: (R,)),
}
GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
              "dL_drotations", "dL_dview2gaussian"]

_ref = None


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def ref_lib():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(REF_LIB)
        lib.ref_last_error.restype = c_char_p
        lib.ref_state_create.restype = c_void_p
        lib.ref_state_destroy.argtypes = [c_void_p]
        lib.ref_forward.restype = c_int
        lib.ref_forward.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int] + [c_void_p] * 5 + \
            [c_float] + [c_void_p] * 6 + [c_float, c_float, c_float, c_void_p, c_int, c_void_p, c_void_p, c_int]
        lib.ref_backward.restype = c_int
        lib.ref_backward.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int] + [c_void_p] * 5 + \
            [c_float] + [c_void_p] * 5 + [c_float, c_float, c_float, c_void_p, c_void_p, c_void_p] + [c_void_p] * 10 + [c_int]
        lib.ref_state_get.restype = c_longlong
        lib.ref_state_get.argtypes = [c_void_p, c_char_p, c_void_p, c_longlong]
        lib.ref_integrate.restype = c_int
        lib.ref_integrate.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int] + [c_void_p] * 6 + \
            [c_float] + [c_void_p] * 6 + [c_float, c_float, c_float, c_void_p, c_int] + [c_void_p] * 4 + [c_int]
        lib.ref_mark_visible.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        _ref = lib
    return _ref


def _p(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


class RefRun:
    """One forward (+ optional backward) of the reference; keeps the opaque state alive."""

    def __init__(self):
        self.lib = ref_lib()
        self.state = self.lib.ref_state_create()

    def __del__(self):
        try:
            self.lib.ref_state_destroy(self.state)
        except Exception:
            pass

    def forward(self, c: dict, decode_state: bool = True) -> dict:
        dev = c["means3D"].device
        P, W, H = c["means3D"].shape[0], c["W"], c["H"]
        M = c["shs"].shape[1] if c.get("shs") is not None else 0
        self.out_color = torch.zeros((9, H, W), dtype=torch.float32, device=dev)
        self.radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        R = self.lib.ref_forward(self.state, P, c["D"], M, _p(c["bg"]), W, H, _p(c["means3D"]), _p(c.get("shs")),
                                 _p(c.get("colors_precomp")), _p(c["opacities"]), _p(c.get("scales")),
                                 c["scale_modifier"], _p(c.get("rotations")), _p(c.get("cov3D_precomp")),
                                 _p(c.get("view2gaussian_precomp")), _p(c["viewmatrix"]), _p(c["projmatrix"]),
                                 _p(c["campos"]), c["tanfovx"], c["tanfovy"], c["kernel_size"], None, 0,
                                 self.out_color.data_ptr(), self.radii.data_ptr(), 0)
        if R < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        torch.cuda.synchronize()
        self.R, self.P, self.W, self.H, self.M = R, P, W, H, M
        out = {"num_rendered": R, "out_color": self.out_color, "radii": self.radii}
        if decode_state:
            N, T = W * H, ((W + 15) // 16) * ((H + 15) // 16)
            for name, (dt, shp) in STATE_SPECS.items():
                if R == 0 and name in ("point_list", "point_list_keys"):
                    out[name] = torch.empty(0, dtype=dt, device=dev)
                    continue
                t = torch.empty(shp(P, N, T, R), dtype=dt, device=dev)
                n = self.lib.ref_state_get(self.state, name.encode(), _p(t), t.numel() * t.element_size())
                if n < 0:
                    raise RuntimeError(self.lib.ref_last_error().decode())
                out[name] = t
            torch.cuda.synchronize()
        return out

    def backward(self, c: dict, dL_dout: torch.Tensor) -> dict:
        dev = c["means3D"].device
        P, M = self.P, self.M
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        g = {"dL_dmeans2D": z(P, 3), "dL_dcolors": z(P, 3), "dL_dopacity": z(P, 1), "dL_dmeans3D": z(P, 3),
             "dL_dcov3D": z(P, 6), "dL_dsh": z(P, M, 3), "dL_dscales": z(P, 3), "dL_drotations": z(P, 4),
             "dL_dview2gaussian": z(P, 10)}
        dconic = z(P, 2, 2)
        dL = dL_dout.contiguous()
        torch.cuda.synchronize()
        rc = self.lib.ref_backward(self.state, P, c["D"], M, self.R, _p(c["bg"]), self.W, self.H, _p(c["means3D"]),
                                   _p(c.get("shs")), _p(c.get("colors_precomp")), _p(c.get("view2gaussian_precomp")),
                                   _p(c.get("scales")), c["scale_modifier"], _p(c.get("rotations")),
                                   _p(c.get("cov3D_precomp")), _p(c["viewmatrix"]), _p(c["projmatrix"]), _p(c["campos"]),
                                   c["tanfovx"], c["tanfovy"], c["kernel_size"], None, self.radii.data_ptr(),
                                   dL.data_ptr(), g["dL_dmeans2D"].data_ptr(), dconic.data_ptr(),
                                   g["dL_dopacity"].data_ptr(), g["dL_dcolors"].data_ptr(), g["dL_dmeans3D"].data_ptr(),
                                   g["dL_dcov3D"].data_ptr(), _p(g["dL_dsh"]), g["dL_dscales"].data_ptr(),
                                   g["dL_drotations"].data_ptr(), g["dL_dview2gaussian"].data_ptr(), 0)
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        torch.cuda.synchronize()
        return g


def ref_integrate(c: dict, points3D: torch.Tensor) -> dict:
    """The reference's Rasterizer::integrate on case `c` and query points [PN,3]."""
    lib = ref_lib()
    st = lib.ref_state_create()
    dev = c["means3D"].device
    P, W, H, PN = c["means3D"].shape[0], c["W"], c["H"], points3D.shape[0]
    M = c["shs"].shape[1] if c.get("shs") is not None else 0
    out_color = torch.zeros((9, H, W), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    alpha = torch.ones((PN,), dtype=torch.float32, device=dev)
    rgb = torch.zeros((PN, 3), dtype=torch.float32, device=dev)
    sub = torch.zeros((H, W, 2), dtype=torch.float32, device=dev)
    pts = points3D.contiguous()
    torch.cuda.synchronize()
    R = lib.ref_integrate(st, PN, P, c["D"], M, _p(c["bg"]), W, H, _p(pts), _p(c["means3D"]), _p(c.get("shs")),
                          _p(c.get("colors_precomp")), _p(c["opacities"]), _p(c.get("scales")), c["scale_modifier"],
                          _p(c.get("rotations")), _p(c.get("cov3D_precomp")), _p(c.get("view2gaussian_precomp")),
                          _p(c["viewmatrix"]), _p(c["projmatrix"]), _p(c["campos"]), c["tanfovx"], c["tanfovy"],
                          c["kernel_size"], sub.data_ptr(), 0, out_color.data_ptr(), radii.data_ptr(), alpha.data_ptr(),
                          rgb.data_ptr(), 0)
    torch.cuda.synchronize()
    if R < 0:
        raise RuntimeError(lib.ref_last_error().decode())
    lib.ref_state_destroy(st)
    return {"num_rendered": R, "out_color": out_color, "radii": radii, "alpha_integrated": alpha, "color_integrated": rgb}


def ours_integrate(c: dict, points3D: torch.Tensor) -> dict:
    from f3d_gaus_b200.diff_gof_rasterization import _C
    e = torch.Tensor([])
    g = lambda k: c[k] if c.get(k) is not None else e
    R, color, alpha, rgb, radii, *_ = _C.integrate_gaussians_to_points(
        c["bg"], points3D, c["means3D"], g("colors_precomp"), c["opacities"], g("scales"), g("rotations"),
        c["scale_modifier"], g("cov3D_precomp"), g("view2gaussian_precomp"), c["viewmatrix"], c["projmatrix"],
        c["tanfovx"], c["tanfovy"], c["kernel_size"], e, c["H"], c["W"], g("shs"), c["D"], c["campos"], False, False)
    torch.cuda.synchronize()
    return {"num_rendered": int(R), "out_color": color, "radii": radii, "alpha_integrated": alpha, "color_integrated": rgb}


class OursRun:
    """Same interface on top of the product's public `_C` surface (C ABI underneath)."""

    def forward(self, c: dict, decode_state: bool = True) -> dict:
        from f3d_gaus_b200.diff_gof_rasterization import _C, state_array
        e = torch.Tensor([])
        g = lambda k: c[k] if c.get(k) is not None else e
        self.args = c
        R, color, radii, geom, binning, img = _C.rasterize_gaussians(
            c["bg"], c["means3D"], g("colors_precomp"), c["opacities"], g("scales"), g("rotations"),
            c["scale_modifier"], g("cov3D_precomp"), g("view2gaussian_precomp"), c["viewmatrix"], c["projmatrix"],
            c["tanfovx"], c["tanfovy"], c["kernel_size"], e, c["H"], c["W"], g("shs"), c["D"], c["campos"], False,
            bool(c.get("debug", False)))
        self.saved = (int(R), radii, geom, binning, img)
        out = {"num_rendered": int(R), "out_color": color, "radii": radii}
        if decode_state:
            P, W, H = c["means3D"].shape[0], c["W"], c["H"]
            for name in STATE_SPECS:
                t = state_array(name, P, W, H, int(R), geom, binning, img)
                if name in ("final_T", "n_contrib"):
                    t = t.reshape(t.shape[0], -1)
                out[name] = t
            torch.cuda.synchronize()
        return out

    def backward(self, c: dict, dL_dout: torch.Tensor) -> dict:
        from f3d_gaus_b200.diff_gof_rasterization import _C
        e = torch.Tensor([])
        g = lambda k: c[k] if c.get(k) is not None else e
        R, radii, geom, binning, img = self.saved
        outs = _C.rasterize_gaussians_backward(
            c["bg"], c["means3D"], radii, g("colors_precomp"), g("scales"), g("rotations"), c["scale_modifier"],
            g("cov3D_precomp"), g("view2gaussian_precomp"), c["viewmatrix"], c["projmatrix"], c["tanfovx"],
            c["tanfovy"], c["kernel_size"], e, dL_dout, g("shs"), c["D"], c["campos"], geom, R, binning, img,
            bool(c.get("debug", False)))
        torch.cuda.synchronize()
        return dict(zip(GRAD_NAMES, outs))
