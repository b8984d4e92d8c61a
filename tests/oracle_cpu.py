"""numpy/ctypes driver of the CPU oracle (oracle/gof_oracle.c -> oracle/_ref/libgof_oracle.so).
Test infrastructure only; the product never imports this."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_int64, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libgof_oracle.so")
LIB64 = os.path.join(ROOT, "oracle", "_ref", "libgof_oracle_f64.so")
_lib = None
_lib64 = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "gof_oracle.c")
        if not os.path.exists(LIB) or not os.path.exists(LIB64) or os.path.getmtime(src) > os.path.getmtime(LIB):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
        L = ctypes.CDLL(LIB)
        L.oracle_num_threads.restype = c_int
        L.oracle_scan.restype = c_int64
        L.oracle_binning.restype = c_int
        _lib = L
    return _lib


def lib64():
    """The float64 build of the backward preprocess (conditioning reference)."""
    global _lib64
    if _lib64 is None:
        lib()   # (re)builds both
        _lib64 = ctypes.CDLL(LIB64)
        assert _lib64.oracle_real_bytes() == 8
    return _lib64


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def preprocess(c: dict) -> dict:
    """c: flat case dict of numpy arrays / scalars (see cases.py)."""
    P = c["means3D"].shape[0]
    shs = _f(c.get("shs"))
    M = shs.shape[1] if shs is not None else 0
    out = {
        "radii": np.zeros(P, np.int32), "means2D": np.zeros((P, 2), np.float32), "depths": np.zeros(P, np.float32),
        "view2gaussian": np.zeros((P, 10), np.float32), "rgb": np.zeros((P, 3), np.float32),
        "conic_opacity": np.zeros((P, 4), np.float32), "tiles_touched": np.zeros(P, np.uint32),
        "clamped": np.zeros((P, 3), np.uint8),
    }
    keep = [_f(c["means3D"]), _f(c.get("scales")), _f(c.get("rotations")), _f(c["opacities"]), shs,
            _f(c.get("cov3D_precomp")), _f(c.get("colors_precomp")), _f(c.get("view2gaussian_precomp")),
            _f(c["viewmatrix"]), _f(c["projmatrix"]), _f(c["campos"])]
    lib().oracle_preprocess(
        c_int(P), c_int(int(c["D"])), c_int(M), _p(keep[0]), _p(keep[1]), c_float(c["scale_modifier"]), _p(keep[2]),
        _p(keep[3]), _p(keep[4]), _p(keep[5]), _p(keep[6]), _p(keep[7]), _p(keep[8]), _p(keep[9]), _p(keep[10]),
        c_int(int(c["W"])), c_int(int(c["H"])), c_float(c["tanfovx"]), c_float(c["tanfovy"]), c_float(c["kernel_size"]),
        _p(out["radii"]), _p(out["means2D"]), _p(out["depths"]), _p(out["view2gaussian"]), _p(out["rgb"]),
        _p(out["conic_opacity"]), _p(out["tiles_touched"]), _p(out["clamped"]))
    return out


def binning(W, H, means2D, depths, radii, tiles_touched) -> dict:
    W, H = int(W), int(H)
    P = radii.shape[0]
    tt = np.ascontiguousarray(tiles_touched, np.uint32)
    offs = np.zeros(P, np.uint32)
    R = int(lib().oracle_scan(c_int(P), _p(tt), _p(offs)))
    T = ((W + 15) // 16) * ((H + 15) // 16)
    keys = np.zeros(R, np.uint64)
    plist = np.zeros(R, np.uint32)
    ranges = np.zeros((T, 2), np.uint32)
    m2, d, r = _f(means2D), _f(depths), np.ascontiguousarray(radii, np.int32)
    rc = lib().oracle_binning(c_int(P), c_int(W), c_int(H), _p(m2), _p(d), _p(r), _p(offs), c_int64(R), None, None,
                              _p(keys), _p(plist), _p(ranges))
    assert rc == 0
    return {"num_rendered": R, "point_offsets": offs, "point_list_keys": keys, "point_list": plist, "ranges": ranges}


def render_forward(c, ranges, point_list, v2g, conic_opacity, features) -> dict:
    W, H = int(c["W"]), int(c["H"])
    N = W * H
    out = {"out_color": np.zeros((9, H, W), np.float32), "final_T": np.zeros((4, N), np.float32),
           "n_contrib": np.zeros((2, N), np.uint32)}
    keep = [np.ascontiguousarray(ranges, np.uint32), np.ascontiguousarray(point_list, np.uint32), _f(v2g),
            _f(conic_opacity), _f(features), _f(c["bg"])]
    lib().oracle_render_forward(c_int(W), c_int(H), c_float(c["tanfovx"]), c_float(c["tanfovy"]), _p(keep[0]),
                                _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(keep[5]),
                                _p(out["out_color"]), _p(out["final_T"]), _p(out["n_contrib"]))
    return out


def render_backward(c, ranges, point_list, v2g, conic_opacity, means2D, features, final_T, n_contrib, dL_dout) -> dict:
    W, H = int(c["W"]), int(c["H"])
    P = v2g.shape[0]
    out = {"dL_dmeans2D": np.zeros((P, 3), np.float32), "dL_dopacity": np.zeros((P, 1), np.float32),
           "dL_dcolors": np.zeros((P, 3), np.float32), "dL_dview2gaussian": np.zeros((P, 10), np.float32)}
    keep = [np.ascontiguousarray(ranges, np.uint32), np.ascontiguousarray(point_list, np.uint32), _f(v2g),
            _f(conic_opacity), _f(means2D), _f(features), _f(c["bg"]), _f(final_T),
            np.ascontiguousarray(n_contrib, np.uint32), _f(dL_dout)]
    lib().oracle_render_backward(c_int(P), c_int(W), c_int(H), c_float(c["tanfovx"]), c_float(c["tanfovy"]),
                                 *[_p(k) for k in keep], _p(out["dL_dmeans2D"]), _p(out["dL_dopacity"]),
                                 _p(out["dL_dcolors"]), _p(out["dL_dview2gaussian"]))
    return out


def preprocess_backward(c, radii, clamped, dL_dv2g, dL_dcolor, f64: bool = False) -> dict:
    """f64=True: same formulas, float32 inputs, arithmetic and outputs in double."""
    P = c["means3D"].shape[0]
    shs = _f(c.get("shs"))
    M = shs.shape[1] if shs is not None else 0
    dt = np.float64 if f64 else np.float32
    out = {"dL_dmeans3D": np.zeros((P, 3), dt), "dL_dsh": np.zeros((P, M, 3), dt),
           "dL_dscales": np.zeros((P, 3), dt), "dL_drotations": np.zeros((P, 4), dt)}
    keep = [_f(c["means3D"]), np.ascontiguousarray(radii, np.int32), shs, np.ascontiguousarray(clamped, np.uint8),
            _f(c["scales"]), _f(c["rotations"]), _f(c["viewmatrix"]), _f(c["campos"]), _f(dL_dv2g), _f(dL_dcolor)]
    (lib64() if f64 else lib()).oracle_preprocess_backward(c_int(P), c_int(int(c["D"])), c_int(M), *[_p(k) for k in keep],
                                     _p(out["dL_dmeans3D"]), _p(out["dL_dsh"]) if M else None,
                                     _p(out["dL_dscales"]), _p(out["dL_drotations"]))
    return out


def integrate(c: dict, points3D, ranges, point_list, v2g, conic_opacity, features) -> dict:
    """Rasterizer::integrate on the CPU (oracle_integrate): state arrays as produced by preprocess() + binning()."""
    W, H = int(c["W"]), int(c["H"])
    N = W * H
    pts = _f(points3D)
    PN = pts.shape[0]
    out = {"out_color": np.zeros((9, H, W), np.float32), "final_T": np.zeros(N, np.float32),
           "n_contrib": np.zeros(N, np.uint32), "alpha_integrated": np.ones(PN, np.float32),
           "color_integrated": np.zeros((PN, 3), np.float32)}
    keep = [_f(c["viewmatrix"]), pts, np.ascontiguousarray(ranges, np.uint32), np.ascontiguousarray(point_list, np.uint32),
            _f(v2g), _f(conic_opacity), _f(features), _f(c["bg"])]
    L = lib()
    L.oracle_integrate.restype = c_int
    L.oracle_integrate(c_int(v2g.shape[0]), c_int(PN), c_int(W), c_int(H), c_float(c["tanfovx"]), c_float(c["tanfovy"]),
                       *[_p(k) for k in keep], _p(out["out_color"]), _p(out["final_T"]), _p(out["n_contrib"]),
                       _p(out["alpha_integrated"]), _p(out["color_integrated"]))
    return out


def forward_all(c: dict) -> dict:
    """Whole forward pipeline on the CPU (the cpu_baseline unit of work)."""
    pre = preprocess(c)
    b = binning(c["W"], c["H"], pre["means2D"], pre["depths"], pre["radii"], pre["tiles_touched"])
    feats = pre["rgb"] if c.get("colors_precomp") is None else c["colors_precomp"]
    v2g = pre["view2gaussian"] if c.get("view2gaussian_precomp") is None else c["view2gaussian_precomp"]
    r = render_forward(c, b["ranges"], b["point_list"], v2g, pre["conic_opacity"], feats)
    return {**pre, **b, **r}


def case_to_numpy(c: dict) -> dict:
    import torch
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
