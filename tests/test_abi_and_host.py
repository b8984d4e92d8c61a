"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/gof_b200.h
declares; host-side logic (cameras, synthetic clouds, argument validation, state sizes)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from f3d_gaus_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gof_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(gof_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in gof_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"


def test_host_only_calls():
    from f3d_gaus_b200 import _lib
    assert b"sm_100a" in _lib.lib.gof_version()
    g, i, b = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    assert _lib.lib.gof_state_sizes(65536, 256, 256, 200000, ctypes.byref(g), ctypes.byref(i), ctypes.byref(b)) == 0
    assert g.value > 65536 * (4 + 8 + 16 + 64 + 4 + 4 + 3)
    assert i.value > 256 * 256 * 24
    assert b.value > 200000 * (8 + 8 + 4 + 4 + 64)
    assert _lib.lib.gof_state_sizes(-1, 256, 256, 0, None, None, None) == _lib.GOF_EINVAL
    assert "bad sizes" in _lib.last_error()


def test_missing_library_fails_loudly(tmp_path):
    from f3d_gaus_b200 import _lib
    with pytest.raises(ImportError):
        _lib.load(str(tmp_path / "nope.so"))


def test_no_cpu_path():
    from f3d_gaus_b200.diff_gof_rasterization import _C
    e = torch.Tensor([])
    with pytest.raises(RuntimeError):
        _C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 3), e, torch.zeros(4, 1), torch.zeros(4, 3),
                               torch.zeros(4, 4), 1.0, e, e, torch.eye(4), torch.eye(4), 0.5, 0.5, 0.0, e, 32, 32,
                               torch.zeros(4, 1, 3), 0, torch.zeros(3), False, False)
    with pytest.raises(RuntimeError):
        _C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 2), e, e, e, e, 1.0, e, e, e, e, 0.5, 0.5, 0.0, e, 32, 32,
                               e, 0, e, False, False)


def test_rasterizer_argument_validation():
    from f3d_gaus_b200.diff_gof_rasterization import (GaussianRasterizationSettings, GaussianRasterizationSettings_GOF,
                                                     GaussianRasterizer, GaussianRasterizer_GOF)
    assert GaussianRasterizationSettings is GaussianRasterizationSettings_GOF
    assert GaussianRasterizer is GaussianRasterizer_GOF
    rs = GaussianRasterizationSettings_GOF(32, 32, 0.5, 0.5, 0.0, torch.zeros(1), torch.zeros(3), 1.0, torch.eye(4),
                                           torch.eye(4), 1, torch.zeros(3), False, False)
    r = GaussianRasterizer_GOF(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(x, x, torch.zeros(4, 1), shs=None, colors_precomp=None, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(x, x, torch.zeros(4, 1), shs=torch.zeros(4, 4, 3), scales=x, rotations=None)
    with pytest.raises(RuntimeError, match="CUDA device"):     # integrate exists; like every entry point it has no CPU path
        r.integrate(x, x, x, torch.zeros(4, 1), shs=torch.zeros(4, 4, 3), scales=x, rotations=torch.zeros(4, 4))


def test_install_drop_in():
    import f3d_gaus_b200
    f3d_gaus_b200.install_drop_in()
    from diff_gof_rasterization import GaussianRasterizationSettings_GOF, GaussianRasterizer_GOF, _C  # noqa: F401
    assert all(hasattr(_C, n) for n in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible",
                                        "integrate_gaussians_to_points"))


def test_known_answer_cameras():
    """canonical full_proj and novel view #2 of 8 computed from the reference code."""
    from f3d_gaus_b200 import cameras
    c = cameras.canonical_camera()
    assert torch.allclose(c.world_view[0], torch.eye(4), atol=1e-6)
    assert torch.allclose(c.centers[0], torch.zeros(3), atol=1e-6)
    want = torch.tensor([[8.666594, 0, 0, 0], [0, 8.666594, 0, 0], [0, 0, 7.667, 1], [0, 0, -28.891445, 0]])
    assert torch.allclose(c.full_proj[0], want, atol=2e-5)
    o = cameras.orbit_cameras(8)
    wv2 = torch.tensor([[0.970444, 0, 0.241326, 0], [-0.008054, 0.999443, 0.032386, 0],
                        [-0.241192, -0.033372, 0.969904, 0], [1.849216, 0.255863, 0.230749, 1]])
    assert torch.allclose(o.world_view[2], wv2, atol=2e-6)
    assert torch.allclose(o.centers[2], torch.tensor([-1.850246, -0.248301, 0.230749]), atol=2e-6)
    # first and last orbit views coincide (steps 0 and 1 of the circle)
    assert torch.allclose(o.world_view[0], o.world_view[7], atol=1e-5)
    assert abs(math.tan(13.164 * math.pi / 360) - 0.11538559) < 1e-7


def test_synthetic_is_seeded_and_shaped():
    from f3d_gaus_b200 import synthetic
    a, b, c = synthetic.f3d_like(3, 32), synthetic.f3d_like(3, 32), synthetic.f3d_like(4, 32)
    for k, shape in (("xyz", (1, 1024, 3)), ("opacity", (1, 1024, 1)), ("scaling", (1, 1024, 3)),
                     ("rotation", (1, 1024, 4)), ("features_dc", (1, 1024, 1, 3)), ("features_rest", (1, 1024, 3, 3))):
        assert tuple(a[k].shape) == shape
        assert torch.equal(a[k], b[k])
    assert not torch.equal(a["xyz"], c["xyz"])
    z = a["xyz"][0, :, 2]
    assert 6.667 <= float(z.min()) and float(z.max()) <= 8.667
    assert torch.allclose(a["rotation"].norm(dim=-1), torch.ones(1, 1024), atol=1e-5)
    m = synthetic.concat_sets([a, c])
    assert m["xyz"].shape == (1, 2048, 3)


def test_ply_export_roundtrip(tmp_path):
    """ply.flat_attributes reproduces load_ply(path=None) of visualize.py:146-179; save_ply/read_ply round-trip."""
    from f3d_gaus_b200 import ply, synthetic
    pc = synthetic.f3d_like(3, 16)
    xyz, f_dc, f_rest, opac, scale, rot = ply.flat_attributes(pc, 0)
    P = 256
    assert xyz.shape == (P, 3) and f_dc.shape == (P, 3) and f_rest.shape == (P, 45) and opac.shape == (P, 1)
    assert torch.equal(f_dc, pc["features_dc"][0].reshape(P, 3)) and not f_rest.any()
    path = str(tmp_path / "sub" / "scene.ply")
    assert ply.save_ply(pc, 0, path) == P
    back = ply.read_ply(path)
    assert list(back)[:6] == ["x", "y", "z", "nx", "ny", "nz"] and len(back) == 6 + 3 + 45 + 1 + 3 + 4
    assert np.array_equal(back["x"], xyz[:, 0].numpy()) and np.array_equal(back["rot_3"], rot[:, 3].numpy())
    assert np.array_equal(back["opacity"], opac[:, 0].numpy()) and np.array_equal(back["scale_1"], scale[:, 1].numpy())
    assert not back["nx"].any() and not back["f_rest_44"].any()


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirrors of the ABI structs (_lib.py) have the size and field offsets the C compiler gives the
    declarations in include/gof_b200.h (the header must also compile as plain C)."""
    import shutil
    import subprocess
    from f3d_gaus_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    structs = {"GofParams": _lib.GofParams, "GofInputs": _lib.GofInputs, "GofGrads": _lib.GofGrads,
               "GofHeadParams": _lib.GofHeadParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gof_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'printf("{name} %zu", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'printf(" %zu", offsetof({name}, {field}));')
        lines.append('printf("\\n");')
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for line in filter(None, out):
        name, size, *offsets = line.split()
        cls = structs[name]
        assert ctypes.sizeof(cls) == int(size), name
        assert [getattr(cls, f).offset for f, _ in cls._fields_] == [int(o) for o in offsets], name


def test_bench_workloads_and_formulas():
    """bench.py imports without a GPU and names the four BASELINE configs; the algorithmic-byte formulas are SURVEY.md 8(d)'s."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert set(bench.WORKLOAD_CLASSES) == set(bench.WORKLOADS) == {"nvs256", "train256", "cycle3", "batch512"}
    assert all(bench.MIN_TIMED_STEPS[w] >= 10 for w in bench.WORKLOADS)
    P, R, W, H = 65536, 186000, 256, 256
    assert bench.bytes_render_fwd(P, R, W, H) == 8 * 256 + 60 * R + 12 * P + 60 * W * H          # ~15.9 MB @ configs[1]
    assert bench.bytes_render_bwd(P, R, W, H) == 8 * 256 + 80 * R + 60 * W * H + 68 * P
    assert abs(bench.bytes_render_fwd(P, R, W, H) / 1e6 - 15.9) < 0.1
    assert bench.Batch512.frames_per_step == 64 and bench.Cycle3.frames_per_step == 10
    peak, src = bench.measured_peak_gbs()
    assert 5000 < peak < 9000 and ("measured" in src or "fallback" in src)


def test_save_contrib_flag_follows_requires_grad(monkeypatch):
    """The Python layer asks the forward for contributor masks exactly when a differentiable input is passed."""
    from f3d_gaus_b200 import _lib
    monkeypatch.delenv("GOF_SAVE_CONTRIB", raising=False)
    monkeypatch.delenv("GOF_EXACT_BLEND", raising=False)
    a, b = torch.zeros(3), torch.zeros(3, requires_grad=True)
    assert _lib.default_flags(a, None) == 0
    assert _lib.default_flags(a, b) == _lib.FLAG_SAVE_CONTRIB
    monkeypatch.setenv("GOF_SAVE_CONTRIB", "0")
    assert _lib.default_flags(a, b) == 0
    monkeypatch.setenv("GOF_SAVE_CONTRIB", "1")
    monkeypatch.setenv("GOF_EXACT_BLEND", "1")
    assert _lib.default_flags(a) == (_lib.FLAG_SAVE_CONTRIB | _lib.FLAG_EXACT_BLEND)
