"""Randomised GPU parity against the reference build over a wide parameter space: image sizes that are not
multiples of the tile size, wide and narrow fields of view, Gaussians from far smaller than a pixel to larger
than the image, anisotropy up to 1:100, un-normalised quaternions, points behind / across the near plane,
opacities below the 1/255 cut and at 1, SH degrees 0-3 and precomputed colours, kernel_size and scale_modifier.
This is the test of the conic pre-test's soundness proof (csrc/conic.cuh): any wrongly skipped pair shows up as
a contributor-count or image mismatch."""
import math
import os

import pytest
import torch

import refgpu

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refgpu.ref_available(), reason="oracle/_ref/libgof_ref.so not built")


def random_case(seed: int, device="cuda") -> dict:
    g = torch.Generator().manual_seed(1000 + seed)
    r = lambda *s: torch.rand(*s, generator=g)
    rn = lambda *s: torch.randn(*s, generator=g)
    P = int(torch.randint(1, 6000, (1,), generator=g))
    W = int(torch.randint(17, 400, (1,), generator=g))
    H = int(torch.randint(17, 400, (1,), generator=g))
    fov = float(10 + 80 * r(1))
    tanfov = math.tan(math.radians(fov) / 2)
    depth_lo, depth_hi = (0.05, 6.0) if seed % 3 == 0 else (1.0, 12.0)     # every third case straddles the near plane
    z = depth_lo + (depth_hi - depth_lo) * r(P)
    if seed % 4 == 1:
        z[: P // 10] *= -1.0                                               # some behind the camera
    spread = 1.4 * tanfov
    xyz = torch.stack([(2 * r(P) - 1) * spread * z.abs(), (2 * r(P) - 1) * spread * z.abs(), z], dim=-1)
    base = torch.exp(math.log(1e-3) + (math.log(1.0) - math.log(1e-3)) * r(P, 1))      # 1e-3 .. 1, log-uniform
    aniso = torch.exp(math.log(100.0) * r(P, 3) * (1.0 if seed % 2 else 0.3))          # up to 1:100
    scales = (base * aniso / aniso.max(dim=1, keepdim=True).values).clamp_min(1e-4)
    rot = rn(P, 4)
    if seed % 5 != 0:
        rot = torch.nn.functional.normalize(rot, dim=-1)                   # every fifth case: un-normalised quaternions
    else:
        rot = rot * (0.5 + r(P, 1))
    op = torch.sigmoid(3.0 * rn(P, 1))
    op[: P // 20] = 0.003                                                  # below the 1/255 cut
    op[P // 20: P // 10] = 1.0
    D = seed % 4
    wv = torch.eye(4)
    a = float(0.3 * (r(1) - 0.5))
    wv[0, 0], wv[0, 2], wv[2, 0], wv[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    wv[3, :3] = 0.2 * (r(3) - 0.5)
    n, f = 0.1, 100.0
    proj = torch.zeros(4, 4)
    proj[0, 0] = 1 / tanfov; proj[1, 1] = 1 / tanfov; proj[2, 2] = (n + f) / (f - n); proj[2, 3] = 1.0; proj[3, 2] = -(f * n) / (f - n)
    d = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
    c = {"W": W, "H": H, "D": D, "tanfovx": tanfov, "tanfovy": tanfov * (0.8 + 0.4 * float(r(1))),
         "kernel_size": 0.1 if seed % 3 == 1 else 0.0, "scale_modifier": 1.0 if seed % 4 else 1.5,
         "bg": d(r(3)), "means3D": d(xyz), "opacities": d(op), "scales": d(scales), "rotations": d(rot),
         "viewmatrix": d(wv), "projmatrix": d(wv @ proj), "campos": d(wv.inverse()[3, :3])}
    if seed % 6 == 5:
        c["colors_precomp"] = d(r(P, 3))
        c["D"] = 0
    else:
        M = (D + 1) ** 2
        sh = 0.3 * rn(P, M, 3)
        sh[:, 0] += 1.0
        c["shs"] = d(sh)
    return c


@needs_ref
@pytest.mark.parametrize("exact", [False, True], ids=["fast_blend", "exact_blend"])
@pytest.mark.parametrize("seed", list(range(int(os.environ.get("GOF_FUZZ_SEEDS", "24")))))      # GOF_FUZZ_SEEDS=300 for a long soak
def test_random_case_matches_reference(seed, exact, monkeypatch):
    monkeypatch.setenv("GOF_EXACT_BLEND", "1" if exact else "0")
    c = random_case(seed)
    ref = refgpu.RefRun().forward(c)
    ours = refgpu.OursRun().forward(c)
    bits = lambda t: t.contiguous().view(torch.int32) if t.dtype == torch.float32 else t
    vis = ref["radii"] > 0
    assert ours["num_rendered"] == ref["num_rendered"]
    assert torch.equal(ours["radii"], ref["radii"])
    for k in ("depths", "means2D", "conic_opacity", "view2gaussian"):
        assert torch.equal(bits(ours[k][vis]), bits(ref[k][vis])), k
    assert torch.equal(ours["point_list_keys"], ref["point_list_keys"])
    assert torch.equal(ours["point_list"], ref["point_list"])
    assert torch.equal(ours["ranges"], ref["ranges"])
    assert torch.equal(ours["n_contrib"], ref["n_contrib"]), "contributor counts differ: a pair was skipped or added"
    for ch in (0, 1, 2, 6, 7):
        assert torch.equal(bits(ours["out_color"][ch]), bits(ref["out_color"][ch])), f"channel {ch}"
    assert torch.equal(bits(ours["final_T"][0]), bits(ref["final_T"][0]))
    if exact:
        assert torch.equal(bits(ours["out_color"]), bits(ref["out_color"]))
    else:
        fin = torch.isfinite(ref["out_color"]) & torch.isfinite(ours["out_color"])
        assert (ours["out_color"][fin] - ref["out_color"][fin]).abs().max().item() <= 1e-4


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2, 5, 7, 11])
def test_random_case_backward_blend(seed):
    c = random_case(seed)
    g = torch.Generator().manual_seed(seed)
    dL = torch.randn(9, c["H"], c["W"], generator=g).cuda()
    r = refgpu.RefRun(); r.forward(c, decode_state=False); ref = r.backward(c, dL)
    o = refgpu.OursRun(); o.forward(c, decode_state=False); ours = o.backward(c, dL)
    for k in ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dsh", "dL_dview2gaussian"):
        a, b = ours[k].double(), ref[k].double()
        if b.numel() == 0:
            continue
        fin = torch.isfinite(a) & torch.isfinite(b)
        rel = (a[fin] - b[fin]).norm().item() / max(b[fin].norm().item(), 1e-30)
        assert rel <= 1e-3, (k, rel)
