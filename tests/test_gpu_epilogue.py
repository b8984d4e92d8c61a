"""The fused L2 epilogue (csrc/epilogue.cu: gof_render_epilogue[_batch], gof_render_epilogue_backward_batch) against
the reference's torch post-processing (src/gaussian_renderer/__init__.py:881-909,1043-1053).

Oracles, in this order:
  * tests/epilogue_torch_ref.py evaluated in FLOAT64 on the same float32 raster -- the exact value of the reference's
    formula;
  * the same restatement in float32 on the GPU -- what the reference's op sequence produces;
  * the reference's OWN depth_to_normal, unmodified, when oracle/_ref/pyref is staged.
Bars: rendered_normal within 1e-5 of the reference's float32 result; depth_normal within 1e-5 of the float64 value
(the reference's own float32 result is up to ~1e-4 away from it at oblique cameras: it adds the camera origin to both
points before subtracting them) and never further from it than the reference is; border pixels exactly 0; the backward
of the normal channels within 1e-5 (relative L2) of torch.autograd through the float64 restatement, that of the median
depth (a discrete divergence, ~2e-5 for any float32 evaluation including the reference's graph) within 1e-4.
"""
import math

import pytest
import torch

import cases
import epilogue_torch_ref as tref
import pyref
import refgpu

pytestmark = pytest.mark.gpu

CASES = {
    "f3d_canon": lambda: cases.f3d_case(0, 128, 256, None, device="cuda"),
    "f3d_view2": lambda: cases.f3d_case(1, 128, 256, 2, device="cuda"),
    "f3d_view5_r128": lambda: cases.f3d_case(2, 96, 128, 5, device="cuda"),
    "unit_ragged": lambda: cases.unit_case(0, 4096, 200, 136, device="cuda"),
}


def _raster(c):
    o = refgpu.OursRun().forward(c, decode_state=False)
    fov_x = 2 * math.atan(c["tanfovx"])
    fov_y = 2 * math.atan(c["tanfovy"])
    return o["out_color"], fov_x, fov_y


def _fused(img, c, fx, fy):
    from f3d_gaus_b200 import gaussian_renderer as gr
    return gr.fused_epilogue(img, c["viewmatrix"], c["W"], c["H"], fx, fy)


@pytest.mark.parametrize("name", list(CASES))
def test_fused_epilogue_matches_reference_postprocessing(name):
    c = CASES[name]()
    img, fx, fy = _raster(c)
    W, H = c["W"], c["H"]
    nw, dn = _fused(img, c, fx, fy)
    nw32, dn32 = tref.postprocess(img, c["viewmatrix"], W, H, fx, fy)                       # the reference's float32 ops
    nw64, dn64 = tref.postprocess(img.double(), c["viewmatrix"].double(), W, H, fx, fy)    # exact value of its formula
    # rendered normal: against the reference's float32 result and the exact value
    assert (nw - nw32).abs().max().item() <= 1e-5
    assert (nw.double() - nw64).abs().max().item() <= 1e-5
    # depth normal: border exactly zero, interior within 1e-5 of the exact value and at least as close as the reference
    border = torch.ones((H, W), dtype=torch.bool, device=img.device)
    border[1:-1, 1:-1] = False
    assert float(dn[:, border].abs().max()) == 0.0
    e_ours = (dn.double() - dn64).abs().max().item()
    e_ref = (dn32.double() - dn64).abs().max().item()
    assert e_ours <= 1e-5, (e_ours, e_ref)
    assert e_ours <= e_ref + 1e-7, (e_ours, e_ref)
    assert (dn - dn32).abs().max().item() <= e_ref + 1e-5
    if pyref.available():
        ref_mod = pyref.reference_renderer("dropin")
        theirs = ref_mod.depth_to_normal(c["viewmatrix"], W, H, fx, fy, img[6:7]).permute(2, 0, 1)
        assert (theirs - dn32).abs().max().item() <= 2e-6            # the restatement IS the reference's function
        assert (dn - theirs).abs().max().item() <= e_ref + 1e-5


def test_epilogue_batch_equals_per_view():
    import ctypes
    from f3d_gaus_b200 import _lib
    from f3d_gaus_b200 import gaussian_renderer as gr
    cs = [cases.f3d_case(1, 96, 128, v, device="cuda") for v in (0, 2, 5)]
    imgs, fx, fy = [], None, None
    for c in cs:
        img, fx, fy = _raster(c)
        imgs.append(img)
    raster = torch.stack(imgs).contiguous()
    vm = torch.stack([c["viewmatrix"].reshape(16) for c in cs]).contiguous()
    nw, dn = gr._epilogue_forward(raster, vm, 3, 128, 128, fx, fy)
    for v, c in enumerate(cs):
        a, b = gr.fused_epilogue(imgs[v], c["viewmatrix"], 128, 128, fx, fy)
        assert torch.equal(nw[v], a) and torch.equal(dn[v], b)


@pytest.mark.parametrize("name", ["f3d_view2", "unit_ragged"])
def test_fused_epilogue_backward_matches_autograd(name):
    from f3d_gaus_b200 import gaussian_renderer as gr
    c = CASES[name]()
    img, fx, fy = _raster(c)
    W, H = c["W"], c["H"]
    g = torch.Generator().manual_seed(11)
    g_nw = torch.randn(3, H, W, generator=g).cuda()
    g_dn = torch.randn(3, H, W, generator=g).cuda()
    # ours: the hand-written backward
    x = img.clone().requires_grad_(True)
    nw, dn = gr.fused_epilogue(x, c["viewmatrix"], W, H, fx, fy)
    assert nw.requires_grad and dn.requires_grad
    ((nw * g_nw).sum() + (dn * g_dn).sum()).backward()
    # oracle: autograd through the float64 restatement of the reference's ops
    x64 = img.double().clone().requires_grad_(True)
    nw64, dn64 = tref.postprocess(x64, c["viewmatrix"].double(), W, H, fx, fy)
    ((nw64 * g_nw.double()).sum() + (dn64 * g_dn.double()).sum()).backward()
    # and through its float32 form (what the reference's training graph computes)
    x32 = img.clone().requires_grad_(True)
    nw32, dn32 = tref.postprocess(x32, c["viewmatrix"], W, H, fx, fy)
    ((nw32 * g_nw).sum() + (dn32 * g_dn).sum()).backward()
    rel = lambda a, b: (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)
    # only the normal channels and the median depth receive gradient
    for ch in (0, 1, 2, 7, 8):
        assert float(x.grad[ch].abs().max()) == 0.0 and float(x64.grad[ch].abs().max()) == 0.0
    e_n, e_d = rel(x.grad[3:6], x64.grad[3:6]), rel(x.grad[6], x64.grad[6])
    r_d = rel(x32.grad[6], x64.grad[6])
    assert e_n <= 1e-5, e_n
    # the depth gradient is a discrete divergence (differences of neighbouring pixels' terms): any float32 evaluation,
    # the reference's graph included, sits ~2e-5 from the exact value; ours must be in that class and within 1e-4
    assert e_d <= 1e-4 and e_d <= 2 * r_d + 1e-5, (e_d, r_d)


def test_renderer_training_path_uses_fused_epilogue_and_backpropagates():
    """render_predicted_more_v2_gof with autograd: normals carry a graph through the fused epilogue down to the
    Gaussian parameters (no torch post-processing ops in between)."""
    from f3d_gaus_b200 import cameras, synthetic
    from f3d_gaus_b200.gaussian_renderer import render_predicted_more_v2_gof
    dev = "cuda"
    pc = {k: v.to(dev).requires_grad_(True) for k, v in synthetic.f3d_like(0, 64).items()}
    cams = cameras.orbit_cameras(8)
    cfg = synthetic.cfg_for(128)
    o = render_predicted_more_v2_gof(pc, 0, cams.world_view[2:3].to(dev), cams.full_proj[2:3].to(dev),
                                     cams.centers[2:3].to(dev), torch.zeros(3, device=dev), cfg)
    def graph_nodes(t):
        seen, todo = set(), [t.grad_fn]
        while todo:
            f = todo.pop()
            if f is None or f in seen:
                continue
            seen.add(f)
            todo += [nf for nf, _ in f.next_functions]
        return {type(f).__name__ for f in seen}
    nodes = graph_nodes(o["rendered_normal"]) | graph_nodes(o["depth_normal"])
    assert any("FusedEpilogue" in x for x in nodes), nodes
    assert not any(x.startswith(("Linalg", "Mm", "Cross", "Div")) for x in nodes), nodes      # no torch post-processing ops
    loss = (1 - (o["rendered_normal"] * o["depth_normal"]).sum(0)).mean() + o["render"].mean()
    loss.backward()
    for k in ("xyz", "scaling", "rotation", "opacity", "features_dc"):
        assert pc[k].grad is not None and bool(torch.isfinite(pc[k].grad).all()) and float(pc[k].grad.abs().max()) > 0, k
