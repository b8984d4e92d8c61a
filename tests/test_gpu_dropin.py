""""Drops into F3D-Gaus unchanged": the reference's OWN Python, unmodified, running on top of the product.

oracle/_ref/pyref holds byte-for-byte copies of RAST/diff_gof_rasterization/__init__.py and
src/gaussian_renderer/__init__.py staged by oracle/Makefile (tests/pyref.py).  Three stacks are exercised:

  A  reference renderer  ->  product's diff_gof_rasterization package (install_drop_in)  ->  libgof_b200
  B  reference rasterizer package (its _RasterizeGaussians / GaussianRasterizer_GOF)     ->  product's _C -> libgof_b200
  C  reference renderer  ->  reference rasterizer package                                ->  product's _C -> libgof_b200

and compared with the product's own wrappers (gaussian_renderer.render_predicted_more_v2_gof, its autograd.Function).
Raster-derived outputs must be bit-identical (same kernels underneath); the normals of the fused epilogue within 1e-5 of
the reference's torch post-processing where that is well conditioned (tests/test_gpu_epilogue.py has the float64 analysis).
"""
import pytest
import torch

import pyref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/pyref not staged (make -C oracle pyref)")]

RASTER_KEYS = ("render", "rendered_depth", "rendered_alpha", "distortion_map", "radii", "visibility_filter")


def _scene(S, res, seed=0, grad=False):
    from f3d_gaus_b200 import cameras, synthetic
    pc = {k: v.to("cuda") for k, v in synthetic.f3d_like(seed, S).items()}
    if grad:
        pc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    cams = cameras.orbit_cameras(8)
    cfg = synthetic.cfg_for(res)
    return pc, cams, cfg


def _cam(cams, v):
    return cams.world_view[v:v + 1].cuda(), cams.full_proj[v:v + 1].cuda(), cams.centers[v:v + 1].cuda()


@pytest.mark.parametrize("stack", ["dropin", "refpy"])
@pytest.mark.parametrize("S,res,view", [(256, 256, 2), (96, 128, 5)])
def test_reference_renderer_no_grad(stack, S, res, view):
    """Stacks A and C under no_grad (how visualize.py:288-340 calls it) against the product's wrapper."""
    from f3d_gaus_b200.gaussian_renderer import render_predicted_more_v2_gof as ours_fn
    ref_mod = pyref.reference_renderer(stack)
    pc, cams, cfg = _scene(S, res)
    wv, fp, cc = _cam(cams, view)
    bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
    with torch.no_grad():
        theirs = ref_mod.render_predicted_more_v2_gof(pc, 0, wv, fp, cc, bg, cfg)
        ours = ours_fn(pc, 0, wv, fp, cc, bg, cfg)
    assert set(theirs) == set(ours)
    for k in RASTER_KEYS:
        assert theirs[k].shape == ours[k].shape and theirs[k].dtype == ours[k].dtype, k
        assert torch.equal(theirs[k], ours[k]), k
    assert theirs["viewspace_points"].shape == ours["viewspace_points"].shape
    assert (theirs["rendered_normal"] - ours["rendered_normal"]).abs().max().item() <= 1e-5
    # depth normal: the reference's float32 evaluation loses ~1e-4 at oblique cameras (test_gpu_epilogue.py); the
    # bulk of the pixels agrees far better than that
    d = (theirs["depth_normal"] - ours["depth_normal"]).abs()
    assert d.max().item() <= 5e-4 and d.mean().item() <= 1e-5
    border = torch.ones((res, res), dtype=torch.bool, device="cuda")
    border[1:-1, 1:-1] = False
    assert float(ours["depth_normal"][:, border].abs().max()) == 0.0 and float(theirs["depth_normal"][:, border].abs().max()) == 0.0


def test_reference_autograd_function_over_product_C():
    """Stack B: forward + backward through the reference's own _RasterizeGaussians over the product's `_C`."""
    import math
    from f3d_gaus_b200.diff_gof_rasterization import GaussianRasterizationSettings_GOF as OurSettings, GaussianRasterizer_GOF as OurRast
    ref_pkg = pyref.reference_rasterizer_package()
    assert ref_pkg._C is not None and hasattr(ref_pkg, "_RasterizeGaussians")
    pc, cams, cfg = _scene(96, 128, seed=1)
    wv, fp, cc = _cam(cams, 2)
    tanfov = math.tan(13.164 * math.pi / 360)
    bg = torch.tensor([0.3, 0.1, 0.2], device="cuda")
    dL = torch.randn(9, 128, 128, generator=torch.Generator().manual_seed(5)).cuda()
    shs0 = torch.cat([pc["features_dc"][0], pc["features_rest"][0]], dim=1).contiguous()

    def run(Settings, Rast):
        leaves = {k: pc[k][0].clone().requires_grad_(True) for k in ("xyz", "opacity", "scaling", "rotation")}
        shs = shs0.clone().requires_grad_(True)
        m2d = torch.zeros_like(leaves["xyz"], requires_grad=True)
        rs = Settings(image_height=128, image_width=128, tanfovx=tanfov, tanfovy=tanfov, kernel_size=0.0,
                      subpixel_offset=torch.zeros((128, 128, 2), device="cuda"), bg=bg, scale_modifier=1.0, viewmatrix=wv,
                      projmatrix=fp, sh_degree=1, campos=cc, prefiltered=False, debug=False)
        color, radii = Rast(raster_settings=rs)(means3D=leaves["xyz"], means2D=m2d, shs=shs, colors_precomp=None,
                                                opacities=leaves["opacity"], scales=leaves["scaling"],
                                                rotations=leaves["rotation"], cov3D_precomp=None, view2gaussian_precomp=None)
        (color * dL).sum().backward()
        g = {k: v.grad for k, v in leaves.items()}
        g["shs"], g["means2D"] = shs.grad, m2d.grad
        return color.detach(), radii, g

    c_ref, r_ref, g_ref = run(ref_pkg.GaussianRasterizationSettings_GOF, ref_pkg.GaussianRasterizer_GOF)
    c_our, r_our, g_our = run(OurSettings, OurRast)
    assert torch.equal(c_ref, c_our) and torch.equal(r_ref, r_our)
    rel = lambda a, b: (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)
    for k in ("opacity", "shs", "means2D"):
        assert g_ref[k] is not None and rel(g_ref[k], g_our[k]) <= 1e-5, (k, rel(g_ref[k], g_our[k]))
    for k in ("xyz", "scaling", "rotation"):              # same kernels; only the order of the float atomics differs
        assert g_ref[k] is not None and bool(torch.isfinite(g_ref[k]).all())
        assert g_ref[k].shape == g_our[k].shape


def test_reference_full_stack_training_step():
    """Stack C with autograd: the reference's renderer + its torch post-processing + its autograd bridge, backward
    through the product's kernels; against the product's wrapper (fused epilogue with the hand-written backward)."""
    from f3d_gaus_b200.gaussian_renderer import render_predicted_more_v2_gof as ours_fn
    ref_mod = pyref.reference_renderer("refpy")
    res = 128
    g = torch.Generator().manual_seed(9)
    w_rgb = torch.randn(3, res, res, generator=g).cuda()
    w_n = torch.randn(3, res, res, generator=g).cuda()

    def step(fn):
        pc, cams, cfg = _scene(64, res, seed=2, grad=True)
        wv, fp, cc = _cam(cams, 2)
        o = fn(pc, 0, wv, fp, cc, torch.zeros(3, device="cuda"), cfg)
        loss = (o["render"] * w_rgb).sum() + (o["rendered_normal"] * w_n).sum() + o["rendered_depth"].sum() \
            + 0.1 * (o["distortion_map"]).sum()
        loss.backward()
        return o, {k: v.grad for k, v in pc.items()}

    o_ref, g_ref = step(ref_mod.render_predicted_more_v2_gof)
    o_our, g_our = step(ours_fn)
    for k in RASTER_KEYS:
        assert torch.equal(o_ref[k], o_our[k]), k
    rel = lambda a, b: (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)
    # well-conditioned gradients (through colour / opacity): the two stacks agree to float32 rounding
    for k in ("opacity", "features_dc", "features_rest"):
        assert g_ref[k] is not None and g_our[k] is not None
        assert rel(g_ref[k], g_our[k]) <= 1e-4, (k, rel(g_ref[k], g_our[k]))
    for k in ("xyz", "scaling", "rotation"):
        assert g_ref[k] is not None and g_our[k] is not None and bool(torch.isfinite(g_our[k]).all())


def test_mark_visible_through_reference_module():
    """GaussianRasterizer_GOF.markVisible of the reference package over the product's _C.mark_visible."""
    import math
    ref_pkg = pyref.reference_rasterizer_package()
    pc, cams, cfg = _scene(64, 128)
    wv, fp, cc = _cam(cams, 3)
    tanfov = math.tan(13.164 * math.pi / 360)
    rs = ref_pkg.GaussianRasterizationSettings_GOF(128, 128, tanfov, tanfov, 0.0, torch.zeros(1, device="cuda"),
                                                   torch.zeros(3, device="cuda"), 1.0, wv, fp, 1, cc, False, False)
    xyz = pc["xyz"][0].clone()
    xyz[::7, 2] -= 9.0                      # push some points behind the near plane
    vis = ref_pkg.GaussianRasterizer_GOF(rs).markVisible(xyz)
    assert vis.dtype == torch.bool and vis.shape == (xyz.shape[0],)
    assert 0 < int(vis.sum()) < xyz.shape[0]
