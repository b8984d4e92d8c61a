"""Predictor output head: golden vectors from the reference's own forward (CPU), the numpy
oracle, and the CUDA kernel (gof_predictor_head) against both."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import head_oracle  # noqa: E402

GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "head", "head_*.npz")))
RTOL, ATOL = 2e-6, 2e-6          # float32: one or two ulp of |xyz| <= 16 (bmm accumulation order, exp/expf)
EXACT = ("features_dc", "unet_depth")


def _load(path):
    z = np.load(path)
    B, V, res, with_offset, sh, iso, inv_x, inv_y, origin = [int(v) for v in z["cfg"]]
    kw = dict(B=B, V=V, res=res, fov_deg=float(z["fov"]), with_offset=bool(with_offset), sh_degree=sh,
              isotropic=bool(iso), inverted_x=bool(inv_x), inverted_y=bool(inv_y), squre_clip=float(z["squre_clip"]))
    co = z["in_x"].reshape(B * V, -1, res, res)[:, 3:4] if origin else None
    return z, kw, co


def _cfg(kw, origin=False):
    return {"model": {"training_resolution": kw["res"], "fov": kw["fov_deg"], "inverted_x": kw["inverted_x"],
                      "inverted_y": kw["inverted_y"], "max_sh_degree": kw["sh_degree"], "isotropic": kw["isotropic"],
                      "origin_distances": origin, "network_with_offset": kw["with_offset"],
                      "network_without_offset": not kw["with_offset"]}}


def _check(got: dict, want: dict, where: str):
    for k, w in want.items():
        g = got[k]
        assert tuple(g.shape) == tuple(w.shape), (where, k, g.shape, w.shape)
        if w.size == 0:
            continue
        if k in EXACT:
            assert np.array_equal(g, w), (where, k)
        else:
            assert np.allclose(g, w, rtol=RTOL, atol=ATOL), (where, k, float(np.abs(g - w).max()))


def _random_case(B, V, res, with_offset, sh, seed):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_head_golden as mk
    net, depth, _, v2w, quat = mk.make_inputs(B, V, res, with_offset, sh, False, seed)
    kw = dict(B=B, V=V, res=res, fov_deg=13.164, with_offset=with_offset, sh_degree=sh, isotropic=False,
              inverted_x=False, inverted_y=True, squre_clip=10000.0)
    return net.numpy(), depth.numpy(), v2w.numpy(), quat.numpy(), kw


def test_golden_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_head_oracle_vs_golden(path):
    """The numpy restatement against the reference's forward (run unmodified on CPU, tests/golden/make_head_golden.py)."""
    z, kw, co = _load(path)
    rx, ry = head_oracle.ray_tables(kw["res"], kw["fov_deg"], kw["inverted_x"], kw["inverted_y"])
    assert np.array_equal(z["ray_dirs"][0, 0, 0, :], rx) and np.array_equal(z["ray_dirs"][0, 1, :, 0], ry)
    assert np.all(z["ray_dirs"][0, 2] == 1.0)
    got = head_oracle.head(z["in_net"], z["in_depth"], z["in_view_to_world"], z["in_quat"], const_offset=co, **kw)
    _check(got, {k[4:]: z[k] for k in z.files if k.startswith("out_")}, "oracle")
    assert np.array_equal(got["rotation"], z["out_rotation"])       # unfused float32 products: bit-identical


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_host_ray_tables_are_the_reference_buffer(path):
    from f3d_gaus_b200.predictor_head import ray_tables
    z, kw, _ = _load(path)
    x, y = ray_tables(_cfg(kw))
    assert np.array_equal(z["ray_dirs"][0, 0, 0, :], x.numpy()) and np.array_equal(z["ray_dirs"][0, 1, :, 0], y.numpy())


@pytest.mark.parametrize("path", [p for p in GOLDEN if "iso" not in p and "origin" not in p],
                         ids=lambda p: os.path.basename(p)[:-4])
def test_torch_restatement_is_bit_identical_on_cpu(path):
    """tests/head_torch_ref.py (the timing baseline of tools/bench_head.py and the GPU-side checker) reproduces the
    reference's forward bit for bit when both run on the CPU."""
    import head_torch_ref
    z, kw, _ = _load(path)
    h = head_torch_ref.TorchHead(_cfg(kw), "cpu")
    t = torch.from_numpy
    o = h(t(z["in_net"]), t(z["in_depth"]), t(z["in_view_to_world"]), t(z["in_quat"]), kw["B"], kw["V"], kw["squre_clip"])
    for k, v in o.items():
        assert np.array_equal(v.numpy(), z["out_" + k]), k


def test_config1_plumbing_shapes_on_cpu():
    """BASELINE configs[0]: one 256x256 image -> per-pixel Gaussian parameters, no raster.  The UNet is the reference's
    (out of scope); everything after it, on CPU through the oracle: dict keys and [1, 65 536, .] shapes of the
    reference's output contract (src/gaussian_predictor.py:972-1002)."""
    net, depth, v2w, quat, kw = _random_case(1, 1, 256, True, 1, seed=0)
    out = head_oracle.head(net, depth, v2w, quat, **kw)
    want = {"xyz": (1, 65536, 3), "opacity": (1, 65536, 1), "scaling": (1, 65536, 3), "rotation": (1, 65536, 4),
            "features_dc": (1, 65536, 1, 3), "features_rest": (1, 65536, 3, 3), "unet_depth": (1, 65536, 1)}
    assert {k: v.shape for k, v in out.items()} == want
    assert all(np.isfinite(v).all() for v in out.values())
    assert (out["opacity"] > 0).all() and (out["opacity"] < 1).all() and (out["scaling"] > 0).all()
    assert np.allclose(np.linalg.norm(out["rotation"], axis=-1), 1.0, atol=1e-5)      # unit quaternions stay unit


def test_head_needs_cuda():
    from f3d_gaus_b200.predictor_head import PredictorHead
    with pytest.raises(RuntimeError, match="no CPU path"):
        PredictorHead({"model": {"training_resolution": 16, "fov": 13.164, "max_sh_degree": 1,
                                 "network_with_offset": True, "network_without_offset": False}}, device="cpu")


# ------------------------------------------------ GPU -------------------------------------------------------------
def _run_cuda(z_net, z_depth, v2w, quat, kw, co=None, origin=False, sh_transform=None):
    from f3d_gaus_b200.predictor_head import PredictorHead
    dev = torch.device("cuda", torch.cuda.current_device())
    head = PredictorHead(_cfg(kw, origin), dev)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = head(t(z_net), t(z_depth), t(v2w), t(quat), kw["B"], kw["V"], squre_clip=kw["squre_clip"], const_offset=t(co),
               sh_transform=t(sh_transform))
    for k, v in out.items():
        assert v.is_contiguous() and v.dtype == torch.float32, k
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_head_cuda_vs_golden(path):
    z, kw, co = _load(path)
    got = _run_cuda(z["in_net"], z["in_depth"], z["in_view_to_world"], z["in_quat"], kw, co, origin=co is not None)
    _check(got, {k[4:]: z[k] for k in z.files if k.startswith("out_")}, "cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("B,V,res,with_offset,sh", [(8, 1, 256, True, 1), (1, 3, 100, False, 1), (2, 2, 72, True, 0)])
def test_head_cuda_vs_oracle(B, V, res, with_offset, sh):
    """Full size (8 x 65 536 Gaussians, the shipped configuration) and ragged sizes (B*V*N not a multiple of the block)."""
    net, depth, v2w, quat, kw = _random_case(B, V, res, with_offset, sh, seed=11)
    want = head_oracle.head(net, depth, v2w, quat, **kw)
    got = _run_cuda(net, depth, v2w, quat, kw)
    _check(got, want, "cuda-vs-oracle")


@pytest.mark.gpu
def test_head_cuda_vs_torch_ops_on_gpu():
    """Against the reference's torch op sequence executed on the same GPU (cuBLAS bmm, torch's CUDA exp / sigmoid)."""
    import head_torch_ref
    net, depth, v2w, quat, kw = _random_case(4, 2, 256, True, 1, seed=21)
    dev = torch.device("cuda", torch.cuda.current_device())
    h = head_torch_ref.TorchHead(_cfg(kw), dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    # full-precision float32 matmuls, like the reference's defaults
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        want = h(t(net), t(depth), t(v2w), t(quat), kw["B"], kw["V"], kw["squre_clip"])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    got = _run_cuda(net, depth, v2w, quat, kw)
    _check(got, {k: v.cpu().numpy() for k, v in want.items()}, "cuda-vs-torch")
    for k in ("opacity", "scaling", "rotation"):       # same expf / division / unfused products: bit-identical
        assert np.array_equal(got[k], want[k].cpu().numpy()), k


@pytest.mark.gpu
def test_head_explicit_sh_transform_is_identical():
    """sh_transform passed explicitly (the module's registered matrices) == derived inside the kernel."""
    net, depth, v2w, quat, kw = _random_case(2, 2, 64, True, 1, seed=5)
    a = _run_cuda(net, depth, v2w, quat, kw)
    b = _run_cuda(net, depth, v2w, quat, kw, sh_transform=head_oracle.sh_transform(v2w))
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
def test_head_errors_and_rasterizer_hand_off():
    from f3d_gaus_b200.predictor_head import PredictorHead
    from f3d_gaus_b200 import cameras, synthetic
    from f3d_gaus_b200.gaussian_renderer import render_views
    dev = torch.device("cuda", torch.cuda.current_device())
    net, depth, v2w, quat, kw = _random_case(1, 1, 64, True, 1, seed=3)
    cfg = synthetic.cfg_for(64)
    cfg["model"].update(network_with_offset=True, network_without_offset=False)
    head = PredictorHead(cfg, dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    with pytest.raises(RuntimeError, match="network_out must be"):
        head(t(net)[:, :20], t(depth), t(v2w), t(quat), 1, 1)
    eye = torch.eye(4, device=dev).reshape(1, 1, 4, 4)
    q0 = torch.tensor([[[1.0, 0, 0, 0]]], device=dev)
    pc = head(t(net), t(depth), eye, q0, 1, 1)
    cams = cameras.orbit_cameras(8)
    o = render_views(pc, 0, cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev),
                     torch.zeros(3, device=dev), cfg, epilogue=False)
    assert o["render"].shape == (8, 3, 64, 64) and bool(torch.isfinite(o["raster"]).all())
    assert float(o["rendered_alpha"].max()) > 0.0


@pytest.mark.gpu
def test_cycle_loop_with_network_and_fused_head():
    """cycle.from_network: a stand-in UNet (one conv, reference call signature) + the fused head inside the
    cycle-aggregative loop; the merged set has (1 + K) * H * W Gaussians and renders."""
    from f3d_gaus_b200 import cameras, cycle, synthetic
    from f3d_gaus_b200.gaussian_renderer import render_views
    dev = torch.device("cuda", torch.cuda.current_device())
    res, K = 64, 2
    cfg = synthetic.cfg_for(res)
    cfg["model"].update(network_with_offset=True, network_without_offset=False, cross_view_attention=True)
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(4, 23, 3, padding=1).to(dev)
    with torch.no_grad():
        conv.weight.mul_(0.05)
        conv.bias.copy_(torch.tensor([0.0] * 3 + [2.0] + [-4.6] * 3 + [1.0, 0, 0, 0] + [0.5] * 3 + [0.0] * 9))

    def network(x, film_camera_emb=None, N_views_xa=1):
        with torch.no_grad():
            return conv(x)

    predict = cycle.from_network(network, cfg, dev)
    pc = {k: v.to(dev) for k, v in synthetic.f3d_like(0, res).items()}
    orbit = cameras.orbit_cameras(8)
    cams = cameras.Cameras(*[t[[2, 5]].to(dev) for t in orbit])
    bg = torch.zeros(3, device=dev)
    merged, frames = cycle.cycle_aggregate(pc, predict, cams, cfg, bg)
    assert merged["xyz"].shape == (1, (1 + K) * res * res, 3) and merged["features_rest"].shape[2:] == (3, 3)
    assert bool(torch.isfinite(merged["xyz"]).all())
    o = render_views(merged, 0, cams.world_view, cams.full_proj, cams.centers, bg, cfg, epilogue=False)
    assert bool(torch.isfinite(o["raster"]).all()) and float(o["rendered_alpha"].max()) > 0.5
