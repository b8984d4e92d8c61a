"""GPU tests of the batched entry point (gof_forward_batch / rasterize_views / render_views): frame v of a
batch must be bit-identical to the single-frame call with camera v, in both hand-off modes."""
import pytest
import torch

import cases
import refgpu

pytestmark = pytest.mark.gpu


def _scene(S, res, device="cuda", seed=0):
    from f3d_gaus_b200 import cameras, synthetic
    pc = {k: v.to(device) for k, v in synthetic.f3d_like(seed, S).items()}
    cams = cameras.orbit_cameras(8)
    cfg = synthetic.cfg_for(res)
    return pc, cams, cfg


def _per_view(pc, cams, res, v, device="cuda"):
    c = cases.make_case({k: t.cpu() for k, t in pc.items()}, cams.world_view[v], cams.full_proj[v], cams.centers[v],
                        W=res, H=res, fov_deg=13.164, device=device)
    o = refgpu.OursRun().forward(c)
    return c, o


def _batch(pc, cams, res, workspace=None, views=range(8)):
    from f3d_gaus_b200.diff_gof_rasterization import rasterize_views
    import math
    idx = list(views)
    dev = pc["xyz"].device
    shs = torch.cat([pc["features_dc"][0], pc["features_rest"][0]], dim=1).contiguous()
    tanfov = math.tan(13.164 * math.pi / 360)
    return rasterize_views(torch.zeros(3, device=dev), pc["xyz"][0], None, pc["opacity"][0], pc["scaling"][0],
                           pc["rotation"][0], 1.0, cams.world_view[idx].to(dev), cams.full_proj[idx].to(dev), tanfov,
                           tanfov, 0.0, res, res, shs, 1, cams.centers[idx].to(dev), workspace=workspace)


@pytest.mark.parametrize("S,res", [(64, 128), (256, 256), (96, 200)])
def test_batch_equals_per_view(S, res):
    from f3d_gaus_b200.diff_gof_rasterization import state_array_batch
    pc, cams, cfg = _scene(S, res)
    R, color, radii, geom, binning, img = _batch(pc, cams, res)
    torch.cuda.synchronize()
    P, V = S * S, 8
    nc = state_array_batch("n_contrib", P, res, res, V, sum(R), geom, binning, img)
    fT = state_array_batch("final_T", P, res, res, V, sum(R), geom, binning, img)
    rng = state_array_batch("ranges", P, res, res, V, sum(R), geom, binning, img)
    plist = state_array_batch("point_list", P, res, res, V, sum(R), geom, binning, img)
    keys = state_array_batch("point_list_keys", P, res, res, V, sum(R), geom, binning, img)
    T = ((res + 15) // 16) ** 2
    start = 0
    for v in range(V):
        c, o = _per_view(pc, cams, res, v)
        assert R[v] == o["num_rendered"]
        assert torch.equal(radii[v], o["radii"])
        assert torch.equal(color[v].view(torch.int32), o["out_color"].view(torch.int32)), f"view {v}"
        assert torch.equal(nc[v].reshape(2, -1), o["n_contrib"])
        assert torch.equal(fT[v].reshape(4, -1).view(torch.int32), o["final_T"].view(torch.int32))
        # the batch's tile lists are the per-view lists, concatenated in view order
        assert torch.equal(plist[start:start + R[v]], o["point_list"])
        assert torch.equal(keys[start:start + R[v]], o["point_list_keys"])
        rv = rng[v * T:(v + 1) * T]
        touched = rv[:, 1] > rv[:, 0]
        assert torch.equal((rv - start)[touched], o["ranges"][touched])
        assert not rv[~touched].any()
        start += R[v]


def test_workspace_sync_free_and_overflow_retry():
    from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace
    pc, cams, cfg = _scene(128, 256)
    R0, color0, radii0, *_ = _batch(pc, cams, 256)
    ws = BatchWorkspace("cuda:0")
    R, color, radii, *_ = _batch(pc, cams, 256, workspace=ws)
    assert R is None
    torch.cuda.synchronize()
    assert ws.finish() == R0
    assert torch.equal(color.view(torch.int32), color0.view(torch.int32)) and torch.equal(radii, radii0)
    # a blob that is too small: overflow is reported, buffers grow, the re-run is correct
    ws2 = BatchWorkspace("cuda:0")
    ws2.capacity_hint = 1000
    _batch(pc, cams, 256, workspace=ws2)
    assert ws2.finish() is None
    R, color, radii, *_ = _batch(pc, cams, 256, workspace=ws2)
    assert ws2.finish() == R0
    assert torch.equal(color.view(torch.int32), color0.view(torch.int32))


def test_render_views_matches_render_predicted():
    from f3d_gaus_b200.gaussian_renderer import render_predicted_more_v2_gof, render_views
    pc, cams, cfg = _scene(128, 256)
    dev = "cuda"
    wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    out = render_views(pc, 0, wv, fp, cc, bg, cfg)
    with torch.no_grad():
        for v in (0, 3, 7):
            o = render_predicted_more_v2_gof(pc, 0, wv[v:v + 1], fp[v:v + 1], cc[v:v + 1], bg, cfg)
            for k in ("render", "rendered_depth", "rendered_alpha", "distortion_map", "rendered_normal", "depth_normal"):
                assert torch.equal(out[k][v], o[k]), (k, v)
            assert torch.equal(out["radii"][v], o["radii"])


def test_big_tile_global_sort_path():
    """> 4096 Gaussians in one tile: the bucket is sorted in place in global memory."""
    c = cases.unit_case(3, 6000, 32, 32, device="cuda")
    c["means3D"] = c["means3D"].clone()
    c["means3D"][:, :2] *= 0.05          # everything lands in the same few tiles
    o = refgpu.OursRun().forward(c)
    rng = o["ranges"].long()
    assert int((rng[:, 1] - rng[:, 0]).max()) > 4096
    keys = o["point_list_keys"]
    assert bool((keys[1:] >= keys[:-1]).all())
    if refgpu.ref_available():
        r = refgpu.RefRun().forward(c)
        assert torch.equal(o["point_list"], r["point_list"]) and torch.equal(keys, r["point_list_keys"])
        assert torch.equal(o["out_color"].view(torch.int32), r["out_color"].view(torch.int32)) or \
            (o["out_color"] - r["out_color"]).abs().max().item() <= 2e-5


def test_config4_cycle_aggregative_loop_parity():
    """BASELINE configs[3]: cycle-aggregative 3-view loop at 256x256 -> 196,608 aggregated Gaussians.  The loop
    (cycle.cycle_aggregate, batched renders, stand-in predictor) builds the merged set; the merged set is
    then rendered by the reference build and by ours: integer state and image must agree."""
    from f3d_gaus_b200 import cameras, cycle, synthetic
    from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace
    dev = "cuda"
    res = 256
    cfg = synthetic.cfg_for(res)
    pc = {k: v.to(dev) for k, v in synthetic.f3d_like(0, 256).items()}
    orbit = cameras.orbit_cameras(8)
    pick = [2, 5]                                   # two aggregation views -> 3 sets
    cams = cameras.Cameras(*[t[pick].to(dev) for t in orbit])
    ws = BatchWorkspace("cuda:0")
    merged, frames = cycle.cycle_aggregate(pc, cycle.unproject_predictor(cfg), cams, cfg, torch.zeros(3, device=dev),
                                           workspace=ws)
    assert merged["xyz"].shape == (1, 3 * 65536, 3)
    assert frames["rgb"].shape == (1, 2, 3, res, res) and float(frames["alpha"].max()) > 0.5
    # render the aggregated set from a third view, reference vs ours
    v = 3
    c = cases.make_case({k: t.cpu() for k, t in merged.items()}, orbit.world_view[v], orbit.full_proj[v], orbit.centers[v],
                        W=res, H=res, fov_deg=13.164, device=dev)
    ours = refgpu.OursRun().forward(c)
    assert ours["num_rendered"] > 400000
    if refgpu.ref_available():
        ref = refgpu.RefRun().forward(c)
        assert ours["num_rendered"] == ref["num_rendered"]
        assert torch.equal(ours["point_list"], ref["point_list"])
        assert torch.equal(ours["n_contrib"], ref["n_contrib"])
        d = (ours["out_color"] - ref["out_color"]).abs().max().item()
        assert d <= 1e-4, d
        for ch in (0, 1, 2, 6, 7):
            assert torch.equal(ours["out_color"][ch].view(torch.int32), ref["out_color"][ch].view(torch.int32))


@pytest.mark.parametrize("kind", ["unit", "f3d"])
def test_batched_backward_equals_sum_of_per_view_backwards(kind):
    """gof_backward_batch: gradients of V views in one pass == the sum of the per-view backwards (what autograd
    accumulates when the reference renders the views one call at a time)."""
    from f3d_gaus_b200 import cameras, synthetic
    from f3d_gaus_b200.diff_gof_rasterization import (GaussianRasterizationSettings_GOF, GaussianRasterizer_GOF,
                                                     backward_accumulators, rasterize_views_autograd)
    import math
    dev = "cuda"
    if kind == "unit":
        pc = synthetic.unit_cloud(0, 3000)
        wv0, proj, _ = synthetic.perspective_camera(60.0)
        yaw = torch.tensor([0.0, 0.1, -0.15, 0.05])
        wvs, fps, ccs = [], [], []
        for a in yaw.tolist():
            Rm = torch.eye(4)
            Rm[0, 0], Rm[0, 2], Rm[2, 0], Rm[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
            Rm[3, 0] = 0.2 * a
            wvs.append(Rm); fps.append(Rm @ proj); ccs.append(Rm.inverse()[3, :3])
        wv, fp, cc = torch.stack(wvs).to(dev), torch.stack(fps).to(dev), torch.stack(ccs).to(dev)
        res, tanfov, D = 128, math.tan(math.radians(30.0)), 1
    else:
        pc = synthetic.f3d_like(1, 64)
        cams = cameras.orbit_cameras(8)
        wv, fp, cc = cams.world_view[:4].to(dev), cams.full_proj[:4].to(dev), cams.centers[:4].to(dev)
        res, tanfov, D = 128, math.tan(13.164 * math.pi / 360), 1
    V = wv.shape[0]
    leaves = {k: pc[k][0].to(dev).clone().requires_grad_(True) for k in ("xyz", "opacity", "scaling", "rotation")}
    shs = torch.cat([pc["features_dc"][0], pc["features_rest"][0]], dim=1).to(dev).clone().requires_grad_(True)
    bg = torch.tensor([0.1, 0.3, 0.2], device=dev)
    g = torch.Generator().manual_seed(3)
    dL = torch.randn(V, 9, res, res, generator=g).to(dev)

    import numpy as np
    import oracle_cpu
    n = lambda t: t.detach().cpu().numpy()
    rel = lambda a, b: (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)
    P = leaves["xyz"].shape[0]

    # ---- batched: one forward + one backward for the V views, through autograd ----
    m2d = torch.zeros_like(leaves["xyz"], requires_grad=True)
    c2, _ = rasterize_views_autograd(leaves["xyz"], m2d, leaves["opacity"], shs=shs, scales=leaves["scaling"],
                                     rotations=leaves["rotation"], bg=bg, viewmatrices=wv, projmatrices=fp, campos=cc,
                                     tanfovx=tanfov, tanfovy=tanfov, image_height=res, image_width=res, sh_degree=D)
    (c2 * dL).sum().backward()
    g2 = {"xyz": leaves["xyz"].grad, "scaling": leaves["scaling"].grad, "rotation": leaves["rotation"].grad,
          "opacity": leaves["opacity"].grad, "shs": shs.grad, "means2D": m2d.grad}
    acc = backward_accumulators(leaves["xyz"].device, V, P)     # [V,P,20]: what the batched K10 consumed, per view

    # ---- per view: V reference-shaped forward + backward calls, gradients summed as autograd would ----
    names = {"xyz": "dL_dmeans3D", "scaling": "dL_dscales", "rotation": "dL_drotations", "opacity": "dL_dopacity",
             "shs": "dL_dsh", "means2D": "dL_dmeans2D"}
    quad = ("xyz", "scaling", "rotation")
    g1 = {k: 0.0 for k in names}
    E_pv = {k: 0.0 for k in quad}
    E_b = {k: 0.0 for k in quad}
    for v in range(V):
        c = {"W": res, "H": res, "D": D, "tanfovx": tanfov, "tanfovy": tanfov, "kernel_size": 0.0, "scale_modifier": 1.0,
             "bg": bg, "means3D": leaves["xyz"].detach(), "opacities": leaves["opacity"].detach(),
             "scales": leaves["scaling"].detach(), "rotations": leaves["rotation"].detach(), "shs": shs.detach(),
             "viewmatrix": wv[v].contiguous(), "projmatrix": fp[v].contiguous(), "campos": cc[v].contiguous()}
        o = refgpu.OursRun()
        f = o.forward(c, decode_state=True)
        assert torch.equal(f["out_color"].view(torch.int32), c2[v].detach().view(torch.int32)), f"view {v}"
        gv = o.backward(c, dL[v])
        for k, kk in names.items():
            g1[k] = g1[k] + gv[kk]
        # both sets of per-view blend gradients through the float64 oracle of the per-Gaussian map
        cn = oracle_cpu.case_to_numpy(c)
        for dst, dq, dcol in ((E_pv, gv["dL_dview2gaussian"], gv["dL_dcolors"]), (E_b, acc[v, :, 0:10], acc[v, :, 10:13])):
            ex = oracle_cpu.preprocess_backward(cn, n(f["radii"]), n(f["clamped"]), n(dq.contiguous()), n(dcol.contiguous()), f64=True)
            for k in quad:
                dst[k] = dst[k] + ex[names[k]]

    for k in ("opacity", "shs", "means2D"):
        assert rel(g2[k].reshape(g1[k].shape), g1[k]) <= 1e-5, (k, rel(g2[k].reshape(g1[k].shape), g1[k]))
    # Quadric gradients (xyz, scaling, rotation).  K10 is the exact (double) map of the float32 blend gradients it is
    # handed, per view; the two paths hand it blend gradients that differ by the order of their float atomics (~1e-7),
    # and the map amplifies that by ~(t/s)^2 (~4e3 on the unit cloud, ~6e5 at F3D-Gaus scales).  So the bound is
    # DERIVED, not chosen: each path must reproduce the float64 map of ITS OWN per-view blend gradients within the bar,
    # and the two may differ by what the exact propagation of the two inputs gives.
    for k in quad:
        exact_b, exact_pv = torch.from_numpy(np.asarray(E_b[k])), torch.from_numpy(np.asarray(E_pv[k]))
        assert rel(g2[k].cpu(), exact_b) <= 1e-4, (k, "batched", rel(g2[k].cpu(), exact_b))
        assert rel(g1[k].cpu(), exact_pv) <= 1e-4, (k, "per view", rel(g1[k].cpu(), exact_pv))
        carried = rel(exact_b, exact_pv)
        assert rel(g2[k], g1[k]) <= 1e-3 + 1.1 * carried, (k, rel(g2[k], g1[k]), carried)


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("chunks,zero_copy,channels_last", [(3, False, False), (1, True, False), (1, True, True), (3, True, True)])
def test_host_frame_sink(chunks, zero_copy, channels_last, exact, monkeypatch):
    """Frames read back to pinned host memory: DMA copies pipelined behind the passes, or stored by the blend kernel
    itself (gof_set_frame_sink); both bit-identical to the device result."""
    from f3d_gaus_b200.gaussian_renderer import HostFrameSink, render_views
    monkeypatch.setenv("GOF_EXACT_BLEND", "1" if exact else "0")          # both blend kernels have a sink instantiation
    pc, cams, cfg = _scene(128, 256)
    dev = "cuda"
    wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    want = render_views(pc, 0, wv, fp, cc, bg, cfg, epilogue=False)
    sink = HostFrameSink(8, 256, 256, dev, chunks=chunks, zero_copy=zero_copy, channels_last=channels_last)
    sink.host.fill_(-7.0)
    for _ in range(2):                      # first call sizes the binning blobs (may overflow and grow)
        host = sink.render(pc, 0, wv, fp, cc, bg, cfg)
        R = sink.finish()
        if R is not None:
            break
    assert R == want["num_rendered"]
    assert torch.equal(host[:, 0:3], want["render"].cpu())
    assert torch.equal(host[:, 3:4], want["rendered_depth"].cpu())
    assert torch.equal(host[:, 4:5], want["rendered_alpha"].cpu())


def test_frame_sink_abi_errors():
    """gof_set_frame_sink: pageable host memory is refused; a sink smaller than [V,5,H,W] fails the forward call, and
    the sink is one-shot (the next call does not write it)."""
    import ctypes
    from f3d_gaus_b200 import _lib
    from f3d_gaus_b200.gaussian_renderer import render_views
    pc, cams, cfg = _scene(32, 64)
    dev = torch.device("cuda", torch.cuda.current_device())
    wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
    bg = torch.zeros(3, device=dev)
    ctx = _lib.context(dev.index)
    pageable = torch.empty(8 * 5 * 64 * 64)
    assert _lib.lib.gof_set_frame_sink(ctx, pageable.data_ptr(), pageable.numel() * 4, 0) == _lib.GOF_EINVAL
    assert "pinned" in _lib.last_error()
    small = torch.empty(5 * 64 * 64, device=dev)
    _lib.check(_lib.lib.gof_set_frame_sink(ctx, small.data_ptr(), small.numel() * 4, _lib.SINK_HWC), "set")
    with pytest.raises(RuntimeError, match="frame sink too small"):
        render_views(pc, 0, wv, fp, cc, bg, cfg, epilogue=False)
    on_dev = torch.full((8, 5, 64, 64), -7.0, device=dev)
    want = render_views(pc, 0, wv, fp, cc, bg, cfg, epilogue=False, sink=on_dev)
    assert torch.equal(on_dev[:, 0:3], want["render"]) and torch.equal(on_dev[:, 4:5], want["rendered_alpha"])
    on_dev.fill_(-7.0)
    render_views(pc, 0, wv, fp, cc, bg, cfg, epilogue=False)          # no sink pending any more
    assert bool((on_dev == -7.0).all())


@pytest.mark.parametrize("res", [100, 72])
@pytest.mark.parametrize("channels_last", [False, True])
def test_frame_sink_ragged_tiles(res, channels_last):
    """Image sizes that are not multiples of the 16x16 tile (partial tiles take the per-pixel store path)."""
    from f3d_gaus_b200.gaussian_renderer import render_views
    pc, cams, cfg = _scene(64, res)
    dev = torch.device("cuda", torch.cuda.current_device())
    wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
    bg = torch.tensor([0.3, 0.1, 0.2], device=dev)
    sink = torch.full((8, res, res, 5) if channels_last else (8, 5, res, res), -7.0, device=dev)
    if channels_last:
        sink = sink.permute(0, 3, 1, 2)
    o = render_views(pc, 0, wv, fp, cc, bg, cfg, epilogue=False, sink=sink)
    assert torch.equal(sink[:, 0:3], o["render"])
    assert torch.equal(sink[:, 3:4], o["rendered_depth"])
    assert torch.equal(sink[:, 4:5], o["rendered_alpha"])


def test_pinned_scene_single_upload():
    """PinnedScene: one pinned slab + one device slab; the device views equal the per-key copies and feed the renderer."""
    from f3d_gaus_b200 import synthetic
    from f3d_gaus_b200.gaussian_renderer import PinnedScene
    dev = torch.device("cuda", torch.cuda.current_device())
    pc_cpu = synthetic.f3d_like(2, 64)
    scene = PinnedScene(pc_cpu, dev)
    assert scene.host_slab.is_pinned() and scene.nbytes == sum(v.numel() * 4 for v in pc_cpu.values())
    d = scene.upload()
    torch.cuda.synchronize()
    for k, v in pc_cpu.items():
        assert d[k].shape == v.shape and d[k].data_ptr() % 256 == 0 and torch.equal(d[k].cpu(), v), k
    scene.host["opacity"].mul_(0.5)                      # the pinned views are writable in place
    assert torch.equal(scene.upload()["opacity"].cpu(), pc_cpu["opacity"] * 0.5)


@pytest.mark.parametrize("zero_copy", [True, False])
def test_scene_streamer_pipeline(zero_copy):
    """SceneStreamer: several scenes in flight (H2D of k+1 | render k | frames of k-1 to host); every collected frame
    block is bit-identical to render_views of that scene, in submission order, including a scene whose first render
    overflows the (deliberately undersized) binning blob and is re-rendered inside collect()."""
    from f3d_gaus_b200 import cameras, synthetic
    from f3d_gaus_b200.gaussian_renderer import PinnedScene, SceneStreamer, render_views
    dev = torch.device("cuda", torch.cuda.current_device())
    res = 128
    cams = cameras.orbit_cameras(8)
    wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
    cfg = synthetic.cfg_for(res)
    bg = torch.tensor([0.2, 0.1, 0.3], device=dev)
    scenes = [PinnedScene(synthetic.f3d_like(seed, 96), dev) for seed in range(5)]
    want = []
    for sc in scenes:
        o = render_views({k: v.to(dev) for k, v in sc.host.items()}, 0, wv, fp, cc, bg, cfg, epilogue=False)
        want.append((torch.cat([o["render"], o["rendered_depth"], o["rendered_alpha"]], dim=1).cpu(), o["num_rendered"]))
    st = SceneStreamer(8, res, res, dev, wv, fp, cc, bg, cfg, slots=2, zero_copy=zero_copy)
    for s in st.slot:
        s["ws"].capacity_hint = 2000            # far too small: the first render of each slot overflows
    got = []
    for sc in scenes:
        if st.pending == st.slots:
            frames, R = st.collect()
            got.append((frames.clone(), R))
        st.submit(sc)
    with pytest.raises(RuntimeError):
        st.submit(scenes[0]); st.submit(scenes[0]); st.submit(scenes[0])
    while st.pending:
        frames, R = st.collect()
        got.append((frames.clone(), R))
    got = got[:len(scenes)]
    assert len(got) == len(scenes)
    for k, ((f, R), (wf, wR)) in enumerate(zip(got, want)):
        assert R == wR, k
        assert torch.equal(f, wf), k


def test_cycle_loop_deferred_overflow_check():
    """cycle_aggregate(check_overflow=False) never synchronises; an undersized workspace then shows in the caller's
    workspace.finish() (None), the merged set derived from the NaN-poisoned frames is discarded, and the re-run on the
    grown workspace equals the checked loop bit for bit."""
    from f3d_gaus_b200 import cameras, cycle, synthetic
    from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace
    dev = "cuda"
    res = 128
    cfg = synthetic.cfg_for(res)
    pc = {k: v.to(dev) for k, v in synthetic.f3d_like(1, 64).items()}
    orbit = cameras.orbit_cameras(8)
    cams = cameras.Cameras(*[t[[1, 6]].to(dev) for t in orbit])
    bg = torch.zeros(3, device=dev)
    predict = cycle.unproject_predictor(cfg)
    want, want_frames = cycle.cycle_aggregate(pc, predict, cams, cfg, bg, workspace=BatchWorkspace("cuda:0"))

    ws = BatchWorkspace("cuda:0")
    ws.capacity_hint = 256                              # far too small for two 128^2 frames of 4096 Gaussians
    got, frames = cycle.cycle_aggregate(pc, predict, cams, cfg, bg, workspace=ws, check_overflow=False)
    torch.cuda.synchronize()
    assert ws.finish() is None                          # the caller's check: this step has to be computed again
    assert torch.isnan(frames["rgb"]).any()
    got, frames = cycle.cycle_aggregate(pc, predict, cams, cfg, bg, workspace=ws, check_overflow=False)
    torch.cuda.synchronize()
    assert ws.finish() is not None
    for k in want:
        assert torch.equal(got[k], want[k]), k
    assert torch.equal(frames["depth"], want_frames["depth"])
