"""GPU tests of the one-frame-per-call path (the reference-shaped `_C.rasterize_gaussians`):

* the latency variant of the forward blend (render_fwd_split_kernel: half-tile CTAs, sweeper / blender warp pairs) must
  leave exactly the frame and the per-pixel state of the throughput kernel, in both blend modes, on whole, ragged and
  tiny images;
* the speculative num_rendered hand-off (binning blob allocated for 1.25x the previous call's R, R delivered through
  mapped pinned memory by the tile scan) must re-run exactly when the guess was too small.
"""
import pytest
import torch

import cases
import refgpu

pytestmark = pytest.mark.gpu

SPLIT_CASES = {
    "f3d_256": lambda: cases.f3d_case(0, 256, 256, 1),          # the headline frame: 256 tiles, lists up to ~1900 records
    "f3d_96_ragged": lambda: cases.f3d_case(3, 96, 200, 5),     # 200 x 200: partial tiles on two borders
    "unit_wide": lambda: cases.unit_case(5, 20000, 250, 130, sh_degree=2, bg=(0.3, 0.1, 0.7)),
    "tiny": lambda: cases.unit_case(6, 40, 40, 24),             # lists shorter than one chunk, empty tiles
}


@pytest.mark.parametrize("exact", [False, True], ids=["fast", "exact"])
@pytest.mark.parametrize("name", sorted(SPLIT_CASES))
def test_split_kernel_equals_tile_kernel(name, exact, monkeypatch):
    from f3d_gaus_b200.diff_gof_rasterization import state_array
    monkeypatch.setenv("GOF_EXACT_BLEND", "1" if exact else "0")
    c = cases.case_to(SPLIT_CASES[name](), "cuda")
    outs = []
    for split_max in ("0", "1000000"):
        monkeypatch.setenv("GOF_FWD_SPLIT_MAX_TILES", split_max)
        o = refgpu.OursRun().forward(c)
        torch.cuda.synchronize()
        outs.append(o)
    a, b = outs
    assert a["num_rendered"] == b["num_rendered"]
    for k in ("out_color", "final_T"):
        assert torch.equal(a[k].view(torch.int32), b[k].view(torch.int32)), k
    assert torch.equal(a["n_contrib"], b["n_contrib"])
    if refgpu.ref_available():
        r = refgpu.RefRun().forward(c)
        assert torch.equal(b["n_contrib"], r["n_contrib"])
        if exact:
            assert torch.equal(b["out_color"].view(torch.int32), r["out_color"].view(torch.int32))


def test_split_kernel_overflow_poison():
    """Sync-free mode with an undersized binning blob, one small frame (64 tiles: the split kernel): the overflow is
    reported by finish() and the frame is NaN, never a plausible image -- like the tile kernel's."""
    from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace
    import test_gpu_batch as tb
    pc, cams, cfg = tb._scene(64, 128)
    R0, color0, *_ = tb._batch(pc, cams, 128, views=[2])
    ws = BatchWorkspace("cuda:0")
    ws.capacity_hint = 64                                       # far too small for one 128^2 frame of 4096 Gaussians
    R, color, *_ = tb._batch(pc, cams, 128, workspace=ws, views=[2])
    assert R is None and ws.finish() is None
    assert torch.isnan(color[:, :8]).all()
    R, color, *_ = tb._batch(pc, cams, 128, workspace=ws, views=[2])       # grown workspace: the re-run is exact
    assert ws.finish() == R0
    assert torch.equal(color.view(torch.int32), color0.view(torch.int32))


def test_speculative_handoff_reruns_exactly():
    """big frame, small frame, big frame again: the third call's speculative binning blob (sized from the small frame)
    overflows, and the exact re-run must reproduce the first call bit for bit."""
    big = cases.case_to(cases.f3d_case(0, 128, 256, 2), "cuda")
    small = cases.case_to(cases.unit_case(6, 40, 40, 24), "cuda")
    run = refgpu.OursRun()
    first = run.forward(big)
    tiny = run.forward(small)
    assert tiny["num_rendered"] * 2 + 8192 < first["num_rendered"]
    again = run.forward(big)
    assert again["num_rendered"] == first["num_rendered"]
    for k in ("out_color", "final_T", "point_list", "ranges", "n_contrib"):
        assert torch.equal(first[k].view(torch.int32), again[k].view(torch.int32)), k
    # and the calls after it speculate from the big frame again
    third = run.forward(big, decode_state=False)
    assert torch.equal(third["out_color"].view(torch.int32), first["out_color"].view(torch.int32))
