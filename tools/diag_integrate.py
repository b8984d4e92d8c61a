"""Integration-path diagnostics: where ours and the reference's Rasterizer::integrate differ."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases, refgpu
from test_gpu_integrate import CASES, query_points

for name in sys.argv[1:] or list(CASES):
    c = CASES[name]("cuda")
    pts = query_points(c, 30000)
    ref = refgpu.ref_integrate(c, pts)
    ours = refgpu.ours_integrate(c, pts)
    print("==", name, "R", ours["num_rendered"], ref["num_rendered"])
    for ch in (0, 1, 2, 6, 7, 8):
        d = ours["out_color"][ch] - ref["out_color"][ch]
        bad = d.abs() > 1e-4
        print(f"  ch{ch}: bad {int(bad.sum())}/{d.numel()}  max|d| {d.abs().max().item():.3e}  ours>ref {int((d > 1e-4).sum())} ours<ref {int((d < -1e-4).sum())}")
    d = ours["alpha_integrated"] - ref["alpha_integrated"]
    bad = d.abs() > 1e-4
    print(f"  alpha_integrated: bad {int(bad.sum())}/{d.numel()} max {d.abs().max().item():.3e} mean {d.abs().mean().item():.3e} ours>ref {int((d>1e-4).sum())} ours<ref {int((d<-1e-4).sum())}")
    d = (ours["color_integrated"] - ref["color_integrated"]).abs().max(dim=1).values
    print(f"  color_integrated: bad {int((d > 1e-4).sum())}/{d.numel()} max {d.max().item():.3e}")
    r2 = refgpu.ref_integrate(c, pts)
    print("  ref-vs-ref identical:", torch.equal(r2["out_color"], ref["out_color"]), torch.equal(r2["alpha_integrated"], ref["alpha_integrated"]))
    import time
    def timeit(fn, it=5):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(it): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / it * 1e3
    print(f"  time: ours {timeit(lambda: refgpu.ours_integrate(c, pts)):.2f} ms   ref {timeit(lambda: refgpu.ref_integrate(c, pts)):.2f} ms")
