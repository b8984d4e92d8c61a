"""Host-side profile of the per-frame reference-shaped call (dev tool)."""
import cProfile, pstats, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from f3d_gaus_b200 import cameras, synthetic
from f3d_gaus_b200.gaussian_renderer import render_predicted_more_v2_gof
dev = torch.device("cuda", 0)
pc = {k: v.to(dev) for k, v in synthetic.f3d_like(0, 256).items()}
cams = cameras.orbit_cameras(8)
wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
cfg = synthetic.cfg_for(256)
bg = torch.zeros(3, device=dev)
def loop(n):
    with torch.no_grad():
        for i in range(n):
            v = i % 8
            o = render_predicted_more_v2_gof(pc, 0, wv[v:v + 1], fp[v:v + 1], cc[v:v + 1], bg, cfg)
    return o
loop(16); torch.cuda.synchronize()
t0 = time.perf_counter(); loop(400); torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 400
print(f"per frame wall {t * 1e6:.1f} us")
pr = cProfile.Profile(); pr.enable(); loop(400); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
