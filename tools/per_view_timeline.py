"""Where does a frame of the one-frame-per-call API go?  (dev tool)

Per frame: wall time of the back-to-back loop, the GPU time of its stages (gof_profile events), the host time spent
inside the library call (launches + the wait for num_rendered) and in the Python around it.
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from f3d_gaus_b200 import _lib, cameras, synthetic
from f3d_gaus_b200.diff_gof_rasterization import _C
from f3d_gaus_b200.gaussian_renderer import render_predicted_more_v2_gof

dev = torch.device("cuda", 0)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pc = {k: v.to(dev) for k, v in synthetic.f3d_like(0, 256).items()}
cams = cameras.orbit_cameras(8)
wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
cfg = synthetic.cfg_for(res)
bg = torch.zeros(3, device=dev)

inner = [0.0, 0]
real = _lib.lib.gof_forward
def timed(*a):
    t0 = time.perf_counter()
    r = real(*a)
    inner[0] += time.perf_counter() - t0
    inner[1] += 1
    return r

def loop(n):
    with torch.no_grad():
        for i in range(n):
            v = i % 8
            render_predicted_more_v2_gof(pc, 0, wv[v:v + 1], fp[v:v + 1], cc[v:v + 1], bg, cfg)

def measure(n, label):
    loop(32); torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(n); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / n
    _lib.profile_enable(0, True); _lib.profile_read(0)
    loop(n); torch.cuda.synchronize()
    pr = _lib.profile_read(0); _lib.profile_enable(0, False)
    gpu = {k: v / pr["fwd_calls"] * 1e3 for k, v in pr["fwd_ms"].items()}
    _lib.lib.gof_forward = timed
    inner[0] = 0.0; inner[1] = 0
    t0 = time.perf_counter(); loop(n); torch.cuda.synchronize(); wall2 = (time.perf_counter() - t0) / n
    _lib.lib.gof_forward = real
    print(f"{label}: wall {wall*1e6:.1f} us/frame ({1/wall:.0f} frames/s); GPU stages us {({k: round(v, 1) for k, v in gpu.items()})} "
          f"sum {sum(gpu.values()):.1f}; in-library host time {inner[0]/inner[1]*1e6:.1f} us, python around it {(wall2 - inner[0]/n)*1e6:.1f} us")

measure(800, f"render_predicted_more_v2_gof {res}^2")
# host cost of the pieces
def t_host(fn, n=2000):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return t * 1e6
print("torch.empty(9,H,W): %.1f us" % t_host(lambda: torch.empty((9, res, res), dtype=torch.float32, device=dev)))
print("torch.empty(bytes): %.1f us" % t_host(lambda: torch.empty(20_000_000, dtype=torch.uint8, device=dev)))
print("zeros_like(xyz)   : %.1f us" % t_host(lambda: torch.zeros_like(pc["xyz"][0])))
