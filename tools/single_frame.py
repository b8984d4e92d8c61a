"""One-frame-per-call forwards of the headline scene, per orbit view: GPU time per frame from CUDA events over a
back-to-back loop of `_C.rasterize_gaussians` calls, and the stage times of gof_profile (dev tool; also the ncu target
for the single-frame launches: `ncu -k regex:render_fwd ... python tools/single_frame.py 256 3`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from f3d_gaus_b200 import _lib, cameras, synthetic
from f3d_gaus_b200.diff_gof_rasterization import _C

dev = torch.device("cuda", 0)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pc = {k: v.to(dev) for k, v in synthetic.f3d_like(0, 256).items()}
cams = cameras.orbit_cameras(8)
wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
cfg = synthetic.cfg_for(res)
import math
tan = math.tan(cfg["model"]["fov"] * math.pi / 360)
bg = torch.zeros(3, device=dev)
e = torch.Tensor([])
xyz, op, sc, rot = pc["xyz"][0].contiguous(), pc["opacity"][0].contiguous(), pc["scaling"][0].contiguous(), pc["rotation"][0].contiguous()
shs = torch.cat([pc["features_dc"][0], pc["features_rest"][0]], dim=1).contiguous()
D = cfg["model"]["max_sh_degree"]

def frame(v):
    return _C.rasterize_gaussians(bg, xyz, e, op, sc, rot, 1.0, e, e, wv[v], fp[v], tan, tan, 0.0, e, res, res, shs, D, cc[v], False, False)

tot = 0.0
for v in range(8):
    for _ in range(3): R = frame(v)[0]
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): frame(v)
    t.record(); torch.cuda.synchronize()
    us = s.elapsed_time(t) / iters * 1e3
    _lib.profile_enable(0, True); _lib.profile_read(0)
    for _ in range(max(iters // 4, 1)): frame(v)
    torch.cuda.synchronize()
    pr = _lib.profile_read(0); _lib.profile_enable(0, False)
    st = {k: round(x / pr["fwd_calls"] * 1e3, 1) for k, x in pr["fwd_ms"].items()}
    tot += us
    print(f"view {v}: R={int(R)} {us:.1f} us/frame  stages {st}")
print(f"mean {tot / 8:.1f} us/frame -> {8e6 / tot:.0f} frames/s")
