"""BASELINE configs[4], one GPU's share: 8 scenes x 8 orbit views at 512x512, 65,536 Gaussians per scene
(batch=64 scenes over 8 GPUs, scene-sharded).  Ours: one batched pass per scene (render_views, workspace mode);
reference: its own build, one Rasterizer::forward per frame.  Device time with CUDA events, scenes resident.

    python tools/bench_config5.py [--scenes 8] [--steps 5]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases, refgpu
from f3d_gaus_b200 import cameras, synthetic
from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace
from f3d_gaus_b200.gaussian_renderer import render_views

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=8)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--res", type=int, default=512)
a = ap.parse_args()
dev = torch.device("cuda", 0)
V, RES = 8, a.res
cams = cameras.orbit_cameras(V)
wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
cfg = synthetic.cfg_for(RES)
bg = torch.zeros(3, device=dev)
scenes = [{k: v.to(dev) for k, v in synthetic.f3d_like(s, 256).items()} for s in range(a.scenes)]
wss = [BatchWorkspace(dev) for _ in scenes]

def ours_step():
    for pc, ws in zip(scenes, wss):
        render_views(pc, 0, wv, fp, cc, bg, cfg, workspace=ws, epilogue=False)

def timed(fn, steps):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps

ours_step(); torch.cuda.synchronize()
for ws in wss:
    if ws.finish() is None: pass
ours_step(); torch.cuda.synchronize()
R = [ws.finish() for ws in wss]
assert all(r is not None for r in R)
t_ours = timed(ours_step, a.steps)
out = {"config": f"{a.scenes} scenes x {V} views at {RES}x{RES}, 65536 Gaussians per scene (one GPU's share of BASELINE configs[4])",
       "frames_per_step": a.scenes * V, "num_rendered_per_frame_mean": sum(map(sum, R)) / (a.scenes * V),
       "ours_ms_per_step": t_ours, "ours_frames_per_s": a.scenes * V / (t_ours * 1e-3)}
if refgpu.ref_available():
    flat = [[cases.make_case({k: v.cpu() for k, v in pc.items()}, cams.world_view[v], cams.full_proj[v], cams.centers[v],
                             W=RES, H=RES, fov_deg=13.164, device=dev) for v in range(V)] for pc in scenes]
    run = refgpu.RefRun()
    def ref_step():
        for sc in flat:
            for c in sc:
                run.forward(c, decode_state=False)
    t_ref = timed(ref_step, max(1, a.steps // 2))
    out.update({"reference_ms_per_step": t_ref, "reference_frames_per_s": a.scenes * V / (t_ref * 1e-3),
                "speedup": t_ref / t_ours})
print(json.dumps(out))
