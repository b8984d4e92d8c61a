"""Where the end-to-end step goes: H2D, batched render, D2H, host-side launch cost (dev tool)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from f3d_gaus_b200 import cameras, synthetic
from f3d_gaus_b200.gaussian_renderer import render_views, HostFrameSink
from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace

dev = torch.device("cuda", 0)
pc_cpu = synthetic.f3d_like(0, 256)
host_pc = {k: v.pin_memory() for k, v in pc_cpu.items()}
dev_pc = {k: torch.empty_like(v, device=dev) for k, v in pc_cpu.items()}
cams = cameras.orbit_cameras(8)
wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
cfg = synthetic.cfg_for(256)
bg = torch.zeros(3, device=dev)
ws = BatchWorkspace(dev)
out_dev = torch.empty((8, 5, 256, 256), device=dev)
out_host = torch.empty((8, 5, 256, 256)).pin_memory()

def wall(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def host_only(fn, n=30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize(); return t

def h2d():
    for k in host_pc: dev_pc[k].copy_(host_pc[k], non_blocking=True)
def render():
    return render_views(dev_pc, 0, wv, fp, cc, bg, cfg, workspace=ws, epilogue=False)
def pack(o):
    out_dev[:, 0:3].copy_(o["render"]); out_dev[:, 3:4].copy_(o["rendered_depth"]); out_dev[:, 4:5].copy_(o["rendered_alpha"])
def d2h():
    out_host.copy_(out_dev, non_blocking=True)
h2d(); o = render(); torch.cuda.synchronize(); ws.finish(); o = render(); torch.cuda.synchronize(); ws.finish()
print(f"H2D 6.0 MB            : {wall(h2d):.3f} ms")
print(f"render_views (8 views): {wall(render):.3f} ms   host-side launch cost {host_only(render):.3f} ms")
print(f"pack 3 copies         : {wall(lambda: pack(o)):.3f} ms")
print(f"D2H 10.5 MB           : {wall(d2h):.3f} ms")
def full():
    h2d(); pack(render()); d2h()
print(f"full step             : {wall(full):.3f} ms")
for chunks, zc, cl in ((2, False, False), (1, False, False), (1, True, False), (1, True, True), (2, True, True)):
    sk = HostFrameSink(8, 256, 256, dev, chunks=chunks, zero_copy=zc, channels_last=cl)
    def fullk():
        h2d(); sk.render(dev_pc, 0, wv, fp, cc, bg, cfg)
    fullk(); torch.cuda.synchronize(); sk.finish(); fullk(); torch.cuda.synchronize(); sk.finish()
    how = ("kernel stores to pinned host, " + ("[V,H,W,5]" if cl else "[V,5,H,W]")) if zc else "pack + DMA copy"
    print(f"full step, sink chunks={chunks} {how:42s}: {wall(fullk):.3f} ms  host-side {host_only(fullk):.3f} ms")
for cl in (False, True):
    oh = torch.empty((8, 256, 256, 5)).pin_memory().permute(0, 3, 1, 2) if cl else out_host
    def render_sink():
        return render_views(dev_pc, 0, wv, fp, cc, bg, cfg, workspace=ws, epilogue=False, sink=oh)
    print(f"render_views + kernel stores to pinned host {'[V,H,W,5]' if cl else '[V,5,H,W]'} (no H2D): {wall(render_sink):.3f} ms")
od = torch.empty((8, 5, 256, 256), device=dev)
print(f"render_views + kernel stores to a DEVICE sink: {wall(lambda: render_views(dev_pc, 0, wv, fp, cc, bg, cfg, workspace=ws, epilogue=False, sink=od)):.3f} ms")
