"""Multi-view training step: forward + backward of V views of one Gaussian set (dev tool).

    python tools/train_bench.py [V] [S] [res]     ->  one JSON line

ours_batched  : rasterize_views_autograd (gof_forward_batch + gof_backward_batch, gradients summed over views in-kernel)
ours_per_view : GaussianRasterizer_GOF called V times under autograd (the reference-shaped API)
reference     : the unmodified reference build (oracle/_ref/libgof_ref.so), V x (forward + backward) + the V-way sum of
                its nine gradient tensors that autograd would do
All with the scene resident, seeded dL/dout, CUDA events, L2 flushed between steps."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases, refgpu
from f3d_gaus_b200 import cameras, synthetic
from f3d_gaus_b200.diff_gof_rasterization import (GaussianRasterizationSettings_GOF, GaussianRasterizer_GOF,
                                                 rasterize_views_autograd)

V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
res = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device("cuda", 0)
pc = synthetic.f3d_like(0, S)
orbit = cameras.orbit_cameras(8)
idx = [i % 8 for i in range(V)]
wv, fp, cc = orbit.world_view[idx].to(dev), orbit.full_proj[idx].to(dev), orbit.centers[idx].to(dev)
tanfov = math.tan(math.radians(13.164) / 2)
bg = torch.zeros(3, device=dev)
leaves = {k: pc[k][0].to(dev).requires_grad_(True) for k in ("xyz", "opacity", "scaling", "rotation")}
shs = torch.cat([pc["features_dc"][0], pc["features_rest"][0]], dim=1).to(dev).requires_grad_(True)
D = 1
dL = torch.randn(V, 9, res, res, generator=torch.Generator().manual_seed(3)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def clear():
    for t in list(leaves.values()) + [shs]:
        t.grad = None


def batched():
    clear()
    m2d = torch.zeros_like(leaves["xyz"], requires_grad=True)
    c, _ = rasterize_views_autograd(leaves["xyz"], m2d, leaves["opacity"], shs=shs, scales=leaves["scaling"],
                                    rotations=leaves["rotation"], bg=bg, viewmatrices=wv, projmatrices=fp, campos=cc,
                                    tanfovx=tanfov, tanfovy=tanfov, image_height=res, image_width=res, sh_degree=D)
    (c * dL).sum().backward()


def per_view():
    clear()
    m2d = torch.zeros_like(leaves["xyz"], requires_grad=True)
    loss = 0.0
    for v in range(V):
        rs = GaussianRasterizationSettings_GOF(res, res, tanfov, tanfov, 0.0, torch.zeros(1, device=dev), bg, 1.0,
                                               wv[v], fp[v], D, cc[v], False, False)
        c, _ = GaussianRasterizer_GOF(rs)(leaves["xyz"], m2d, leaves["opacity"], shs=shs, scales=leaves["scaling"],
                                          rotations=leaves["rotation"])
        loss = loss + (c * dL[v]).sum()
    loss.backward()


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    ms = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / n


line = {"what": f"training step: forward + backward of {V} views, {S * S} Gaussians, {res}x{res}", "views": V,
        "ours_batched_ms": round(timed(batched), 3), "ours_per_view_api_ms": round(timed(per_view), 3)}
if refgpu.ref_available():
    cs = [cases.make_case({k: t for k, t in pc.items()}, orbit.world_view[i], orbit.full_proj[i], orbit.centers[i], W=res, H=res,
                          fov_deg=13.164, device=dev) for i in idx]
    ref = refgpu.RefRun()

    def reference():
        acc = None
        for v, c in enumerate(cs):
            ref.forward(c, decode_state=False)
            g = ref.backward(c, dL[v])
            acc = g if acc is None else {k: acc[k] + g[k] for k in g}
        return acc

    line["reference_ms"] = round(timed(reference), 3)
    line["speedup_batched"] = round(line["reference_ms"] / line["ours_batched_ms"], 2)
    line["speedup_per_view_api"] = round(line["reference_ms"] / line["ours_per_view_api_ms"], 2)
line["steps_per_s_batched"] = round(1e3 / line["ours_batched_ms"], 1)
print(json.dumps(line))
