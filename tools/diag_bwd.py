"""Backward diagnostics: ours-vs-ref and ref-vs-ref error statistics per gradient tensor."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases, refgpu
from test_gpu_parity import FWD_CASES

def stats(a, b, rtol=1e-3):
    a, b = a.double(), b.double()
    if b.numel() == 0: return "empty"
    scale = b.pow(2).mean().sqrt().item()
    err = (a - b).abs()
    bad = err > rtol * (b.abs() + scale)
    rel = (a - b).norm().item() / max(b.norm().item(), 1e-30)
    return f"relL2={rel:.2e} out={bad.double().mean().item():.2e} maxerr={err.max().item():.2e} rms={scale:.2e}"

for name in sys.argv[1:] or ["unit_p4096_200x136", "f3d_s64_r256_view2", "f3d_s256_r256_view2"]:
    c = FWD_CASES[name]("cuda")
    dL = cases.grad_seed(c)
    refs = []
    for _ in range(2):
        r = refgpu.RefRun(); r.forward(c, decode_state=False); refs.append(r.backward(c, dL))
    o = refgpu.OursRun(); o.forward(c, decode_state=False); ours = o.backward(c, dL)
    o2 = refgpu.OursRun(); o2.forward(c, decode_state=False); ours2 = o2.backward(c, dL)
    print("==", name)
    for k in refgpu.GRAD_NAMES:
        print(f"  {k:18s} ours-ref: {stats(ours[k], refs[0][k])}")
        print(f"  {'':18s} ref-ref : {stats(refs[1][k], refs[0][k])}")
        print(f"  {'':18s} ours-ours: {stats(ours2[k], ours[k])}")
    # worst dL_dscales rows
    k = "dL_dscales"
    err = (ours[k] - refs[0][k]).abs().max(dim=1).values
    idx = err.topk(5).indices
    for i in idx.tolist():
        print("   worst", k, i, "ours", ours[k][i].tolist(), "ref", refs[0][k][i].tolist(), "ref2", refs[1][k][i].tolist(),
              "scale", c["scales"][i].tolist(), "v2g ours", ours["dL_dview2gaussian"][i].tolist(), "v2g ref", refs[0]["dL_dview2gaussian"][i].tolist())
