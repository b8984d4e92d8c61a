"""Predictor output head: the fused kernel against the reference's torch op sequence on the same GPU (dev tool).

    python tools/bench_head.py [BV] [res]      ->  one JSON line (also written by tools/gpu_round.sh into profiles/)

Algorithmic bytes per pixel: (C + 1) * 4 read (network planes + depth) + (14 + 9 * sh) * 4 written; the kernel is a pure
HBM stream, so `frac` = achieved / measured copy bandwidth (MEASURED_PEAKS.json, else the profiling guide's fallback)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import head_torch_ref, make_head_golden as mk
from f3d_gaus_b200.predictor_head import PredictorHead

BV = int(sys.argv[1]) if len(sys.argv) > 1 else 8
res = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
cfg = {"model": {"training_resolution": res, "fov": 13.164, "inverted_x": False, "inverted_y": True, "max_sh_degree": 1,
                 "isotropic": False, "origin_distances": False, "network_with_offset": True, "network_without_offset": False}}
net, depth, _, v2w, quat = mk.make_inputs(BV, 1, res, True, 1, False, 0)
net, depth, v2w, quat = net.to(dev), depth.to(dev), v2w.to(dev), quat.to(dev)
ours, ref = PredictorHead(cfg, dev), head_torch_ref.TorchHead(cfg, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    ms = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / n


t_call = timed(lambda: ours(net, depth, v2w, quat, BV, 1))          # through the Python wrapper (7 output allocations + ctypes)
t_ref = timed(lambda: ref(net, depth, v2w, quat, BV, 1))
# the kernel alone: the same call captured in a CUDA graph and replayed (no host work between the events)
graph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    ours(net, depth, v2w, quat, BV, 1)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=side):
        keep = ours(net, depth, v2w, quat, BV, 1)
torch.cuda.synchronize()
t_ours = timed(graph.replay)
C = net.shape[1]
bytes_alg = BV * res * res * ((C + 1) * 4 + (14 + 9) * 4)
peak, src = 6650.0, "fallback (B200_PROFILING.md)"
try:
    peak, src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
except Exception:
    pass
gbps = bytes_alg / (t_ours * 1e-3) / 1e9
print(json.dumps({"what": "predictor output head (gof_predictor_head), L2 flushed between launches", "images": BV, "res": res,
                  "channels": C, "kernel_us": round(t_ours * 1e3, 2), "call_us": round(t_call * 1e3, 2),
                  "torch_op_sequence_us": round(t_ref * 1e3, 2), "speedup_call": round(t_ref / t_call, 1), "algorithmic_bytes": bytes_alg,
                  "roofline": {"bound": "hbm", "achieved": round(gbps, 1), "peak": peak, "unit": "GB/s",
                               "frac": round(gbps / peak, 3), "peak_source": src}}))
