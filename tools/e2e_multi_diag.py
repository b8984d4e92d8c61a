"""Where the end-to-end step goes at N GPUs (dev tool; run under torchrun like bench.py).
Variants of the e2e step, host-timed per step with the ranks aligned by a barrier, max and mean over ranks."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from f3d_gaus_b200 import cameras, synthetic, sharding
from f3d_gaus_b200.gaussian_renderer import HostFrameSink, render_views
from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group(backend="nccl", device_id=dev)
info = {"rank": rank, "affinity_cpus": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    info["nvml_cpu_affinity_words"] = [hex(x) for x in pynvml.nvmlDeviceGetCpuAffinity(h, 4)]
    info["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
except Exception as e:      # noqa: BLE001
    info["nvml"] = repr(e)
if os.environ.get("DIAG_BIND") == "1":
    try:
        words = pynvml.nvmlDeviceGetCpuAffinity(h, 16)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound_to"] = len(cpus)
    except Exception as e:  # noqa: BLE001
        info["bind_error"] = repr(e)

pc_cpu = synthetic.f3d_like(rank, 256)
host_pc = {k: v.pin_memory() for k, v in pc_cpu.items()}
dev_pc = {k: torch.empty_like(v, device=dev) for k, v in pc_cpu.items()}
for k in host_pc:
    dev_pc[k].copy_(host_pc[k])
cams = cameras.orbit_cameras(8)
wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
cfg = synthetic.cfg_for(256)
bg = torch.zeros(3, device=dev)
sink = HostFrameSink(8, 256, 256, dev)
ws = BatchWorkspace(dev)
peer = sharding.PeerFrameGather(world, 8, 256, 256, dev) if world > 1 else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def h2d():
    for k in host_pc:
        dev_pc[k].copy_(host_pc[k], non_blocking=True)


def render_sink():
    sink.render(dev_pc, 0, wv, fp, cc, bg, cfg)


def render_only():
    render_views(dev_pc, 0, wv, fp, cc, bg, cfg, workspace=ws, epilogue=False)


def exchange():
    if peer:
        peer.push(sink.last_raster, first_scene=rank)


variants = {
    "full (h2d + render + sink + exchange)": lambda: (h2d(), render_sink(), exchange()),
    "h2d + render + sink": lambda: (h2d(), render_sink()),
    "render + sink": render_sink,
    "render only": render_only,
    "h2d only": h2d,
    "exchange only": exchange,
}
render_sink(); torch.cuda.synchronize(); sink.finish(); render_only(); torch.cuda.synchronize(); ws.finish()
render_sink(); torch.cuda.synchronize(); sink.finish(); render_only(); torch.cuda.synchronize(); ws.finish()
out = {}
for name, fn in variants.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([sum(ts) / len(ts), sorted(ts)[len(ts) // 2]], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        out[name] = {"mean_ms_max_over_ranks": round(float(mx[0]), 3), "mean_ms_avg_over_ranks": round(float(sm[0]) / world, 3),
                     "median_ms_max_over_ranks": round(float(mx[1]), 3)}
    else:
        out[name] = {"mean_ms": round(float(t[0]), 3), "median_ms": round(float(t[1]), 3)}
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, info)
else:
    gathered = [info]
if rank == 0:
    print(json.dumps({"world": world, "bind": os.environ.get("DIAG_BIND", "0"), "variants": out, "ranks": gathered}, indent=1))
if world > 1:
    dist.destroy_process_group()
