#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (BASELINE.json: "256x256 NVS frames/sec @65k Gaussians;
renderCUDA HBM GB/s vs roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload nvs256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload `nvs256` (BASELINE configs[1]): one seeded "f3d-like" scene of 65,536 Gaussians (SH degree 1)
per GPU, rendered at 256x256 from the 8 orbit views of the cycle-aggregative loop.  One STEP = one
pass of the hot path over that batch = 8 full forwards (preprocess -> binning -> blend), issued as
ONE batched call (gof_forward_batch).  `value` = frames/s over all GPUs with the scene resident in
HBM; `e2e` = the same frames through the loop's public API (gaussian_renderer.render_views) starting
from pinned HOST buffers, with the H2D copy of the Gaussian set and the D2H copy of the rendered
rgb/depth/alpha inside the timed region.  `per_view_api` repeats both measurements one frame per call
through the reference-shaped functions (_C.rasterize_gaussians / render_predicted_more_v2_gof).
Timing: CUDA events per step on the launch stream, an L2 flush (256 MB write) between steps outside
the events, max over ranks.

`--impl reference` times the UNMODIFIED reference rasterizer (oracle/_ref/libgof_ref.so: its CUDA
sources compiled for sm_100a -- the reference has no CPU implementation of this path) on the same
workload; if that library is absent it falls back to the CPU oracle port.  The product arm never
touches oracle/: only the `cpu_baseline` leg (rank 0, N=1) and the reference arm do.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VIEWS = 8
P_SIDE = 256            # 256*256 = 65,536 Gaussians
RES = 256
L2_FLUSH_BYTES = 256 << 20


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        # keep stdout for the one JSON line: NCCL_DEBUG=VERSION printf()s its banner to stdout, other levels log to the file
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world: int) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, world: int) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_workload(rank: int, device):
    from f3d_gaus_b200 import cameras, synthetic
    import cases
    pc_cpu = synthetic.f3d_like(seed=rank, S=P_SIDE)
    cams = cameras.orbit_cameras(VIEWS)
    cfg = synthetic.cfg_for(RES)
    flat = [cases.make_case(pc_cpu, cams.world_view[v], cams.full_proj[v], cams.centers[v], W=RES, H=RES,
                            fov_deg=cfg["model"]["fov"], device=device) for v in range(VIEWS)]
    return pc_cpu, cams, cfg, flat


def algorithmic_bytes_render_fwd(P, R, W, H):
    """A_render_fwd = 8 T + 60 R + 12 P + 60 N per frame."""
    T = ((W + 15) // 16) * ((H + 15) // 16)
    return 8 * T + 60 * R + 12 * P + 60 * W * H


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles():
    """(dram bytes per render launch, issue-side figures) from the committed ncu --set full summary, if present."""
    path = os.path.join(ROOT, "profiles", "render_fwd_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        issue = {k: d[k] for k in ("issue_active_pct", "active_lanes_per_instruction", "warp_instructions") if d.get(k) is not None}
        return d.get("dram_bytes_per_launch"), (issue or None)
    except Exception:
        return None, None


# ------------------------------------------------------------------------------ arms --------
class OursArm:
    """The product.  A step = the 8 views of the scene in ONE batched pass (gof_forward_batch, sync-free
    workspace mode); `per_view_step` = the same 8 frames through the reference-shaped one-frame call."""
    name = "ours"
    KERNELS_PER_STEP = 5        # preprocess, tile_scan, scatter, tile_sort_gather, render_fwd (one launch each per batch)

    def __init__(self, device):
        from f3d_gaus_b200 import _lib
        from f3d_gaus_b200.diff_gof_rasterization import _C, BatchWorkspace, rasterize_views
        self._lib, self._C, self.device = _lib, _C, device
        self.rasterize_views = rasterize_views
        self.ws = BatchWorkspace(device)
        self.ws_e2e = BatchWorkspace(device)
        self.empty = torch.Tensor([])
        self.batch = None
        self.out = None
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.gathered = None
        self.sink = None
        self.peer = None
        self.exchange = None

    def _prepare(self, flat):
        c = flat[0]
        self.batch = dict(vm=torch.stack([f["viewmatrix"] for f in flat]).contiguous(),
                          pm=torch.stack([f["projmatrix"] for f in flat]).contiguous(),
                          cam=torch.stack([f["campos"] for f in flat]).contiguous())
        self.out = torch.empty((len(flat), 9, c["H"], c["W"]), dtype=torch.float32, device=self.device)

    def step(self, flat):
        if self.batch is None:
            self._prepare(flat)
        c, b = flat[0], self.batch
        self.rasterize_views(c["bg"], c["means3D"], None, c["opacities"], c["scales"], c["rotations"],
                             c["scale_modifier"], b["vm"], b["pm"], c["tanfovx"], c["tanfovy"], c["kernel_size"],
                             c["H"], c["W"], c["shs"], c["D"], b["cam"], workspace=self.ws, out_color=self.out)
        return None

    def finish(self):
        """Per-view num_rendered of the last step (one mailbox read; None => the blob overflowed)."""
        return self.ws.finish()

    def frame(self, c):
        e = self.empty
        return self._C.rasterize_gaussians(c["bg"], c["means3D"], e, c["opacities"], c["scales"], c["rotations"],
                                           c["scale_modifier"], e, e, c["viewmatrix"], c["projmatrix"], c["tanfovx"],
                                           c["tanfovy"], c["kernel_size"], e, c["H"], c["W"], c["shs"], c["D"],
                                           c["campos"], False, False)

    def per_view_step(self, flat):
        R = 0
        for c in flat:
            R += int(self.frame(c)[0])
        return R

    def profile(self, on):
        self._lib.profile_enable(self.device.index, on)

    def profile_read(self):
        return self._lib.profile_read(self.device.index)

    def e2e_step(self, host_pc, dev_pc, cams_dev, cfg, bg, out_dev, out_host):
        """Public API of the loop from pinned host buffers: H2D of the Gaussian set, batched render of the 8 views
        (gaussian_renderer.HostFrameSink), D2H of rgb/depth/alpha into pinned host memory."""
        from f3d_gaus_b200.gaussian_renderer import HostFrameSink
        if self.sink is None:
            # frames are stored into pinned host memory by the blend kernel itself (gof_set_frame_sink);
            # GOF_BENCH_READBACK=dma selects the packed DMA copy instead (tools/e2e_breakdown.py compares them)
            self.sink = HostFrameSink(VIEWS, RES, RES, self.device, chunks=1,
                                      zero_copy=os.environ.get("GOF_BENCH_READBACK", "kernel") != "dma")
        dev_pc = host_pc.upload()                     # gaussian_renderer.PinnedScene: one copy for the whole set
        self.sink.render(dev_pc, 0, cams_dev[0], cams_dev[1], cams_dev[2], bg, cfg)
        if self.world > 1:
            # the path's one exchange step: every rank receives all scenes' frames -- one kernel that
            # packs the consumed channels and stores them into every rank's buffer over NVLink peer memory
            # (sharding.PeerFrameGather); NCCL all_gather if symmetric memory cannot be set up
            from f3d_gaus_b200 import sharding
            if self.peer is None:
                try:
                    self.peer = sharding.PeerFrameGather(self.world, VIEWS, RES, RES, self.device)
                    self.exchange = "fused pack + all-gather over NVLink peer memory (gof_pack_gather, torch symmetric memory)"
                except Exception as ex:      # noqa: BLE001
                    self.peer = False
                    self.exchange = f"NCCL all_gather_into_tensor (peer memory unavailable: {type(ex).__name__})"
            if self.peer:
                self.gathered = self.peer.push(self.sink.last_raster, first_scene=int(os.environ.get("RANK", "0")))
            else:
                r = self.sink.last_raster
                self.gathered = sharding.gather_frames(sharding.pack_frames(r[None, :, 0:3], r[None, :, 6:7], r[None, :, 7:8]), self.world)

    def e2e_finish(self):
        return self.sink.finish()

    def e2e_per_view_step(self, host_pc, dev_pc, cams_dev, cfg, bg, out_dev, out_host):
        """The same through the reference's own one-frame function (render_predicted_more_v2_gof)."""
        from f3d_gaus_b200.gaussian_renderer import render_predicted_more_v2_gof
        dev_pc = host_pc.upload()
        with torch.no_grad():
            for v in range(VIEWS):
                o = render_predicted_more_v2_gof(dev_pc, 0, cams_dev[0][v:v + 1], cams_dev[1][v:v + 1],
                                                 cams_dev[2][v:v + 1], bg, cfg)
                out_dev[v, 0:3].copy_(o["render"])
                out_dev[v, 3:4].copy_(o["rendered_depth"])
                out_dev[v, 4:5].copy_(o["rendered_alpha"])
        out_host.copy_(out_dev, non_blocking=True)


class ReferenceArm:
    """The unmodified reference rasterizer (CUDA, sm_100a build) through its own C++ entry points."""
    name = "reference"

    def __init__(self, device):
        import refgpu
        self.refgpu, self.device = refgpu, device
        self.run = refgpu.RefRun()
        self.lib = self.run.lib
        self.out_color = None

    def frame(self, c):
        P, W, H = c["means3D"].shape[0], c["W"], c["H"]
        # the reference glue allocates + fills these per call (rasterize_points.cu:72-73)
        out_color = torch.full((9, H, W), 0.0, dtype=torch.float32, device=self.device)
        radii = torch.full((P,), 0, dtype=torch.int32, device=self.device)
        p = lambda t: t.data_ptr()
        R = self.lib.ref_forward(self.run.state, P, c["D"], c["shs"].shape[1], p(c["bg"]), W, H, p(c["means3D"]),
                                 p(c["shs"]), None, p(c["opacities"]), p(c["scales"]), c["scale_modifier"],
                                 p(c["rotations"]), None, None, p(c["viewmatrix"]), p(c["projmatrix"]), p(c["campos"]),
                                 c["tanfovx"], c["tanfovy"], c["kernel_size"], None, 0, p(out_color), p(radii), 0)
        if R < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return R, out_color, radii

    def step(self, flat):
        R = 0
        for c in flat:
            R += self.frame(c)[0]
        return R

    def profile(self, on):
        pass

    def profile_read(self):
        return None

    def finish(self):
        return None

    def e2e_finish(self):
        return None

    def e2e_step(self, host_pc, dev_pc, cams_dev, cfg, bg, out_dev, out_host):
        import math
        dev_pc = host_pc.upload()
        tanfov = math.tan(cfg["model"]["fov"] * math.pi / 360)
        shs = torch.cat([dev_pc["features_dc"][0], dev_pc["features_rest"][0]], dim=1).contiguous()
        for v in range(VIEWS):
            c = {"bg": bg, "means3D": dev_pc["xyz"][0], "shs": shs, "opacities": dev_pc["opacity"][0],
                 "scales": dev_pc["scaling"][0], "rotations": dev_pc["rotation"][0], "scale_modifier": 1.0,
                 "viewmatrix": cams_dev[0][v], "projmatrix": cams_dev[1][v], "campos": cams_dev[2][v],
                 "tanfovx": tanfov, "tanfovy": tanfov, "kernel_size": 0.0, "W": RES, "H": RES,
                 "D": cfg["model"]["max_sh_degree"]}
            _, color, _ = self.frame(c)
            out_dev[v, 0:3].copy_(color[0:3])
            out_dev[v, 3:4].copy_(color[6:7])
            out_dev[v, 4:5].copy_(color[7:8])
        out_host.copy_(out_dev, non_blocking=True)


def cpu_oracle_frames_per_s(flat_cpu, budget_s=12.0, max_frames=400):
    """CPU port (oracle/gof_oracle.c, OpenMP over all host cores) on a bounded sample of the same workload: the 8
    frames of a step, cycled until ~`budget_s` seconds of CPU work are done."""
    import oracle_cpu
    cs = [oracle_cpu.case_to_numpy(c) for c in flat_cpu]
    oracle_cpu.forward_all(cs[0])          # warm (page-in, thread pool)
    n, t0 = 0, time.perf_counter()
    while n < max_frames and (time.perf_counter() - t0) < budget_s:
        oracle_cpu.forward_all(cs[n % len(cs)])
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, int(oracle_cpu.lib().oracle_num_threads()), n, dt


def run_cpu_reference_arm(args):
    """Reference arm when oracle/_ref/libgof_ref.so is absent: the CPU oracle port on host cores."""
    import cases
    from f3d_gaus_b200 import cameras, synthetic
    pc_cpu = synthetic.f3d_like(seed=0, S=P_SIDE)
    cams = cameras.orbit_cameras(VIEWS)
    flat = [cases.make_case(pc_cpu, cams.world_view[v], cams.full_proj[v], cams.centers[v], W=RES, H=RES,
                            fov_deg=13.164) for v in range(VIEWS)]
    fps, cores, n, dt = cpu_oracle_frames_per_s(flat, budget_s=min(20.0, 1.0 * max(1, args.steps)))
    line = {"impl": "reference", "metric": "nvs_frames_per_sec_256x256_65k_gaussians", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * VIEWS / fps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "nvs256: 65536 f3d-like Gaussians, 8 orbit views, 256x256, forward"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n} frames in {dt:.1f}s (oracle/gof_oracle.c, OpenMP)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nvs256")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    if args.impl == "reference":
        import refgpu
        if not refgpu.ref_available():
            if int(os.environ.get("RANK", "0")) == 0:
                run_cpu_reference_arm(args)
            return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path is CUDA-only (no CPU fallback)")

    rank, world, local = dist_setup(args.gpus)
    device = torch.device("cuda", local)
    pc_cpu, cams, cfg, flat = build_workload(rank, device)
    arm = OursArm(device) if args.impl == "ours" else ReferenceArm(device)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
    P, K, W_ = P_SIDE * P_SIDE, args.steps, args.warmup

    def timed_device(step_fn):
        """K steps, CUDA events around each on the launch stream, L2 flush outside the events; max over ranks."""
        for _ in range(W_):
            step_fn()
        torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        barrier(world)
        for i in range(K):
            flush.zero_()                      # L2 flush between timed iterations, outside the events
            starts[i].record()
            step_fn()
            ends[i].record()
        barrier(world)
        return max_over_ranks(sum(s.elapsed_time(e) for s, e in zip(starts, ends)), world)

    def timed_e2e(step_fn):
        """Host clock around each step (host buffers in, host buffers out, device idle on both sides)."""
        for _ in range(W_):
            step_fn()
        torch.cuda.synchronize()
        barrier(world)
        ms = 0.0
        for i in range(K):
            flush.zero_()
            barrier(world)          # ranks start each step together: the exchange step would otherwise time their skew
            t0 = time.perf_counter()
            step_fn()
            torch.cuda.synchronize()
            ms += (time.perf_counter() - t0) * 1e3
        barrier(world)
        return max_over_ranks(ms, world)

    # ---------------- kernel-level throughput: inputs resident in HBM -------------------------
    R_views = None
    if args.impl == "ours":
        arm.step(flat)
        torch.cuda.synchronize()
        if arm.finish() is None:               # first call sized the binning blob; overflow => it has been grown
            arm.step(flat)
            torch.cuda.synchronize()
            assert arm.finish() is not None
    arm.profile(True)
    if args.impl == "ours":
        arm.profile_read()
    sampler = ClockSampler(local) if rank == 0 else None
    total_ms = timed_device(lambda: arm.step(flat))
    clocks = sampler.stop() if sampler else None
    prof = arm.profile_read()
    arm.profile(False)
    if args.impl == "ours":
        R_views = arm.finish()
        assert R_views is not None, "binning blob overflowed inside the timed region"
        R_step = sum(R_views)
    else:
        R_step = arm.step(flat)
    frames = VIEWS * K * world
    value = frames / (total_ms * 1e-3)
    per_view_value = None
    if args.impl == "ours":
        pv_ms = timed_device(lambda: arm.per_view_step(flat))
        per_view_value = frames / (pv_ms * 1e-3)

    # ---------------- end to end through the public API, host buffers -------------------------
    from f3d_gaus_b200.staging import PinnedScene      # pure torch (also re-exported by gaussian_renderer)
    host_pc = PinnedScene(pc_cpu, device)          # both arms: the set in one pinned slab, one H2D copy per step
    dev_pc = host_pc.dev
    cams_dev = (cams.world_view.to(device), cams.full_proj.to(device), cams.centers.to(device))
    bg = torch.zeros(3, device=device)
    out_dev = torch.empty((VIEWS, 5, RES, RES), dtype=torch.float32, device=device)
    out_host = torch.empty((VIEWS, 5, RES, RES), dtype=torch.float32).pin_memory()
    e2e_args = (host_pc, dev_pc, cams_dev, cfg, bg, out_dev, out_host)
    if args.impl == "ours":
        arm.e2e_step(*e2e_args)
        torch.cuda.synchronize()
        if arm.e2e_finish() is None:
            arm.e2e_step(*e2e_args)
    e2e_ms = timed_e2e(lambda: arm.e2e_step(*e2e_args))
    if args.impl == "ours":
        assert arm.e2e_finish() is not None, "binning blob overflowed inside the e2e region"
    e2e_value = frames / (e2e_ms * 1e-3)
    e2e_pv_value = None
    if args.impl == "ours":
        e2e_pv_ms = timed_e2e(lambda: arm.e2e_per_view_step(*e2e_args))
        e2e_pv_value = frames / (e2e_pv_ms * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in pc_cpu.values())
    d2h = out_host.numel() * out_host.element_size()

    # ---------------- roofline of the dominant kernel (the forward blend) ---------------------
    roofline = None
    launches = None
    peak, peak_src = measured_peak_gbs()
    if args.impl == "ours" and prof and prof["fwd_calls"]:
        n_calls = prof["fwd_calls"]            # one call = one batch of VIEWS frames = one launch of every kernel
        blend_ms = prof["fwd_ms"]["blend"] / n_calls
        abytes = sum(algorithmic_bytes_render_fwd(P, r, RES, RES) for r in R_views)
        achieved = abytes / (blend_ms * 1e-3) / 1e9
        traffic, issue = traffic_from_profiles()
        roofline = {"bound": "hbm", "kernel": "render_fwd_kernel (one launch blends the 8 frames of a step)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "issue_side_from_ncu": issue,
                    "algorithmic_bytes_per_launch": abytes, "avg_launch_ms": blend_ms,
                    "num_rendered_per_view": R_views,
                    "stage_ms_per_step": {k: v / n_calls for k, v in prof["fwd_ms"].items()},
                    "note": "issue/latency-bound kernel (256*R pair evaluations per frame against 15 MB of algorithmic "
                            "traffic); see DESIGN.md 4"}
        launches = K * arm.KERNELS_PER_STEP

    if rank != 0:
        return
    line = {
        "metric": "nvs_frames_per_sec_256x256_65k_gaussians", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": K, "warmup": W_, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "nvs256: 65536 f3d-like Gaussians (SH deg 1) per GPU, 8 orbit views, 256x256, forward "
                               "(BASELINE configs[1])", "frames_per_step": VIEWS, "num_rendered_per_step": R_step,
                   "l2": "flushed (256 MB write) between timed steps",
                   "api": ("value: gof_forward_batch (8 views per call, sync-free); e2e: HostFrameSink.render from pinned host buffers (H2D of "
                           "the Gaussians, render_views, rgb/depth/alpha stored to pinned host memory by the blend kernel)" if args.impl == "ours"
                           else "one Rasterizer::forward call per frame")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / K,
                "exchange": (getattr(arm, "exchange", None) if (world > 1 and args.impl == "ours") else None)},
        "gpu_launches": launches,
    }
    if args.impl == "ours":
        line["roofline"] = roofline
        line["per_view_api"] = {"value": per_view_value, "e2e": e2e_pv_value, "unit": "frames/s",
                                "what": "the same frames one call per frame: value through _C.rasterize_gaussians "
                                        "(blocking num_rendered hand-off per frame, like the reference), e2e through "
                                        "render_predicted_more_v2_gof"}
        if world == 1 and not args.no_cpu_baseline:
            import cases
            flat_cpu = [cases.case_to(c, "cpu") for c in flat]
            fps, cores, n, dt = cpu_oracle_frames_per_s(flat_cpu)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} frames (the {VIEWS} views of one step, cycled) in {dt:.1f}s "
                                              "(oracle/gof_oracle.c, OpenMP, all host cores)"}
    else:
        line["impl"] = "reference"
        line["gpu_launches"] = None
        line["cpu_baseline"] = {"value": value, "unit": "frames/s", "cores": 0, "kind": "reference",
                                "sample": f"{K} steps x {VIEWS} frames; the reference path has no CPU implementation: "
                                          "this is its unmodified CUDA source compiled for sm_100a "
                                          "(oracle/_ref/libgof_ref.so), run on the GPU"}
    print(json.dumps(line))


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
