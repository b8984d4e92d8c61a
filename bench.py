#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (BASELINE.json: "256x256 NVS frames/sec @65k Gaussians;
renderCUDA HBM GB/s vs roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload nvs256|train256|cycle3|batch512] [--no-others] [--no-cpu-baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs[1..4]; one seeded "f3d-like" scene of 65,536 Gaussians, SH degree 1, per unit):

  nvs256   (headline, configs[1]) the 8 orbit views of one scene at 256x256, forward.  A STEP = those 8 frames,
           issued as ONE batched call (gof_forward_batch).
  train256 (configs[2]) one training step = forward + backward of the same 8 views (rasterize_views_autograd:
           gof_forward_batch + gof_backward_batch, gradients summed over the views in-kernel).
  cycle3   (configs[3]) the cycle-aggregative 3-view loop (visualize.py:288-340): render 2 aggregation views of the
           source set, re-predict a Gaussian set from each (stand-in predictor, same torch code in both arms),
           concatenate to 196,608 Gaussians, render the 8 orbit views of the merged set.  A STEP = 10 frames.
  batch512 (configs[4], one GPU's share) 8 scenes x 8 views at 512x512.  A STEP = 64 frames.

Every workload reports `value` = frames/s over all GPUs with the inputs resident in HBM (CUDA events per step on the
launch stream, L2 flushed by a 256 MB write between steps, max over ranks), `e2e` = the same through the public API
from pinned HOST buffers with the H2D / D2H copies inside the timed region (host clock, device idle on both sides),
and a `roofline` for its dominant kernel (forward blend; backward blend for train256) from CUDA events recorded inside
the library on the launch stream.  The default run measures the headline workload and appends the other three as
`other_workloads`; `--workload X` makes X the headline.  The timed region never has fewer than MIN_TIMED_STEPS[workload] steps:
`steps` in the JSON line is the number of steps actually timed (`steps_requested` = --steps).

`--impl reference` times the UNMODIFIED reference rasterizer (oracle/_ref/libgof_ref.so: its CUDA sources compiled
for sm_100a -- the reference has no CPU implementation of this path) on the same workloads, one Rasterizer::forward /
::backward call per frame as its own loops do; if that library is absent it falls back to the CPU oracle port.  The
product arm never touches oracle/: only the `cpu_baseline` leg (rank 0, N=1) and the reference arm do.
"""
from __future__ import annotations

import argparse
import json
import math
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
L2_FLUSH_BYTES = 256 << 20
MIN_TIMED_STEPS = {"nvs256": 400, "train256": 100, "cycle3": 40, "batch512": 20}    # floor of timed steps per workload
OTHER_STEPS = {"train256": 100, "cycle3": 40, "batch512": 20}
WORKLOADS = ("nvs256", "train256", "cycle3", "batch512")
FOV = 13.164


# ------------------------------------------------------------------ distributed plumbing ----
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


def _reduce(x: float, world: int, op) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(x, world):
    import torch.distributed as dist
    return _reduce(x, world, dist.ReduceOp.MAX if world > 1 else None)


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)          # let the sampler start before the timed region does
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # "under load": the samples taken while the GPU drew more than idle power (the timed region), if any
        load = [s for s, p in zip(sm, pw) if p > (min(pw) + 0.25 * (max(pw) - min(pw)))] if pw else []
        return {"sm_mhz": statistics.median(load or sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_under_load": len(load), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------- formulas ----------
def tiles(W, H):
    return ((W + 15) // 16) * ((H + 15) // 16)


def bytes_render_fwd(P, R, W, H):
    """A_render_fwd = 8 T + 60 R + 12 P + 60 N per frame."""
    return 8 * tiles(W, H) + 60 * R + 12 * P + 60 * W * H


def bytes_render_bwd(P, R, W, H):
    """A_render_bwd = 8 T + 80 R + 60 N + 68 P per frame."""
    return 8 * tiles(W, H) + 80 * R + 60 * W * H + 68 * P


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(kernel: str):
    """(dram bytes per launch, issue-side figures) of `kernel` from the committed ncu --set full summary, if present."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            d = json.load(f).get(kernel, {})
        issue = {k: d[k] for k in ("issue_active_pct", "active_lanes_per_instruction", "warp_instructions") if d.get(k) is not None}
        return d.get("dram_bytes_per_launch"), (issue or None), d.get("workload")
    except Exception:
        return None, None, None


# ------------------------------------------------------------------------- scenes ------------
def make_cams(device):
    from f3d_gaus_b200 import cameras
    cams = cameras.orbit_cameras(VIEWS)
    return cams, cameras.Cameras(*[t.to(device) for t in cams])


def flat_cases(pc_cpu, cams, res, device):
    import cases
    return [cases.make_case(pc_cpu, cams.world_view[v], cams.full_proj[v], cams.centers[v], W=res, H=res, fov_deg=FOV,
                            device=device) for v in range(cams.world_view.shape[0])]


class RefFrames:
    """The unmodified reference rasterizer (CUDA, sm_100a build) through its own C++ entry points, one call per frame,
    with the per-call allocations + fills of its torch glue (rasterize_points.cu:72-73,161-170)."""

    def __init__(self, device):
        import refgpu
        self.run = refgpu.RefRun()
        self.lib = self.run.lib
        self.device = device

    def forward(self, c):
        P, W, H = c["means3D"].shape[0], c["W"], c["H"]
        out_color = torch.full((9, H, W), 0.0, dtype=torch.float32, device=self.device)
        radii = torch.full((P,), 0, dtype=torch.int32, device=self.device)
        p = lambda t: t.data_ptr()
        R = self.lib.ref_forward(self.run.state, P, c["D"], c["shs"].shape[1], p(c["bg"]), W, H, p(c["means3D"]),
                                 p(c["shs"]), None, p(c["opacities"]), p(c["scales"]), c["scale_modifier"],
                                 p(c["rotations"]), None, None, p(c["viewmatrix"]), p(c["projmatrix"]), p(c["campos"]),
                                 c["tanfovx"], c["tanfovy"], c["kernel_size"], None, 0, p(out_color), p(radii), 0)
        if R < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        self.last = (R, radii, P, c["shs"].shape[1], W, H)
        return R, out_color, radii

    def backward(self, c, dL):
        """Rasterizer::backward of the frame `forward` rendered last; the ten zero-filled gradient tensors of the glue."""
        R, radii, P, M, W, H = self.last
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        g = [z(P, 3), z(P, 2, 2), z(P, 1), z(P, 3), z(P, 3), z(P, 6), z(P, M, 3), z(P, 3), z(P, 4), z(P, 10)]
        p = lambda t: t.data_ptr()
        rc = self.lib.ref_backward(self.run.state, P, c["D"], M, R, p(c["bg"]), W, H, p(c["means3D"]), p(c["shs"]), None,
                                   None, p(c["scales"]), c["scale_modifier"], p(c["rotations"]), None, p(c["viewmatrix"]),
                                   p(c["projmatrix"]), p(c["campos"]), c["tanfovx"], c["tanfovy"], c["kernel_size"], None,
                                   p(radii), p(dL), *[p(t) for t in g], 0)
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return g


def case_of(pc, b, wv, fp, cc, bg, res, D=1):
    """Flat argument dict of one frame from a predictor-style dict (reference arm)."""
    tanfov = math.tan(FOV * math.pi / 360)
    shs = torch.cat([pc["features_dc"][b], pc["features_rest"][b]], dim=1).contiguous()
    return {"bg": bg, "means3D": pc["xyz"][b], "shs": shs, "opacities": pc["opacity"][b], "scales": pc["scaling"][b],
            "rotations": pc["rotation"][b], "scale_modifier": 1.0, "viewmatrix": wv.contiguous(), "projmatrix": fp.contiguous(),
            "campos": cc.contiguous(), "tanfovx": tanfov, "tanfovy": tanfov, "kernel_size": 0.0, "W": res, "H": res, "D": D}


def ref_render_fn(ref: RefFrames):
    """`render_views`-shaped function on the reference build (for cycle.cycle_aggregate in the reference arm)."""
    def fn(pc, b, wvs, fps, ccs, bg, cfg, workspace=None, epilogue=False, **_):
        res = int(cfg["model"]["training_resolution"])
        V = wvs.reshape(-1, 16).shape[0]
        wvs, fps, ccs = wvs.reshape(V, 4, 4), fps.reshape(V, 4, 4), ccs.reshape(V, 3)
        got = [ref.forward(case_of(pc, b, wvs[v], fps[v], ccs[v], bg, res)) for v in range(V)]
        fn.Rs = [g[0] for g in got]
        raster = torch.stack([g[1] for g in got])
        return {"raster": raster, "render": raster[:, 0:3], "rendered_depth": raster[:, 6:7], "rendered_alpha": raster[:, 7:8]}
    return fn


# ------------------------------------------------------------------------- workloads ---------
class Workload:
    """One BASELINE config on one arm.  step() = one resident step, e2e_step() = the same from host buffers;
    finish() after a synchronisation returns the num_rendered of the last step (None => a binning blob overflowed and
    was grown: run the step again)."""
    name = ""
    frames_per_step = VIEWS
    res = 256
    kernels_per_step = None
    dominant = "render_fwd_kernel"

    def __init__(self, impl, rank, world, device):
        self.impl, self.rank, self.world, self.device = impl, rank, world, device
        self.ours = impl == "ours"
        self.bg = torch.zeros(3, device=device)
        self.cams_cpu, self.cams = make_cams(device)
        self.cfg = None
        self.R = None
        self.exchange = None
        self.exchange_checked = None
        if not self.ours:
            self.ref = RefFrames(device)

    def finish(self):
        return self.R

    def e2e_finish(self):
        return True

    def roofline_launches(self):
        """[(P, R, W, H)] of the dominant kernel's frames in one step (for the algorithmic bytes)."""
        raise NotImplementedError


class Nvs256(Workload):
    name = "nvs256"
    description = ("nvs256: 65536 f3d-like Gaussians (SH deg 1) per GPU, 8 orbit views, 256x256, forward "
                   "(BASELINE configs[1])")
    kernels_per_step = 5        # preprocess, tile_scan, scatter, tile_sort_gather, render_fwd (one launch each per batch)

    def __init__(self, impl, rank, world, device):
        super().__init__(impl, rank, world, device)
        from f3d_gaus_b200 import synthetic
        from f3d_gaus_b200.staging import PinnedScene
        self.pc_cpu = synthetic.f3d_like(seed=rank, S=P_SIDE)
        self.cfg = synthetic.cfg_for(self.res)
        self.flat = flat_cases(self.pc_cpu, self.cams_cpu, self.res, device)
        self.host_pc = PinnedScene(self.pc_cpu, device)
        self.h2d = sum(v.numel() * v.element_size() for v in self.pc_cpu.values())
        self.d2h = VIEWS * 5 * self.res * self.res * 4
        if self.ours:
            from f3d_gaus_b200.diff_gof_rasterization import _C, BatchWorkspace, rasterize_views
            self._C, self.rasterize_views = _C, rasterize_views
            self.ws = BatchWorkspace(device)
            self.out = torch.empty((VIEWS, 9, self.res, self.res), dtype=torch.float32, device=device)
            self.sink = None
            self.peer = None
            self.empty = torch.Tensor([])
        else:
            self.out_dev = torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32, device=device)
            self.out_host = torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32).pin_memory()

    # resident
    def step(self):
        if self.ours:
            c, cm = self.flat[0], self.cams
            self.rasterize_views(c["bg"], c["means3D"], None, c["opacities"], c["scales"], c["rotations"], 1.0,
                                 cm.world_view, cm.full_proj, c["tanfovx"], c["tanfovy"], 0.0, self.res, self.res, c["shs"],
                                 c["D"], cm.centers, workspace=self.ws, out_color=self.out)
        else:
            self.R = [self.ref.forward(c)[0] for c in self.flat]

    def finish(self):
        return self.ws.finish() if self.ours else self.R

    def per_view_step(self):
        """The same 8 frames one call per frame through the reference-shaped `_C.rasterize_gaussians`."""
        e = self.empty
        for c in self.flat:
            self._C.rasterize_gaussians(c["bg"], c["means3D"], e, c["opacities"], c["scales"], c["rotations"], 1.0, e, e,
                                        c["viewmatrix"], c["projmatrix"], c["tanfovx"], c["tanfovy"], 0.0, e, self.res,
                                        self.res, c["shs"], c["D"], c["campos"], False, False)

    # end to end
    def e2e_step(self):
        dev_pc = self.host_pc.upload()                    # PinnedScene: one H2D copy for the whole set
        cm = self.cams
        if not self.ours:
            for v in range(VIEWS):
                _, color, _ = self.ref.forward(case_of(dev_pc, 0, cm.world_view[v], cm.full_proj[v], cm.centers[v], self.bg, self.res))
                self.out_dev[v, 0:3].copy_(color[0:3])
                self.out_dev[v, 3:4].copy_(color[6:7])
                self.out_dev[v, 4:5].copy_(color[7:8])
            self.out_host.copy_(self.out_dev, non_blocking=True)
            return
        from f3d_gaus_b200.gaussian_renderer import HostFrameSink
        if self.sink is None:
            # frames are stored into pinned host memory by the blend kernel itself (gof_set_frame_sink);
            # GOF_BENCH_READBACK=dma selects the packed DMA copy instead (tools/e2e_breakdown.py compares them)
            self.sink = HostFrameSink(VIEWS, self.res, self.res, self.device, chunks=1,
                                      zero_copy=os.environ.get("GOF_BENCH_READBACK", "kernel") != "dma")
        self.sink.render(dev_pc, 0, cm.world_view, cm.full_proj, cm.centers, self.bg, self.cfg)
        if self.world > 1:
            self._exchange()

    def _exchange(self):
        # the path's one exchange step: every rank receives all scenes' frames -- one kernel that packs
        # the consumed channels and stores them into every rank's buffer over NVLink peer memory
        # (sharding.PeerFrameGather); NCCL all_gather if symmetric memory cannot be set up
        from f3d_gaus_b200 import sharding
        if self.peer is None:
            try:
                self.peer = sharding.PeerFrameGather(self.world, VIEWS, self.res, self.res, self.device)
                self.exchange = "fused pack + all-gather over NVLink peer memory (gof_pack_gather, torch symmetric memory)"
            except Exception as ex:      # noqa: BLE001
                self.peer = False
                self.exchange = f"NCCL all_gather_into_tensor (peer memory unavailable: {type(ex).__name__})"
        r = self.sink.last_raster
        if self.peer:
            self.gathered = self.peer.push(r, first_scene=self.rank)
            if self.exchange_checked is None:
                # self-check, once, during warm-up: the fused gather must equal pack + NCCL all_gather bit for bit
                want = sharding.gather_frames(sharding.pack_frames(r[None, :, 0:3], r[None, :, 6:7], r[None, :, 7:8]), self.world)
                self.exchange_checked = bool(torch.equal(self.gathered, want))
        else:
            self.gathered = sharding.gather_frames(sharding.pack_frames(r[None, :, 0:3], r[None, :, 6:7], r[None, :, 7:8]), self.world)

    def e2e_finish(self):
        return self.sink.finish() if self.ours else True

    def e2e_pipelined(self, K, slots=2):
        """K scenes streamed through a `slots`-deep pipeline from pinned host buffers to pinned host frames
        (gaussian_renderer.SceneStreamer): H2D of scene k+1 | render k | frames of k-1 to host; the host only waits
        in collect().  The consumer is the host, so no GPU-to-GPU gather is issued.  Returns the host-clock ms for the
        K scenes (first submit to last collect)."""
        cm = self.cams
        if self.ours:
            from f3d_gaus_b200.gaussian_renderer import SceneStreamer
            if getattr(self, "streamer", None) is None:
                # frames reach the host as posted writes of the blend kernel on a GPU that has the host to itself, by DMA
                # from a packed device block when several GPUs share it (SceneStreamer docstring; GOF_BENCH_READBACK overrides)
                self.readback = os.environ.get("GOF_BENCH_READBACK", "kernel" if self.world == 1 else "dma")
                self.streamer = SceneStreamer(VIEWS, self.res, self.res, self.device, cm.world_view, cm.full_proj, cm.centers,
                                              self.bg, self.cfg, slots=slots, zero_copy=self.readback != "dma")
            st = self.streamer
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(K):
                if st.pending == st.slots:
                    st.collect()
                st.submit(self.host_pc)
            while st.pending:
                st.collect()
            return (time.perf_counter() - t0) * 1e3
        # reference arm: the same 2-slot host loop around its blocking per-frame calls
        self.readback = "dma"
        if getattr(self, "ref_slots", None) is None:
            self.ref_copy, self.ref_d2h = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
            self.ref_slots = [{"slab": torch.empty(self.host_pc.host_slab.numel(), dtype=torch.uint8, device=self.device),
                               "dev": torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32, device=self.device),
                               "host": torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32).pin_memory(),
                               "h2d": torch.cuda.Event(), "done": torch.cuda.Event()} for _ in range(slots)]
        main = torch.cuda.current_stream(self.device)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pend = []
        for i in range(K):
            sl = self.ref_slots[i % slots]
            if len(pend) == slots:
                pend.pop(0)["done"].synchronize()
            with torch.cuda.stream(self.ref_copy):
                dev_pc = self.host_pc.upload(sl["slab"])
                sl["h2d"].record(self.ref_copy)
            main.wait_event(sl["h2d"])
            for v in range(VIEWS):
                _, color, _ = self.ref.forward(case_of(dev_pc, 0, cm.world_view[v], cm.full_proj[v], cm.centers[v], self.bg, self.res))
                sl["dev"][v, 0:3].copy_(color[0:3])
                sl["dev"][v, 3:4].copy_(color[6:7])
                sl["dev"][v, 4:5].copy_(color[7:8])
            ev = torch.cuda.Event()
            ev.record(main)
            self.ref_d2h.wait_event(ev)
            with torch.cuda.stream(self.ref_d2h):
                sl["host"].copy_(sl["dev"], non_blocking=True)
                sl["done"].record(self.ref_d2h)
            pend.append(sl)
        for sl in pend:
            sl["done"].synchronize()
        return (time.perf_counter() - t0) * 1e3

    def e2e_per_view_step(self):
        """The same through the reference's own one-frame function (render_predicted_more_v2_gof)."""
        from f3d_gaus_b200.gaussian_renderer import render_predicted_more_v2_gof
        if not hasattr(self, "pv_dev"):
            self.pv_dev = torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32, device=self.device)
            self.pv_host = torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32).pin_memory()
        dev_pc = self.host_pc.upload()
        cm = self.cams
        with torch.no_grad():
            for v in range(VIEWS):
                o = render_predicted_more_v2_gof(dev_pc, 0, cm.world_view[v:v + 1], cm.full_proj[v:v + 1],
                                                 cm.centers[v:v + 1], self.bg, self.cfg)
                self.pv_dev[v, 0:3].copy_(o["render"])
                self.pv_dev[v, 3:4].copy_(o["rendered_depth"])
                self.pv_dev[v, 4:5].copy_(o["rendered_alpha"])
        self.pv_host.copy_(self.pv_dev, non_blocking=True)

    def roofline_launches(self):
        return [(P_SIDE * P_SIDE, r, self.res, self.res) for r in self.finish()]


class Train256(Workload):
    name = "train256"
    description = ("train256: one training step = forward + backward of the 8 orbit views of 65536 f3d-like Gaussians at "
                   "256x256, seeded dL/dout (BASELINE configs[2])")
    kernels_per_step = 5 + 3     # forward batch + gacc clear (memset), render_bwd, preprocess_bwd
    dominant = "render_bwd_kernel"

    def __init__(self, impl, rank, world, device):
        super().__init__(impl, rank, world, device)
        from f3d_gaus_b200 import synthetic
        from f3d_gaus_b200.staging import PinnedScene
        self.pc_cpu = synthetic.f3d_like(seed=rank, S=P_SIDE)
        self.cfg = synthetic.cfg_for(self.res)
        self.flat = flat_cases(self.pc_cpu, self.cams_cpu, self.res, device)
        self.dL = torch.randn(VIEWS, 9, self.res, self.res, generator=torch.Generator().manual_seed(3)).to(device)
        self.tanfov = math.tan(FOV * math.pi / 360)
        self.host_pc = PinnedScene(self.pc_cpu, device)
        self.target_host = torch.rand(VIEWS, 3, self.res, self.res, generator=torch.Generator().manual_seed(4)).pin_memory()
        self.target_dev = torch.empty_like(self.target_host, device=device)
        self.h2d = sum(v.numel() * 4 for v in self.pc_cpu.values()) + self.target_host.numel() * 4
        self.d2h = 4
        if self.ours:
            from f3d_gaus_b200.diff_gof_rasterization import rasterize_views_autograd
            self.rva = rasterize_views_autograd
            c = self.flat[0]
            self.leaves = [c[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")]

    def _ours_fwd(self, leaves):
        xyz, op, sc, rot, shs = leaves
        m2d = torch.zeros_like(xyz, requires_grad=True)
        cm = self.cams
        color, radii = self.rva(xyz, m2d, op, shs=shs, scales=sc, rotations=rot, bg=self.bg, viewmatrices=cm.world_view,
                                projmatrices=cm.full_proj, campos=cm.centers, tanfovx=self.tanfov, tanfovy=self.tanfov,
                                image_height=self.res, image_width=self.res, sh_degree=1)
        return color

    def step(self):
        if self.ours:
            for t in self.leaves:
                t.grad = None
            color = self._ours_fwd(self.leaves)
            color.backward(self.dL)
            self.R = "device"
        else:
            acc, Rs = None, []
            for v, c in enumerate(self.flat):
                Rs.append(self.ref.forward(c)[0])
                g = self.ref.backward(c, self.dL[v])
                acc = g if acc is None else [a + b for a, b in zip(acc, g)]      # the V-way sum autograd performs
            self.R = Rs

    def finish(self):
        if self.ours:
            # num_rendered of the 8 views: one plain forward (outside any timed region), cached
            if getattr(self, "_R", None) is None:
                from f3d_gaus_b200.diff_gof_rasterization import rasterize_views
                c, cm = self.flat[0], self.cams
                self._R = rasterize_views(c["bg"], c["means3D"], None, c["opacities"], c["scales"], c["rotations"], 1.0,
                                          cm.world_view, cm.full_proj, c["tanfovx"], c["tanfovy"], 0.0, self.res, self.res,
                                          c["shs"], c["D"], cm.centers)[0]
            return self._R
        return self.R

    def e2e_step(self):
        """Host buffers in (the Gaussian set + 8 target images), L1 photometric loss on the device, backward, the loss
        value read back."""
        dev_pc = self.host_pc.upload()
        self.target_dev.copy_(self.target_host, non_blocking=True)
        shs = torch.cat([dev_pc["features_dc"][0], dev_pc["features_rest"][0]], dim=1)
        if self.ours:
            leaves = [t.detach().requires_grad_(True) for t in (dev_pc["xyz"][0], dev_pc["opacity"][0], dev_pc["scaling"][0],
                                                               dev_pc["rotation"][0], shs)]
            color = self._ours_fwd(leaves)
            loss = (color[:, 0:3] - self.target_dev).abs().mean()
            loss.backward()
        else:
            cm = self.cams
            acc, loss = None, 0.0
            n = self.target_dev.numel()
            for v in range(VIEWS):
                c = case_of(dev_pc, 0, cm.world_view[v], cm.full_proj[v], cm.centers[v], self.bg, self.res)
                _, color, _ = self.ref.forward(c)
                diff = color[0:3] - self.target_dev[v]
                loss = loss + diff.abs().sum() / n
                dL = torch.zeros_like(color)
                dL[0:3] = torch.sign(diff) / n
                g = self.ref.backward(c, dL)
                acc = g if acc is None else [a + b for a, b in zip(acc, g)]
        self.loss = float(loss.item())                 # D2H of the step's result

    def roofline_launches(self):
        return [(P_SIDE * P_SIDE, r, self.res, self.res) for r in self.finish()]


class Cycle3(Workload):
    name = "cycle3"
    frames_per_step = 2 + VIEWS
    description = ("cycle3: cycle-aggregative 3-view loop at 256x256 (visualize.py:288-340): 2 aggregation renders of the "
                   "65536-Gaussian source set, stand-in re-prediction, concat to 196608 Gaussians, 8 orbit views of the "
                   "merged set; 10 frames per step (BASELINE configs[3])")
    kernels_per_step = 10       # two batched forward passes

    def __init__(self, impl, rank, world, device):
        super().__init__(impl, rank, world, device)
        from f3d_gaus_b200 import cameras, cycle, synthetic
        from f3d_gaus_b200.staging import PinnedScene
        self.cycle = cycle
        self.pc_cpu = synthetic.f3d_like(seed=rank, S=P_SIDE)
        self.cfg = synthetic.cfg_for(self.res)
        self.src = {k: v.to(device) for k, v in self.pc_cpu.items()}
        pick = [2, 5]                                        # the two aggregation views
        self.agg = cameras.Cameras(*[t[pick].contiguous() for t in self.cams])
        self.predict = cycle.unproject_predictor(self.cfg)
        self.host_pc = PinnedScene(self.pc_cpu, device)
        self.h2d = sum(v.numel() * 4 for v in self.pc_cpu.values())
        self.d2h = VIEWS * 5 * self.res * self.res * 4
        self.merged_P = 3 * P_SIDE * P_SIDE
        if self.ours:
            from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace
            from f3d_gaus_b200.gaussian_renderer import HostFrameSink, render_views
            self.render_views = render_views
            self.ws_a, self.ws_b = BatchWorkspace(device), BatchWorkspace(device)
            self.out = torch.empty((VIEWS, 9, self.res, self.res), dtype=torch.float32, device=device)
            self.sink = HostFrameSink(VIEWS, self.res, self.res, device, chunks=1, zero_copy=True)
            self.render_fn = None
        else:
            self.render_fn = ref_render_fn(self.ref)
            self.out_dev = torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32, device=device)
            self.out_host = torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32).pin_memory()

    def _loop(self, src):
        # ours: nothing synchronises inside the loop; finish() reads both workspaces' overflow flags after the step
        merged, frames = self.cycle.cycle_aggregate(src, self.predict, self.agg, self.cfg, self.bg,
                                                    workspace=self.ws_a if self.ours else None, render_fn=self.render_fn,
                                                    check_overflow=False)
        return merged

    def step(self):
        merged = self._loop(self.src)
        cm = self.cams
        if self.ours:
            self.render_views(merged, 0, cm.world_view, cm.full_proj, cm.centers, self.bg, self.cfg, workspace=self.ws_b,
                              epilogue=False, out_color=self.out)
        else:
            self.R = [self.ref.forward(case_of(merged, 0, cm.world_view[v], cm.full_proj[v], cm.centers[v], self.bg, self.res))[0]
                      for v in range(VIEWS)]

    def finish(self):
        if self.ours:
            a, b = self.ws_a.finish(), self.ws_b.finish()
            return None if (a is None or b is None) else (a, b)
        return (self.render_fn.Rs, self.R)

    def e2e_step(self):
        dev_pc = self.host_pc.upload()
        merged = self._loop(dev_pc)
        cm = self.cams
        if self.ours:
            self.sink.render(merged, 0, cm.world_view, cm.full_proj, cm.centers, self.bg, self.cfg)
        else:
            for v in range(VIEWS):
                _, color, _ = self.ref.forward(case_of(merged, 0, cm.world_view[v], cm.full_proj[v], cm.centers[v], self.bg, self.res))
                self.out_dev[v, 0:3].copy_(color[0:3])
                self.out_dev[v, 3:4].copy_(color[6:7])
                self.out_dev[v, 4:5].copy_(color[7:8])
            self.out_host.copy_(self.out_dev, non_blocking=True)

    def e2e_finish(self):
        if not self.ours:
            return True
        got = [self.ws_a.finish(), self.sink.finish()]             # visit both: an overflowed workspace grows in finish()
        return None if any(g is None for g in got) else got

    def roofline_launches(self):
        a, b = self.finish()
        return [(P_SIDE * P_SIDE, r, self.res, self.res) for r in a] + [(self.merged_P, r, self.res, self.res) for r in b]


class Batch512(Workload):
    name = "batch512"
    res = 512
    SCENES = 8
    frames_per_step = 8 * VIEWS
    description = ("batch512: 8 scenes x 8 orbit views at 512x512, 65536 f3d-like Gaussians per scene, forward; one GPU's "
                   "share of the 64-scene batch (BASELINE configs[4])")
    kernels_per_step = 5 * 8

    def __init__(self, impl, rank, world, device):
        super().__init__(impl, rank, world, device)
        from f3d_gaus_b200 import synthetic
        from f3d_gaus_b200.staging import PinnedScene
        self.cfg = synthetic.cfg_for(self.res)
        seeds = [rank * self.SCENES + s for s in range(self.SCENES)]
        self.pcs_cpu = [synthetic.f3d_like(seed=s, S=P_SIDE) for s in seeds]
        self.host_pcs = [PinnedScene(pc, device) for pc in self.pcs_cpu]
        self.scenes = [{k: v.to(device) for k, v in pc.items()} for pc in self.pcs_cpu]
        self.h2d = self.SCENES * sum(v.numel() * 4 for v in self.pcs_cpu[0].values())
        self.d2h = self.SCENES * VIEWS * 5 * self.res * self.res * 4
        if self.ours:
            from f3d_gaus_b200.diff_gof_rasterization import BatchWorkspace
            from f3d_gaus_b200.gaussian_renderer import HostFrameSink, render_views
            self.render_views = render_views
            self.wss = [BatchWorkspace(device) for _ in range(self.SCENES)]
            self.out = torch.empty((VIEWS, 9, self.res, self.res), dtype=torch.float32, device=device)
            self.sinks = [HostFrameSink(VIEWS, self.res, self.res, device, chunks=1, zero_copy=True) for _ in range(self.SCENES)]
        else:
            self.out_dev = torch.empty((VIEWS, 5, self.res, self.res), dtype=torch.float32, device=device)
            self.out_host = torch.empty((self.SCENES, VIEWS, 5, self.res, self.res), dtype=torch.float32).pin_memory()

    def step(self):
        cm = self.cams
        if self.ours:
            for pc, ws in zip(self.scenes, self.wss):
                self.render_views(pc, 0, cm.world_view, cm.full_proj, cm.centers, self.bg, self.cfg, workspace=ws,
                                  epilogue=False, out_color=self.out)
        else:
            self.R = [[self.ref.forward(case_of(pc, 0, cm.world_view[v], cm.full_proj[v], cm.centers[v], self.bg, self.res))[0]
                       for v in range(VIEWS)] for pc in self.scenes]

    def finish(self):
        if self.ours:
            Rs = [ws.finish() for ws in self.wss]
            return None if any(r is None for r in Rs) else Rs
        return self.R

    def e2e_step(self):
        cm = self.cams
        for s in range(self.SCENES):
            dev_pc = self.host_pcs[s].upload()
            if self.ours:
                self.sinks[s].render(dev_pc, 0, cm.world_view, cm.full_proj, cm.centers, self.bg, self.cfg)
            else:
                for v in range(VIEWS):
                    _, color, _ = self.ref.forward(case_of(dev_pc, 0, cm.world_view[v], cm.full_proj[v], cm.centers[v], self.bg, self.res))
                    self.out_dev[v, 0:3].copy_(color[0:3])
                    self.out_dev[v, 3:4].copy_(color[6:7])
                    self.out_dev[v, 4:5].copy_(color[7:8])
                self.out_host[s].copy_(self.out_dev, non_blocking=True)

    def e2e_finish(self):
        if not self.ours:
            return True
        got = [sk.finish() for sk in self.sinks]                  # visit all: an overflowed workspace grows in finish()
        return None if any(g is None for g in got) else got

    def roofline_launches(self):
        return [(P_SIDE * P_SIDE, r, self.res, self.res) for Rs in self.finish() for r in Rs]


WORKLOAD_CLASSES = {"nvs256": Nvs256, "train256": Train256, "cycle3": Cycle3, "batch512": Batch512}


# ------------------------------------------------------------------------- measurement -------
class Timer:
    def __init__(self, world, device, warmup):
        self.world, self.W = world, warmup
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)

    def device(self, step_fn, K):
        """K steps, CUDA events around each on the launch stream, L2 flush outside the events; max over ranks (ms)."""
        for _ in range(self.W):
            step_fn()
        torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        barrier(self.world)
        for i in range(K):
            self.flush.zero_()                      # L2 flush between timed iterations, outside the events
            starts[i].record()
            step_fn()
            ends[i].record()
        barrier(self.world)
        return max_over_ranks(sum(s.elapsed_time(e) for s, e in zip(starts, ends)), self.world)

    def e2e(self, step_fn, K):
        """Host clock around each step (host buffers in, host buffers out, device idle on both sides)."""
        for _ in range(self.W):
            step_fn()
        torch.cuda.synchronize()
        barrier(self.world)
        ms = 0.0
        for i in range(K):
            self.flush.zero_()
            barrier(self.world)          # ranks start each step together: the exchange step would otherwise time their skew
            t0 = time.perf_counter()
            step_fn()
            torch.cuda.synchronize()
            ms += (time.perf_counter() - t0) * 1e3
        barrier(self.world)
        return max_over_ranks(ms, self.world)


def settle(wl, fn, finish):
    """First calls size the binning blobs; an overflow grows them and the step is re-run."""
    for _ in range(3):
        fn()
        torch.cuda.synchronize()
        if finish() is not None:
            return
    raise RuntimeError(f"{wl.name}: binning blob still overflowing after 3 attempts")


def measure(wl: Workload, timer: Timer, K: int, local: int, sample_clocks: bool):
    """Resident + end-to-end measurement of one workload.  Returns the dict that becomes (part of) the JSON line."""
    world, ours = wl.world, wl.ours
    if ours:                                  # (the reference arm never loads the product's library)
        from f3d_gaus_b200 import _lib
    settle(wl, wl.step, wl.finish)
    if ours:
        _lib.profile_enable(wl.device.index, True)
        _lib.profile_read(wl.device.index)
    sampler = ClockSampler(local) if sample_clocks else None
    total_ms = timer.device(wl.step, K)
    clocks = sampler.stop() if sampler else None
    prof = _lib.profile_read(wl.device.index) if ours else None
    if ours:
        _lib.profile_enable(wl.device.index, False)
    Rs = wl.finish()
    assert Rs is not None, "binning blob overflowed inside the timed region"
    frames = wl.frames_per_step * K * world
    out = {"workload": wl.name, "value": frames / (total_ms * 1e-3), "unit": "frames/s", "steps": K,
           "ms_per_step": total_ms / K, "frames_per_step": wl.frames_per_step}

    settle(wl, wl.e2e_step, wl.e2e_finish)
    e2e_ms = timer.e2e(wl.e2e_step, K)
    assert wl.e2e_finish() is not None, "binning blob overflowed inside the e2e region"
    out["e2e"] = {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": wl.h2d,
                  "d2h_bytes_per_step": wl.d2h, "ms_per_step": e2e_ms / K}
    if world > 1 and ours and wl.exchange:
        out["e2e"]["exchange"] = wl.exchange
        out["e2e"]["exchange_checked"] = wl.exchange_checked
    if hasattr(wl, "e2e_pipelined"):
        wl.e2e_pipelined(max(4, timer.W))                      # warm-up (sizes the per-slot workspaces)
        barrier(world)
        pipe_ms = max_over_ranks(wl.e2e_pipelined(K), world)
        barrier(world)
        # The metric is a THROUGHPUT (frames/s), so the end-to-end headline is the streaming loop; the step-at-a-time
        # figure (device idle on both sides of every step, plus the GPU-to-GPU gather at N > 1) stays as e2e_serial.
        out["e2e_serial"] = out["e2e"]
        out["e2e"] = {"value": frames / (pipe_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": wl.h2d,
                      "d2h_bytes_per_step": wl.d2h, "ms_per_step": pipe_ms / K, "mode": "pipelined", "slots": 2,
                      "readback": getattr(wl, "readback", None),
                      "what": "K scenes streamed from pinned host buffers to pinned host frames, 2 in flight: H2D of scene "
                              "k+1 (copy stream) | render k | frames of k-1 to host; host clock over the whole K-step region, "
                              "every step's copies inside it (no artificial L2 flush: each scene's inputs arrive fresh from the host); "
                              "the host is the consumer, so no GPU-to-GPU gather"}

    # roofline of the dominant kernel
    launches = wl.roofline_launches()
    out["num_rendered_per_step"] = int(sum(r for _, r, _, _ in launches))
    if ours and prof:
        peak, peak_src = measured_peak_gbs()
        bwd = wl.dominant == "render_bwd_kernel"
        calls = prof["bwd_calls"] if bwd else prof["fwd_calls"]
        ms = (prof["bwd_ms"]["blend_backward"] if bwd else prof["fwd_ms"]["blend"])
        if calls:
            steps_profiled = K + timer.W                          # the library's events also cover the warm-up steps
            per_step_ms = ms / steps_profiled                     # all launches of the kernel in one step
            n_launch = calls / steps_profiled
            abytes = sum((bytes_render_bwd if bwd else bytes_render_fwd)(*l) for l in launches)
            achieved = abytes / (per_step_ms * 1e-3) / 1e9
            traffic, issue, traffic_wl = traffic_from_profiles(wl.dominant)
            out["roofline"] = {
                "bound": "hbm", "kernel": f"{wl.dominant} ({n_launch:g} launch(es) per step)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic if traffic_wl in (None, wl.name) else None, "peak_source": peak_src,
                "issue_side_from_ncu": issue, "algorithmic_bytes_per_step": abytes, "kernel_ms_per_step": per_step_ms,
                "avg_launch_ms": per_step_ms / n_launch,
                "stage_ms_per_step": {**{k: v / steps_profiled for k, v in prof["fwd_ms"].items()},
                                      **({k: v / steps_profiled for k, v in prof["bwd_ms"].items()} if prof["bwd_calls"] else {})},
                "note": "issue/latency-bound kernel (256*R pair evaluations per frame against ~15 MB of algorithmic "
                        "traffic at 256x256); see DESIGN.md 4"}
        out["gpu_launches"] = K * wl.kernels_per_step
    return out, clocks, Rs


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
    flat = [cases.make_case(pc_cpu, cams.world_view[v], cams.full_proj[v], cams.centers[v], W=256, H=256, fov_deg=FOV)
            for v in range(VIEWS)]
    fps, cores, n, dt = cpu_oracle_frames_per_s(flat, budget_s=min(20.0, 1.0 * max(1, args.steps)))
    line = {"impl": "reference", "metric": "nvs_frames_per_sec_256x256_65k_gaussians", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * VIEWS / fps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": Nvs256.description},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n} frames in {dt:.1f}s (oracle/gof_oracle.c, OpenMP)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


METRIC = {"nvs256": "nvs_frames_per_sec_256x256_65k_gaussians",
          "train256": "train_frames_per_sec_256x256_65k_gaussians_fwd_bwd",
          "cycle3": "cycle_aggregative_frames_per_sec_256x256_196k_gaussians",
          "batch512": "nvs_frames_per_sec_512x512_65k_gaussians_8_scenes"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nvs256", choices=WORKLOADS)
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads block of the default run")
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
    timer = Timer(world, device, args.warmup)
    ours = args.impl == "ours"

    # ---------------- the headline workload ---------------------------------------------------------
    wl = WORKLOAD_CLASSES[args.workload](args.impl, rank, world, device)
    K = max(args.steps, MIN_TIMED_STEPS[args.workload]) if ours else args.steps
    head, clocks, Rs = measure(wl, timer, K, local, sample_clocks=(rank == 0))
    per_view = None
    if ours and args.workload == "nvs256":
        frames = VIEWS * K * world
        pv_ms = timer.device(wl.per_view_step, K)
        e2e_pv_ms = timer.e2e(wl.e2e_per_view_step, K)
        per_view = {"value": frames / (pv_ms * 1e-3), "e2e": frames / (e2e_pv_ms * 1e-3), "unit": "frames/s",
                    "what": "the same frames one call per frame: value through _C.rasterize_gaussians (blocking "
                            "num_rendered hand-off per frame, like the reference), e2e through render_predicted_more_v2_gof"}
    flat_for_cpu = wl.flat if args.workload in ("nvs256", "train256") else None
    del wl
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE configs, same arm, same run -----------------------------
    others = {}
    if args.workload == "nvs256" and not args.no_others:
        for name in ("train256", "cycle3", "batch512"):
            try:
                w = WORKLOAD_CLASSES[name](args.impl, rank, world, device)
                k = OTHER_STEPS[name] if ours else max(3, OTHER_STEPS[name] // 5)
                res, _, _ = measure(w, timer, k, local, sample_clocks=False)
                res["config"] = w.description
                others[name] = res
                del w
                torch.cuda.empty_cache()
            except Exception as ex:      # noqa: BLE001  -- a failing side workload must not cost the headline line
                others[name] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank != 0:
        return
    api = {"nvs256": "value: gof_forward_batch (8 views per call, sync-free); e2e: HostFrameSink.render from pinned host "
                     "buffers (H2D of the Gaussians, render_views, rgb/depth/alpha stored to pinned host memory by the blend kernel)",
           "train256": "value: rasterize_views_autograd + backward(dL) (gof_forward_batch + gof_backward_batch); e2e: H2D of the "
                       "set and 8 target images, L1 loss, backward, loss read back",
           "cycle3": "cycle.cycle_aggregate + render_views of the merged set; e2e: H2D of the source set, frames to pinned host memory",
           "batch512": "render_views per scene (8 views per call, sync-free); e2e: H2D of every scene, HostFrameSink per scene"}
    line = {
        "metric": METRIC[args.workload], "value": head["value"], "unit": "frames/s", "n_gpus": world,
        "steps": head["steps"], "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_CLASSES[args.workload].description, "frames_per_step": head["frames_per_step"],
                   "num_rendered_per_step": head["num_rendered_per_step"],
                   "l2": "flushed (256 MB write) between timed steps"},
        # how each arm is driven lives OUTSIDE `config`, so that the two arms' config blocks compare equal
        "api": api[args.workload] if ours else "one Rasterizer::forward (/ ::backward) call per frame",
        "clocks": clocks,
        "e2e": head["e2e"],
        "gpu_launches": head.get("gpu_launches"),
    }
    if "e2e_serial" in head:
        line["e2e_serial"] = head["e2e_serial"]
    if ours:
        line["roofline"] = head.get("roofline")
        if per_view:
            line["per_view_api"] = per_view
        if world == 1 and not args.no_cpu_baseline and flat_for_cpu is not None:
            import cases
            flat_cpu = [cases.case_to(c, "cpu") for c in flat_for_cpu]
            fps, cores, n, dt = cpu_oracle_frames_per_s(flat_cpu)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} forward frames (the {VIEWS} views of one step, cycled) in {dt:.1f}s "
                                              "(oracle/gof_oracle.c, OpenMP, all host cores)"}
    else:
        line["impl"] = "reference"
        line["gpu_launches"] = None
        line["cpu_baseline"] = {"value": head["value"], "unit": "frames/s", "cores": 0, "kind": "reference",
                                "sample": f"{head['steps']} steps x {head['frames_per_step']} frames; the reference path has no CPU "
                                          "implementation: this is its unmodified CUDA source compiled for sm_100a "
                                          "(oracle/_ref/libgof_ref.so), run on the GPU"}
    if others:
        line["other_workloads"] = others
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
