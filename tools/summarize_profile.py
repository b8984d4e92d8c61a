#!/usr/bin/env python
"""Turn one GPU-box visit (gpurun_out/<tag>/, written by tools/gpu_round.sh) into the tracked
summaries under profiles/:  <round>_launches.md (per-kernel share of a bench step from the ncu launch
list), <round>_blend_ncu.md (key `ncu --set full` metrics of the blend kernels) and
kernel_traffic.json (dram bytes per launch + issue-side figures of the blend kernels, read by bench.py).

    python tools/summarize_profile.py gpurun_out/s3a r01
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "lts__t_sector_hit_rate.pct",
]


def launches(src, out_md, title):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"\(.*", "", r[ki])
        k = re.sub(r"<unnamed>::", "", k)[:100]
        k = (k, r[gi], r[bi])                      # same kernel at different grids = different work
        a = agg.setdefault(k, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(src, ROOT)}` (ncu --metrics gpu__time_duration.sum "
                "--clock-control none; cold-cache, serialised launches: compare SHARES, not absolutes).\n\n")
        f.write("| launches | total us | share | avg us | grid | block | kernel |\n|---:|---:|---:|---:|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0] / 1e3:.2f} | {a[2]} | {a[3]} "
                    f"| `{k[0]}` |\n")
        f.write(f"\nTotal device time in the list: {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches.\n")
    return agg


def full(rep, out_md, title, traffic_json=None, traffic_kernel="render_fwd", traffic_workload="nvs256", traffic_what=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {m: hdr.index(m) for m in METRICS if m in hdr}
    ki = hdr.index("Kernel Name")
    names = [re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("unnamed>::", "").strip() for r in rows[2:]]
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(rep, ROOT)}` (`ncu --set full --clock-control none "
                "--import-source on`), read with `ncu -i ... --page raw --csv`.\n\n")
        f.write("| metric | unit | " + " | ".join(f"{n} #{i}" for i, n in enumerate(names)) + " |\n")
        f.write("|---|---|" + "---:|" * len(names) + "\n")
        for m, i in idx.items():
            f.write(f"| `{m}` | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")
    if traffic_json:
        def to_bytes(r, i):
            v = float(r[i].replace(",", ""))
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[units[i].lower()]
        vals = [to_bytes(r, hdr.index("dram__bytes_read.sum")) + to_bytes(r, hdr.index("dram__bytes_write.sum"))
                for r, n in zip(rows[2:], names) if traffic_kernel in n]
        if vals:
            def metric(name):
                sel = [float(r[hdr.index(name)].replace(",", "")) for r, n in zip(rows[2:], names) if traffic_kernel in n]
                return sum(sel) / len(sel) if (name in hdr and sel) else None
            try:
                with open(traffic_json) as f:
                    table = json.load(f)
            except Exception:
                table = {}
            table[traffic_kernel + "_kernel"] = {
                "dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals), "source": os.path.relpath(rep, ROOT),
                "issue_active_pct": metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "active_lanes_per_instruction": metric("smsp__thread_inst_executed_per_inst_executed.ratio"),
                "warp_instructions": metric("smsp__inst_executed.sum"), "workload": traffic_workload,
                "what": traffic_what}
            with open(traffic_json, "w") as f:
                json.dump(table, f, indent=1)


def main():
    src, tag = os.path.abspath(sys.argv[1]), sys.argv[2]
    prof = os.path.join(ROOT, "profiles")
    os.makedirs(prof, exist_ok=True)
    if os.path.exists(os.path.join(src, "launches.csv")):
        launches(os.path.join(src, "launches.csv"), os.path.join(prof, f"{tag}_launches.md"),
                 f"{tag}: kernel launch list of `bench.py --steps 2 --warmup 3` (N=1)")
    rep = os.path.join(src, "prof_blend.ncu-rep")
    if os.path.exists(rep):
        full(rep, os.path.join(prof, f"{tag}_blend_ncu.md"),
             f"{tag}: ncu --set full, forward pipeline of one bench step (8 views of 65,536 Gaussians, 256x256)",
             os.path.join(prof, "kernel_traffic.json"), "render_fwd", "nvs256",
             "bench.py nvs256 step: one launch blends 8 views of 65536 f3d-like Gaussians, 256x256")
    rep = os.path.join(src, "prof_bwd.ncu-rep")
    if os.path.exists(rep):
        full(rep, os.path.join(prof, f"{tag}_bwd_ncu.md"),
             f"{tag}: ncu --set full, backward blend of one train256 step (8 views of 65,536 Gaussians, 256x256)",
             os.path.join(prof, "kernel_traffic.json"), "render_bwd", "train256",
             "bench.py train256 step: one launch walks back the 8 views of 65536 f3d-like Gaussians, 256x256")
    rep = os.path.join(src, "prof_split.ncu-rep")
    if os.path.exists(rep):
        full(rep, os.path.join(prof, f"{tag}_split_ncu.md"),
             f"{tag}: ncu --set full, one-frame launches (65,536 Gaussians, one 256x256 orbit view): render_fwd_split_kernel and the stages in front of it")
    rep = os.path.join(src, "prof_head.ncu-rep")
    if os.path.exists(rep):
        full(rep, os.path.join(prof, f"{tag}_head_ncu.md"), f"{tag}: ncu --set full, predictor output head (64 images of 256x256, 23 channels)")
    for name in ("bench_ours.json", "bench_reference.json", "quick_bench.log", "head_bench.json", "train_bench.json",
                 "single_frame.log", "per_view_timeline.log"):
        p = os.path.join(src, name)
        if os.path.exists(p):
            with open(p) as f, open(os.path.join(prof, f"{tag}_{name}"), "w") as g:
                g.write(f.read())


if __name__ == "__main__":
    main()
