#!/usr/bin/env python
"""SASS evidence for the ring kernels: per kernel, the opcode histogram and every line that shows a Blackwell /
Hopper-class feature (TMA bulk copy UBLKCP, mbarrier SYNCS.*, async-proxy FENCE, vector REDG, MUFU / FP64 pipe use),
from `cuobjdump -sass` of the built objects.  Writes profiles/<tag>_sass_<kernel>.txt.

    python tools/sass_evidence.py r02
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "f3d_gaus_b200", "csrc", "build")
KERNELS = [("render_fwd.o", "render_fwd_kernelILb0ELb0ELb0E", "render_fwd"), ("render_fwd.o", "render_fwd_split_kernelILb0E", "render_fwd_split"), ("render_fwd.o", "render_fwd_kernelILb0ELb1ELb1E", "render_fwd_sink_mask"),
           ("render_bwd.o", "render_bwd_kernel", "render_bwd"), ("integrate.o", "integrate_pixels_kernel", "integrate"),
           ("binning.o", "tile_sort_gather_kernel", "tile_sort_gather")]
MARK = re.compile(r"UBLKCP|SYNCS|FENCE|REDG|ATOMG|ATOMS|MEMBAR|VOTE|MUFU|UTMA|LDGSTS|griddep|ACQBULK|CCTL")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
for obj, key, name in KERNELS:
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    sel = [b for b in blocks[1:] if key in b.split("\n", 1)[0]]
    if not sel:
        print("not found:", key); continue
    body = sel[0]
    fn = body.split("\n", 1)[0].strip()
    lines = []
    for ln in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            lines.append((m.group(1), m.group(2).strip()))
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in lines)
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_{name}.txt")
    with open(path, "w") as f:
        f.write(f"# cuobjdump -sass f3d_gaus_b200/csrc/build/{obj}  (sm_100a)\n# {fn}\n# {len(lines)} instructions\n\n")
        f.write("## opcode histogram (static)\n")
        for op, n in hist.most_common():
            f.write(f"{n:6d}  {op}\n")
        f.write("\n## lines showing TMA bulk copies (UBLKCP), mbarrier ops (SYNCS.*), proxy fences, L2 reductions (REDG), votes, MUFU\n")
        for a, t in lines:
            if MARK.search(t):
                f.write(f"/*{a}*/  {t} ;\n")
    print("wrote", os.path.relpath(path, ROOT), len(lines), "instructions")
