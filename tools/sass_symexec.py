#!/usr/bin/env python
"""Symbolic executor over the reference build's SASS: prints, as nested fma/mul/add expressions, what the
reference's sm_100a build of integrateCUDA (forward.cu:803-1218) actually evaluates for AA = r^T Sigma r and
BB = 2 b.r of each of its five rays.  nvcc shares products between rays (rays 1/3 share rx, 1/2 and 3/4 share
ry), so the fused/unfused pattern differs per ray; the GOF ray minimum is ill-conditioned enough at F3D-Gaus
scales that these roundings decide alpha to percents.  f3d_gaus_b200/csrc/integrate.cu pins
exactly these associations with __fmaf_rn/__fmul_rn/__fadd_rn, which is what makes it bit-identical.

    make -C oracle ref
    python tools/sass_symexec.py          # needs oracle/_ref/forward.o (test infrastructure, never shipped)

v0..v9 = the 10 view2gaussian floats of the Gaussian; rx0/ry0 = centre ray, rxm/rxp, rym/ryp = -+0.5 px rays.
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    obj = os.path.join(ROOT, "oracle", "_ref", "forward.o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    lines, on = [], False
    for l in sass.splitlines():
        if "Function :" in l:
            on = "integrateCUDA" in l
        m = re.match(r"^\s+/\*([0-9a-f]{4})\*/\s+(.*?)\s*/\*", l)
        if on and m:
            lines.append((m.group(1), m.group(2).rstrip(";").strip()))
    # the per-Gaussian loop body starts where the 10 quadric floats are read back from shared memory
    start = next(i for i, (_, t) in enumerate(lines) if t.startswith("LDS.64") and "+0x8]" in t)
    reg = {"R30": "ry0", "R31": "rx0", "R32": "rym", "R33": "rxm", "R34": "ryp", "R35": "rxp"}
    vmap = {0x0: ("v0", "v1"), 0x8: ("v2", "v3"), 0x10: ("v4", "v5"), 0x18: ("v6", "v7")}

    def R(x):
        x = x.strip()
        neg = x.startswith("-")
        x = x.lstrip("-").strip("|").replace(".reuse", "")
        v = "0" if x == "RZ" else reg.get(x, x)
        return "-" + v if neg else v

    seen = 0
    for addr, text in lines[start:start + 700]:
        if text.startswith("@"):
            text = text.split(None, 1)[1]
        op, _, args = text.partition(" ")
        a = [x.strip() for x in args.split(",")]
        base = op.split(".")[0]
        if op.startswith("LDS"):
            m = re.match(r"\[R\d+(\+0x([0-9a-f]+))?\]", a[1])
            off = int(m.group(2), 16) if m and m.group(2) else 0
            d = int(a[0][1:])
            if op == "LDS.64" and off in vmap:
                reg[f"R{d}"], reg[f"R{d + 1}"] = vmap[off]
            elif off == 0x20:
                reg[a[0]] = "v8"
            elif off == 0x24:
                reg[a[0]] = "v9"
        elif base == "FMUL":
            reg[a[0].replace(".reuse", "")] = f"({R(a[1])}*{R(a[2])})"
        elif base == "FADD":
            reg[a[0]] = f"({R(a[1])}+{R(a[2])})"
        elif op == "FFMA":
            reg[a[0]] = f"fma({R(a[1])},{R(a[2])},{R(a[3])})"
        elif base == "MOV" or op.startswith("IMAD.MOV"):
            reg[a[0]] = R(a[-1])
        elif op == "FCHK":
            seen += 1
            if seen % 2 == 0:          # second division of a ray: (-BB) / AA
                print(f"ray {seen // 2 - 1}:  -BB = {R(a[1])}\n         AA = {R(a[2])}\n")
            if seen == 10:
                break


if __name__ == "__main__":
    main()
