"""Generate the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference
rasterizer (oracle/_ref/libgof_ref.so, built by oracle/Makefile from /root/reference) on a GPU.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'      # on the B200 box
    cp gpurun_out/golden/*.npz tests/golden/                                # back here

Each file holds the inputs, the reference's forward outputs, its decoded forward state and its
backward outputs for a fixed seeded dL_dout_color (cases.grad_seed).  The CPU oracle
(oracle/gof_oracle.c) and the CUDA library are both pinned against these files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
import refgpu  # noqa: E402


# golden cases that also carry the outputs of Rasterizer::integrate: name -> number of query points
INTEGRATE_CASES = {"unit_p1200_120x88_sh3": 4000, "f3d_s32_r96_colors_ks": 3000}


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, build in cases.GOLDEN_CASES.items():
        c_cpu = build()
        c = cases.case_to(c_cpu, "cuda")
        run = refgpu.RefRun()
        fwd = run.forward(c)
        dL = cases.grad_seed(c)
        bwd = run.backward(c, dL)
        blob = {}
        for k, v in c_cpu.items():
            blob["in_" + k] = v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        blob["in_dL_dout"] = dL.cpu().numpy()
        for k, v in fwd.items():
            blob["fwd_" + k] = v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        for k, v in bwd.items():
            blob["bwd_" + k] = v.cpu().numpy()
        if name in INTEGRATE_CASES:
            # point integration (Rasterizer::integrate) on seeded query points around the Gaussians
            g = torch.Generator().manual_seed(11)
            P = c_cpu["means3D"].shape[0]
            idx = torch.randint(0, P, (INTEGRATE_CASES[name],), generator=g)
            pts = c_cpu["means3D"][idx] + 2.0 * c_cpu["scales"][idx].max(dim=1, keepdim=True).values * \
                torch.randn(len(idx), 3, generator=g)
            pts[:20] += 50.0
            integ = refgpu.ref_integrate(c, pts.to("cuda"))
            blob["in_points3D"] = pts.numpy()
            for k in ("out_color", "alpha_integrated", "color_integrated"):
                blob["int_" + k] = integ[k].cpu().numpy()
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, "R =", fwd["num_rendered"], "visible =", int((fwd["radii"] > 0).sum()),
              os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
