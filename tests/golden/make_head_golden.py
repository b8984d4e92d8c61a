"""Generate tests/golden/head/head_*.npz: the post-network part of the reference predictor
(`GaussianSplatPredictor_gtunet.forward`, /root/reference/src/gaussian_predictor.py:883-1008) run UNMODIFIED on CPU.

    python tests/golden/make_head_golden.py            # in the build container (needs /root/reference)

How the reference is driven without its network and without a GPU:
  * the module file is loaded on its own (importing the `src` package pulls omegaconf/prettytable, absent here);
  * the predictor object is created without `__init__` (which builds the UNet); only what the head uses is set up,
    through the reference's OWN methods (`init_ray_dirs`, `init_sh_transform_matrices`, its activation choices are
    copied from :636-638);
  * the UNet is replaced by a stub that returns the seeded "network output" tensor stored in the fixture;
  * `torch.ones` / `torch.zeros` are wrapped for the duration of the call so that the reference's hard-coded
    `device="cuda"` (:963, :999) lands on the CPU.
Each fixture holds the inputs (network output, depth, view_to_world, quaternions, cfg scalars) and every tensor of the
returned dict.  `oracle/head_oracle.py` and the CUDA head (`gof_predictor_head`) are pinned against these files.
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/gaussian_predictor.py"

CASES = {
    # name: (B, V, res, with_offset, sh_degree, isotropic, inverted_x, inverted_y, squre_clip, origin_distances, seed)
    "head_offset_sh1_r32": (2, 1, 32, True, 1, False, False, True, 10000.0, False, 0),      # the shipped config's path
    "head_nooffset_sh1_r24_v2": (1, 2, 24, False, 1, False, False, True, 10000.0, False, 1),
    "head_offset_sh0_iso_clip_r16": (2, 1, 16, True, 0, True, True, False, 0.05, False, 2),
    "head_nooffset_origin_r16": (1, 1, 16, False, 1, False, False, True, 10000.0, True, 3),
}


def load_reference():
    if "torchvision" not in sys.modules:
        try:
            import torchvision  # noqa: F401
        except Exception:
            sys.modules["torchvision"] = types.ModuleType("torchvision")
    spec = importlib.util.spec_from_file_location("ref_gaussian_predictor", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def random_rigid(g):
    """A row-vector-convention view_to_world like visualize.py:253-258 builds: [R^T rows; t] with last column (0,0,0,1)."""
    q = torch.randn(4, generator=g)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=torch.float32)
    M = torch.eye(4)
    M[:3, :3] = R.t()
    M[3, :3] = torch.randn(3, generator=g) * 2.0
    return M


def make_inputs(B, V, res, with_offset, sh_degree, origin_distances, seed):
    g = torch.Generator().manual_seed(seed)
    C = (3 if with_offset else 0) + 1 + 3 + 4 + 3 + (9 if sh_degree > 0 else 0)
    net = torch.randn(B * V, C, res, res, generator=g)
    at = 3 if with_offset else 0
    if with_offset:
        net[:, 0:3] *= 0.01                                   # offsets are small
    net[:, at + 1:at + 4] = net[:, at + 1:at + 4] * 0.3 + math.log(0.01)     # log-scales around 0.01
    depth = 6.667 + 2.0 * torch.rand(B * V, 1, res, res, generator=g)
    x_in = torch.randn(B, V, 4 if origin_distances else 3, res, res, generator=g)
    v2w = torch.stack([random_rigid(g) for _ in range(B * V)]).reshape(B, V, 4, 4)
    quat = torch.randn(B, V, 4, generator=g)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    return net, depth, x_in, v2w, quat


def run_reference(mod, cfg, net, depth, x_in, v2w, quat, squre_clip):
    cls = mod.GaussianSplatPredictor_gtunet
    p = cls.__new__(cls)
    torch.nn.Module.__init__(p)
    p.cfg = cfg
    cls.get_splits_and_inits(p, cfg["model"]["network_with_offset"], cfg)
    p.init_ray_dirs()
    p.depth_act = torch.nn.Sigmoid()
    p.scaling_activation = torch.exp                      # :636-638
    p.opacity_activation = torch.sigmoid
    p.rotation_activation = torch.nn.functional.normalize
    if cfg["model"]["max_sh_degree"] > 0:
        p.init_sh_transform_matrices()

    class Net(torch.nn.Module):
        def forward(self, x, film_camera_emb=None, N_views_xa=1):
            return net

    if cfg["model"]["network_with_offset"]:
        p.network_with_offset = Net()
    else:
        p.network_wo_offset = Net()

    real_ones, real_zeros = torch.ones, torch.zeros

    def on_cpu(fn):
        def wrapped(*a, **k):
            if k.get("device") == "cuda":
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped

    torch.ones, torch.zeros = on_cpu(real_ones), on_cpu(real_zeros)
    try:
        with torch.no_grad():
            out = p(x_in, v2w, quat, focals_pixels=None, squre_clip=squre_clip, unet_depth=depth)
    finally:
        torch.ones, torch.zeros = real_ones, real_zeros
    return out, p.ray_dirs


def main():
    mod = load_reference()
    for name, (B, V, res, with_offset, sh, iso, inv_x, inv_y, clip, origin, seed) in CASES.items():
        cfg = {"model": {"training_resolution": res, "fov": 13.164, "inverted_x": inv_x, "inverted_y": inv_y,
                         "max_sh_degree": sh, "isotropic": iso, "cross_view_attention": True,
                         "origin_distances": origin, "network_with_offset": with_offset,
                         "network_without_offset": not with_offset, "xyz_scale": 1e-6, "xyz_bias": 0.0,
                         "opacity_scale": 1e-3, "opacity_bias": -3.0, "scale_scale": 5e-4, "scale_bias": 0.01}}
        net, depth, x_in, v2w, quat = make_inputs(B, V, res, with_offset, sh, origin, seed)
        out, ray_dirs = run_reference(mod, cfg, net, depth, x_in, v2w, quat, clip)
        blob = {"in_net": net.numpy(), "in_depth": depth.numpy(), "in_view_to_world": v2w.numpy(),
                "in_quat": quat.numpy(), "in_x": x_in.numpy(), "ray_dirs": ray_dirs.numpy(),
                "cfg": np.array([B, V, res, int(with_offset), sh, int(iso), int(inv_x), int(inv_y), int(origin)], dtype=np.int64),
                "fov": np.float64(13.164), "squre_clip": np.float64(clip)}
        for k, v in out.items():
            blob["out_" + k] = v.contiguous().numpy()
        path = os.path.join(HERE, "head", name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, {k: tuple(v.shape) for k, v in out.items()}, f"{os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
