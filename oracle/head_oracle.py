"""CPU oracle (numpy, float32) for the predictor OUTPUT HEAD -- TEST INFRASTRUCTURE, never shipped or imported
by the product.  Only tests/, __graft_entry__.smoke() and tools/ may use it.

Restates what `GaussianSplatPredictor_gtunet.forward` does AFTER the UNet
(/root/reference/src/gaussian_predictor.py:954-1008), one float32 operation per torch operation, in torch's order:

  ray grid            init_ray_dirs                      :657-681
  channel split       get_splits_and_inits               :683-728   ([3,]1,3,4,3[,9])
  position            get_pos_from_network_output        :857-881   ray * depth + offset
  to world            cat 1, bmm view_to_world, / (w+1e-10), squre_clip   :959-970
  activations         sigmoid / exp / normalize(dim=1)   :636-638, :975-977
  rotation to world   transform_rotations -> quaternion_raw_multiply(Mq, q)   :45-63, :839-855
  SH to world         transform_SHs (degree 1)           :821-837
  multi_view_union    [B*V,N,.] -> [B,V*N,.]             :796-800

Pinned against tests/golden/head/head_*.npz, which tests/golden/make_head_golden.py produced by running the reference's
forward itself (unmodified, CPU) on seeded inputs: tests/test_head.py::test_head_oracle_vs_golden.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def ray_tables(res: int, fov_deg: float, inverted_x: bool, inverted_y: bool):
    """x[W], y[H] of init_ray_dirs (:657-681): linspace ends (-res)//2 + 0.5 and res//2 - 0.5 (unit step, exact in
    float32), optional sign flips, then divided by the focal length as a float32 tensor / python scalar division."""
    lo, hi = (-res) // 2 + 0.5, res // 2 - 0.5
    x = np.linspace(lo, hi, res, dtype=np.float64).astype(f32)
    y = np.linspace(hi, lo, res, dtype=np.float64).astype(f32)
    if inverted_x:
        x = -x
    if inverted_y:
        y = -y
    focal = res / (2 * math.tan((fov_deg * np.pi / 180) / 2))        # fov2focal, python double
    return (x / f32(focal)).astype(f32), (y / f32(focal)).astype(f32)


def sh_transform(view_to_world: np.ndarray) -> np.ndarray:
    """transforms of transform_SHs (:821-834): sh_to_v @ V2W[:3,:3] @ v_to_sh, [BV,3,3] float32."""
    v_to_sh = np.array([[0, 0, -1], [-1, 0, 0], [0, 1, 0]], dtype=f32)
    sh_to_v = v_to_sh.T
    R = view_to_world.reshape(-1, 4, 4)[:, :3, :3].astype(f32)
    return np.einsum("ij,bjk,kl->bil", sh_to_v, R, v_to_sh).astype(f32)


def head(net, depth, view_to_world, quat, *, B, V, res, fov_deg, with_offset, sh_degree, isotropic=False,
         inverted_x=False, inverted_y=True, squre_clip=10000.0, const_offset=None):
    """net [B*V,C,res,res], depth [B*V,1,res,res], view_to_world [B,V,4,4], quat [B,V,4] -> dict of float32 arrays
    with the reference's keys and shapes ([B, V*N, ...])."""
    BV, N = B * V, res * res
    net = np.asarray(net, dtype=f32).reshape(BV, -1, N)
    at = 0
    if with_offset:
        offset, at = net[:, 0:3], 3
    opacity, scaling, rotation, dc = net[:, at:at + 1], net[:, at + 1:at + 4], net[:, at + 4:at + 8], net[:, at + 8:at + 11]
    rest = net[:, at + 11:at + 20] if sh_degree > 0 else None
    rx, ry = ray_tables(res, fov_deg, inverted_x, inverted_y)
    ray = np.stack([np.broadcast_to(rx[None, :], (res, res)), np.broadcast_to(ry[:, None], (res, res)),
                    np.ones((res, res), f32)]).reshape(1, 3, N)
    d = np.asarray(depth, dtype=f32).reshape(BV, 1, N)
    if const_offset is not None:
        d = d + np.asarray(const_offset, dtype=f32).reshape(BV, 1, N)
    pos = ray * d + (offset if with_offset else f32(0.0))                       # [BV,3,N]
    pos = np.transpose(pos, (0, 2, 1))                                          # flatten_vector
    hom = np.concatenate([pos, np.ones((BV, N, 1), f32)], axis=2)
    M = np.asarray(view_to_world, dtype=f32).reshape(BV, 4, 4)
    hom = np.einsum("bnk,bkj->bnj", hom, M).astype(f32)
    xyz = hom[:, :, :3] / (hom[:, :, 3:] + f32(1e-10))
    if squre_clip < 10.0:
        xyz[:, :, 0] = np.clip(xyz[:, :, 0], f32(-squre_clip), f32(squre_clip))
        xyz[:, :, 1] = np.clip(xyz[:, :, 1], f32(-squre_clip), f32(squre_clip))
    if isotropic:
        scaling = np.concatenate([scaling[:, :1]] * 3, axis=1)
    out = {
        "xyz": xyz,
        "opacity": np.transpose(f32(1.0) / (f32(1.0) + np.exp(-opacity)), (0, 2, 1)),
        "scaling": np.transpose(np.exp(scaling), (0, 2, 1)),
        "features_dc": np.transpose(dc, (0, 2, 1))[:, :, None, :],
        "unet_depth": np.transpose(np.asarray(depth, dtype=f32).reshape(BV, 1, N), (0, 2, 1)),
    }
    norm = np.sqrt((rotation * rotation).sum(axis=1, keepdims=True, dtype=f32))
    q = np.transpose(rotation / np.maximum(norm, f32(1e-12)), (0, 2, 1))        # F.normalize(dim=1), flattened
    a = np.asarray(quat, dtype=f32).reshape(BV, 1, 4)
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    out["rotation"] = np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                                aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], axis=-1)
    if sh_degree > 0:
        shs = np.transpose(rest, (0, 2, 1)).reshape(BV, N, 3, 3)                # [b n sh rgb]
        T = sh_transform(M)                                                     # [b sh sh']
        out["features_rest"] = np.einsum("bnsr,bst->bntr", shs, T).astype(f32)
    else:
        out["features_rest"] = np.zeros((BV, N, 0, 3), f32)
    return {k: np.ascontiguousarray(v.reshape((B, V * N) + v.shape[2:]), dtype=f32) for k, v in out.items()}
