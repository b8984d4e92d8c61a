"""Torch restatement of the reference's post-network op sequence (src/gaussian_predictor.py:954-1008) -- test and
benchmark infrastructure: the reference module cannot travel to the GPU box, this runs the same torch operations in the
same order on whatever device the inputs live on (flatten_vector :788, bmm :964, activations :975-977,
quaternion_raw_multiply :45-63, transform_SHs :821-837, multi_view_union :796, make_contiguous :793)."""
import torch

from f3d_gaus_b200.predictor_head import ray_tables


def _flat(x):
    return x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)


def _qmul(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


class TorchHead:
    def __init__(self, cfg, device):
        m = cfg["model"]
        self.m = m
        res = int(m["training_resolution"])
        x, y = ray_tables(cfg)
        gx, gy = x[None, :].expand(res, res), y[:, None].expand(res, res)
        self.ray_dirs = torch.stack([gx, gy, torch.ones_like(gx)]).unsqueeze(0).to(device)
        v = torch.tensor([[0, 0, -1], [-1, 0, 0], [0, 1, 0]], dtype=torch.float32, device=device)
        self.v_to_sh, self.sh_to_v = v.unsqueeze(0), v.t().unsqueeze(0)
        self.split = ([3] if m["network_with_offset"] else []) + [1, 3, 4, 3] + ([9] if m["max_sh_degree"] > 0 else [])

    def __call__(self, net, depth, v2w, quat, B, V, squre_clip=10000.0):
        parts = list(net.split(self.split, dim=1))
        offset = parts.pop(0) if self.m["network_with_offset"] else 0.0
        opacity, scaling, rotation, dc = parts[:4]
        pos = self.ray_dirs.expand(depth.shape[0], 3, *self.ray_dirs.shape[2:]).clone() * depth + offset
        pos = _flat(pos)
        pos = torch.cat([pos, torch.ones((pos.shape[0], pos.shape[1], 1), device=pos.device)], dim=2)
        M = v2w.reshape(B * V, 4, 4)
        pos = torch.bmm(pos, M)
        pos = pos[:, :, :3] / (pos[:, :, 3:] + 1e-10)
        if squre_clip < 10.0:
            pos[:, :, 0].clamp_(-squre_clip, squre_clip)
            pos[:, :, 1].clamp_(-squre_clip, squre_clip)
        out = {"xyz": pos, "opacity": _flat(torch.sigmoid(opacity)), "scaling": _flat(torch.exp(scaling)),
               "rotation": _flat(torch.nn.functional.normalize(rotation)), "features_dc": _flat(dc).unsqueeze(2),
               "unet_depth": _flat(depth)}
        q = quat.reshape(B * V, 4)
        out["rotation"] = _qmul(q.unsqueeze(1).expand(*out["rotation"].shape), out["rotation"])
        if self.m["max_sh_degree"] > 0:
            rest = _flat(parts[4])
            rest = rest.reshape(*rest.shape[:2], -1, 3)
            b, n = rest.shape[:2]
            shs = rest.permute(0, 1, 3, 2).reshape(b, n * 3, 3)                     # 'b n sh rgb -> b (n rgb) sh'
            T = torch.bmm(torch.bmm(self.sh_to_v.expand(b, 3, 3), M[:, :3, :3]), self.v_to_sh.expand(b, 3, 3))
            out["features_rest"] = torch.bmm(shs, T).reshape(b, n, 3, 3).permute(0, 1, 3, 2)
        else:
            out["features_rest"] = torch.zeros((pos.shape[0], pos.shape[1], 0, 3), device=pos.device)
        out = {k: t.reshape(B, V, *t.shape[1:]).reshape(B, V * t.shape[1], *t.shape[2:]) for k, t in out.items()}
        return {k: t.contiguous() for k, t in out.items()}
