"""Host-side tests of the scene-sharded runner and the cycle-aggregative loop (CPU, no GPU): the
partition, the single gather over a world-size-2 `gloo` group, and the loop's bookkeeping with a
fake renderer standing in for the CUDA rasterizer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from f3d_gaus_b200 import cameras, cycle, sharding, synthetic


def test_shard_ranges_partition():
    for n in (1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
                for s in range(lo, hi):
                    assert sharding.owner_of(s, n, world) == r
            assert got == list(range(n))


def fake_render(pc, bs, wv, fp, cc, bg, cfg, workspace=None, epilogue=False, **kw):
    """Deterministic stand-in for render_views: every output pixel encodes (scene content, view)."""
    V = wv.reshape(-1, 16).shape[0]
    H = W = cfg["model"]["training_resolution"]
    tag = pc["xyz"][bs].sum() * 1e-3
    vt = wv.reshape(V, 16).sum(dim=1).reshape(V, 1, 1, 1)
    base = torch.ones(V, 9, H, W) * tag + vt
    return {"render": base[:, 0:3] + 0.25, "rendered_depth": base[:, 6:7] + 7.0, "rendered_alpha": base[:, 7:8] * 0 + 0.5,
            "raster": base}


def _scene_block(lo, hi, S=8):
    sets = [synthetic.f3d_like(seed=s, S=S) for s in range(lo, hi)]
    return {k: torch.cat([x[k] for x in sets], dim=0) for k in sets[0]}


def _worker(rank, world, port, n_scenes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cams = cameras.orbit_cameras(4)
    cfg = synthetic.cfg_for(16)
    frames = sharding.render_sharded(_scene_block, n_scenes, cams, cfg, torch.zeros(3), rank=rank, world=world,
                                     render_fn=fake_render)
    torch.save(frames, os.path.join(out_dir, f"frames_{rank}.pt"))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_scenes", [4, 3])
def test_gather_world2_gloo(tmp_path, n_scenes):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_scenes, str(tmp_path)), nprocs=world, join=True)
    cams = cameras.orbit_cameras(4)
    cfg = synthetic.cfg_for(16)
    # single-process result = what every rank must hold after the gather
    want = sharding.render_sharded(_scene_block, n_scenes, cams, cfg, torch.zeros(3), rank=0, world=1, render_fn=fake_render)
    assert want.shape == (n_scenes, 4, sharding.GATHER_CHANNELS, 16, 16)
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"frames_{r}.pt"))
        assert torch.equal(got, want), f"rank {r}"


def test_cycle_aggregate_bookkeeping():
    """visualize.py:288-340: K views rendered per scene, K predictor calls with that view's transform and
    quaternion, sets concatenated on dim 1 in view order, rgb clamped to [0,1] before re-prediction."""
    B, S, res, K = 2, 8, 8, 3
    pc = _scene_block(0, B, S)
    cams = cameras.orbit_cameras(K)
    cfg = synthetic.cfg_for(res)
    calls = []
    base = cycle.unproject_predictor(cfg)

    def predict(novel_img, v2w, quat, depth):
        assert novel_img.shape == (B, 1, 4, res, res) and depth.shape == (B, 1, res, res)
        assert v2w.shape == (B, 1, 4, 4) and quat.shape == (B, 1, 4)
        assert float(novel_img[:, :, 0:3].min()) >= 0.0 and float(novel_img[:, :, 0:3].max()) <= 1.0
        calls.append((v2w[0, 0].clone(), quat[0, 0].clone()))
        return base(novel_img, v2w, quat, depth)

    merged, frames = cycle.cycle_aggregate(pc, predict, cams, cfg, torch.zeros(3), render_fn=fake_render)
    assert len(calls) == K
    for k in range(K):
        assert torch.equal(calls[k][0], cams.view_to_world[k])
        want_q = cameras.matrix_to_quaternion(cams.view_to_world[k][:3, :3].T.contiguous())
        assert torch.allclose(calls[k][1], want_q, atol=1e-6)
    P0 = S * S
    for key in cycle.PC_KEYS:
        assert merged[key].shape[0] == B and merged[key].shape[1] == P0 + K * res * res
        assert torch.equal(merged[key][:, :P0], pc[key])           # the source set comes first, untouched
    assert frames["rgb"].shape == (B, K, 3, res, res)
    # the re-predicted Gaussians of view k sit at the un-projected depth in that view's frame
    k = 1
    blk = merged["xyz"][:, P0 + k * res * res:P0 + (k + 1) * res * res]
    back = torch.cat([blk, torch.ones(B, res * res, 1)], dim=-1) @ cams.world_view[k]
    assert torch.allclose(back[..., 2], frames["depth"][:, k].reshape(B, -1), atol=1e-4)
    assert torch.allclose(merged["rotation"].norm(dim=-1), torch.ones(B, merged["rotation"].shape[1]), atol=1e-5)


def test_view_quaternions_cached_per_camera_tensor():
    """cycle.view_quaternions reads the matrices back once per camera set; an in-place change of the tensor (its
    version counter) or another tensor gives a fresh result."""
    from f3d_gaus_b200 import cameras, cycle
    cams = cameras.orbit_cameras(4)
    v2w = cams.view_to_world.clone()
    q1 = cycle.view_quaternions(v2w)
    assert cycle.view_quaternions(v2w) is q1            # cached: same object back
    want = torch.stack([cameras.matrix_to_quaternion(m[:3, :3].T.contiguous()) for m in v2w])
    assert torch.allclose(q1, want, atol=1e-6)
    v2w[0].copy_(v2w[2])                                # in-place edit bumps the version counter
    q2 = cycle.view_quaternions(v2w)
    assert q2 is not q1 and torch.allclose(q2[0], want[2], atol=1e-6)
    other = cams.view_to_world.clone()
    assert torch.allclose(cycle.view_quaternions(other), want, atol=1e-6)
