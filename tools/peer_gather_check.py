"""2+ GPU check of the fused pack + all-gather over NVLink peer memory (sharding.PeerFrameGather) against the
NCCL all_gather path, plus timing of both.  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_gather_check.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from f3d_gaus_b200 import cameras, sharding, synthetic
from f3d_gaus_b200.gaussian_renderer import render_views

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V, RES = 8, 256
cams = cameras.orbit_cameras(V)
wv, fp, cc = cams.world_view.to(dev), cams.full_proj.to(dev), cams.centers.to(dev)
cfg = synthetic.cfg_for(RES)
pc = {k: v.to(dev) for k, v in synthetic.f3d_like(rank, 128).items()}
o = render_views(pc, 0, wv, fp, cc, torch.zeros(3, device=dev), cfg, epilogue=False)
raster = o["raster"]                                                    # [V,9,H,W]
local_frames = sharding.pack_frames(o["render"].unsqueeze(0), o["rendered_depth"].unsqueeze(0), o["rendered_alpha"].unsqueeze(0))
want = sharding.gather_frames(local_frames, world)                      # NCCL path: [world, V, 5, H, W]
for mc in (False, True):
    pg = sharding.PeerFrameGather(world, V, RES, RES, dev, use_multicast=mc)
    got = pg.push(raster, first_scene=rank)
    torch.cuda.synchronize()
    ok = torch.equal(got, want)
    # write-after-read across steps: a rank that races ahead must not overwrite frames a slower peer is still
    # reading.  Rank 0 is slowed down by a long-running kernel between "consume step k" and "push step k+1".
    for step in range(6):
        scaled = raster * float(step + 2)
        g = pg.push(scaled, first_scene=rank)
        if rank == 0:
            torch.cuda._sleep(20_000_000)                 # ~10 ms of device time before the consumer reads
        snap = g.clone()                                   # the consumer's read, stream-ordered after the push
        ok = ok and torch.equal(snap, want * float(step + 2))   # (no collective here: the ranks must be free to drift)
    torch.cuda.synchronize()
    def t(fn, n=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
    t_peer = t(lambda: pg.push(raster, first_scene=rank))
    t_nccl = t(lambda: sharding.gather_frames(sharding.pack_frames(o["render"].unsqueeze(0), o["rendered_depth"].unsqueeze(0),
                                                                 o["rendered_alpha"].unsqueeze(0)), world))
    if rank == 0:
        print(f"multicast requested={mc} used={bool(pg.multicast)}: equal to NCCL gather: {ok}; fused peer gather {t_peer:.1f} us, "
              f"pack + NCCL all_gather {t_nccl:.1f} us (world {world})", flush=True)
    assert ok
dist.destroy_process_group()
