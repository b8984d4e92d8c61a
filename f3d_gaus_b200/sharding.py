"""Scene-sharded multi-GPU runner: one process per GPU (`torch.distributed`), no data-path collective.

The reference is single-process / single-GPU and loops `for bb in range(bs)` serially
(visualize.py:297,393).  Frames of different scenes are independent (no shared parameters, no
cross-scene reads), so scenes are partitioned in contiguous blocks -- all views of a scene render on
the GPU that holds its Gaussians, which also keeps the cycle-aggregative concat local -- and the only
exchange step is ONE gather of the rendered frames (NCCL over NVLink on GPUs; the same code runs on
`gloo` with CPU tensors for the host-logic tests).  Only the consumed channels travel: rgb, depth,
alpha (what visualize.py:304-306 reads), 5 of the 9 raster channels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

GATHER_CHANNELS = 5   # rgb(3) + median depth + alpha


def shard_range(num_scenes: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of scenes owned by `rank`: [lo, hi).  Block size ceil(num_scenes / world); trailing
    ranks may own fewer (or no) scenes."""
    per = -(-num_scenes // world)
    lo = min(num_scenes, rank * per)
    return lo, min(num_scenes, lo + per)


def owner_of(scene: int, num_scenes: int, world: int) -> int:
    return scene // (-(-num_scenes // world))


def pack_frames(rgb: torch.Tensor, depth: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """[B,V,3,H,W], [B,V,1,H,W], [B,V,1,H,W] -> [B,V,5,H,W] (the gathered payload)."""
    return torch.cat([rgb, depth, alpha], dim=2).contiguous()


def gather_frames(local: torch.Tensor, num_scenes: int, group=None) -> torch.Tensor:
    """All ranks contribute their [B_local, V, C, H, W] block; every rank receives [num_scenes, V, C, H, W] in
    scene order.  One collective (all_gather_into_tensor); shards are padded to the common block size."""
    if not (dist.is_available() and dist.is_initialized()):
        assert local.shape[0] == num_scenes
        return local
    world = dist.get_world_size(group)
    per = -(-num_scenes // world)
    tail = tuple(local.shape[1:])
    send = local
    if local.shape[0] != per:
        send = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
        send[:local.shape[0]].copy_(local)
    recv = torch.empty((world * per,) + tail, dtype=local.dtype, device=local.device)
    try:
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):      # backends without the flat variant
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send.contiguous(), group=group)
        recv = torch.cat(parts, dim=0)
    return recv[:num_scenes]


class PeerFrameGather:
    """The exchange step fused into one kernel over NVLink peer memory: `push(raster, first_scene)` packs the five
    consumed channels of this rank's frames and stores them directly into EVERY rank's gather buffer (peer
    pointers of a torch symmetric-memory allocation; one multimem.st per 16 bytes when the NVSwitch multicast
    address exists), then a symmetric-memory barrier publishes them.  Replaces 3 pack copies + all_gather.
    Needs CUDA + an initialised NCCL process group; `gather_frames` (NCCL / gloo) is the portable path.

    The gather buffer is DOUBLE-BUFFERED: push k writes slot k % 2.  A rank can only start push k+2 (which
    overwrites the slot of push k) after it has passed the barrier of push k+1, and every peer reaches that barrier
    on its stream only after the work it enqueued before it -- including its reads of push k's result.  So a fast
    rank never stores into a buffer a slower peer is still reading, provided the consumer reads the returned tensor
    on the stream `push` was called on (or orders its reads before the next push on that stream), and the returned
    tensor is consumed before the second-next push."""

    def __init__(self, num_scenes: int, V: int, H: int, W: int, device, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.shape = (num_scenes, V, GATHER_CHANNELS, H, W)
        self.buf = symm_mem.empty((2,) + self.shape, dtype=torch.float32, device=self.device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.slot_bytes = self.buf[0].numel() * 4
        self.slot = 0
        base = [int(p) for p in self.hdl.buffer_ptrs]
        self.ptrs = [torch.tensor([p + k * self.slot_bytes for p in base], dtype=torch.int64, device=self.device)
                     for k in range(2)]
        mc = 0
        try:
            if use_multicast and self.hdl.has_multicast_support(self.device.type, self.device.index or 0):
                mc = int(self.hdl.multicast_ptr)
        except Exception:
            mc = 0
        self.multicast = mc

    def push(self, raster: torch.Tensor, first_scene: int) -> torch.Tensor:
        """raster: [B_local, V, 9, H, W] (or [V,9,H,W] for one scene).  Returns the gathered buffer (valid after
        the barrier this call ends with, stream-ordered on the current stream)."""
        num_scenes, V, C, H, W = self.shape
        r = raster.reshape(-1, 9, H * W).contiguous()
        import ctypes
        k = self.slot
        self.slot ^= 1
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self._lib.lib.gof_pack_gather(r.data_ptr(), r.shape[0], H * W, self.ptrs[k].data_ptr(), self.world,
                                               ctypes.c_void_p(self.multicast + k * self.slot_bytes) if self.multicast else None,
                                               first_scene * V, stream)
            self._lib.check(rc, "gof_pack_gather")
            self.hdl.barrier()
        return self.buf[k]


def render_sharded(make_scene, num_scenes: int, cams, cfg: dict, background: torch.Tensor, *, rank: int, world: int,
                   workspace=None, render_fn=None, group=None, gather: bool = True):
    """Render every view of every scene of a global batch, scene-sharded.

    make_scene(lo, hi) -> dict of [hi-lo, P, .] tensors on this rank's device (the rank's block of scenes: in
    the real pipeline the predictor's output for its share of the input images).
    Returns ([num_scenes, V, 5, H, W] on every rank if `gather`, else this rank's [B_local, V, 5, H, W])."""
    from .cycle import render_scene_views
    lo, hi = shard_range(num_scenes, world, rank)
    V = cams.world_view.shape[0]
    H = W = int(cfg["model"]["training_resolution"])
    if hi > lo:
        pc = make_scene(lo, hi)
        rgb, depth, alpha = render_scene_views(pc, cams, cfg, background, workspace, render_fn)
        local = pack_frames(rgb, depth, alpha)
    else:
        local = torch.zeros((0, V, GATHER_CHANNELS, H, W), dtype=torch.float32, device=background.device)
    return gather_frames(local, num_scenes, group) if gather else local
