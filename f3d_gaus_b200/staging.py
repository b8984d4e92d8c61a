"""Host-side staging of a Gaussian set (pure torch: importing this module does not load the CUDA library, so the
reference arm of bench.py can use the same upload path)."""
from __future__ import annotations

import torch


class PinnedScene:
    """A Gaussian set ([B,P,.] per key, the predictor's output dict) staged in ONE pinned host slab with a mirror slab on
    the device: `upload()` is a single async copy instead of one per key (six for the reference's dict), and returns
    the device-side dict of views that the render functions take.  `host[k]` are the writable pinned views."""

    def __init__(self, pc: dict, device):
        self.device = torch.device(device)
        at, layout = 0, {}
        for k, v in pc.items():
            nbytes = v.numel() * 4
            layout[k] = (at, nbytes, tuple(v.shape))
            at += -(-nbytes // 256) * 256
        self.nbytes = sum(n for _, n, _ in layout.values())
        self.host_slab = torch.empty(max(at, 256), dtype=torch.uint8).pin_memory()
        self.dev_slab = torch.empty(max(at, 256), dtype=torch.uint8, device=self.device)
        self.layout = layout
        self.host = self.views(self.host_slab)
        self.dev = self.views(self.dev_slab)
        for k, v in pc.items():
            self.host[k].copy_(v.to(dtype=torch.float32))

    def views(self, slab: torch.Tensor) -> dict:
        """The per-key tensor views of a slab (host or device) that has this scene's layout."""
        return {k: slab[o:o + n].view(torch.float32).view(shape) for k, (o, n, shape) in self.layout.items()}

    def upload(self, dst_slab: torch.Tensor | None = None) -> dict:
        """One async H2D copy on the current stream into the scene's own device slab, or into `dst_slab` (a device
        byte tensor of at least the slab size: streaming loops keep one per pipeline slot)."""
        if dst_slab is None:
            self.dev_slab.copy_(self.host_slab, non_blocking=True)
            return self.dev
        dst = dst_slab[:self.host_slab.numel()]
        dst.copy_(self.host_slab, non_blocking=True)
        return self.views(dst)
