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
        view = lambda slab, o, n, shape: slab[o:o + n].view(torch.float32).view(shape)
        self.host = {k: view(self.host_slab, *l) for k, l in layout.items()}
        self.dev = {k: view(self.dev_slab, *l) for k, l in layout.items()}
        for k, v in pc.items():
            self.host[k].copy_(v.to(dtype=torch.float32))

    def upload(self) -> dict:
        self.dev_slab.copy_(self.host_slab, non_blocking=True)
        return self.dev
