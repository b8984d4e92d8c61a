"""f3d_gaus_b200 -- B200-native (sm_100a) GOF Gaussian rasterizer + render loop for F3D-Gaus.

Layout (only what the hot path needs):
  csrc/                    hand-written CUDA kernels + the extern "C" ABI (include/gof_b200.h)
  _lib.py                  ctypes loader of libgof_b200.so (fails loudly if it is missing)
  diff_gof_rasterization/  drop-in for the reference's rasterizer package
  gaussian_renderer/       drop-in for src/gaussian_renderer (render_predicted_more_v2_gof, render)
  cycle.py                 the cycle-aggregative render loop of visualize.py:281-340
  predictor_head.py        the predictor's post-network output head (src/gaussian_predictor.py:954-1008), one kernel
  sharding.py              scene-sharded multi-GPU runner (one process per GPU, fused peer-memory gather, NCCL fallback)
  staging.py               PinnedScene: a Gaussian set in one pinned host slab, one H2D copy (pure torch)
  ply.py                   Gaussian-set export helpers (visualize.py:146-179)
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_drop_in() -> None:
    """Register this package's modules under the reference's import names, so that
    `from diff_gof_rasterization import GaussianRasterizationSettings_GOF, GaussianRasterizer_GOF`
    (src/gaussian_renderer/__init__.py:10) resolves to the B200 implementation."""
    from . import diff_gof_rasterization as _dgr

    sys.modules["diff_gof_rasterization"] = _dgr
