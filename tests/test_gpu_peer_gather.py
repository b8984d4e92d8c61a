"""Fused pack + all-gather over NVLink peer memory (sharding.PeerFrameGather / gof_pack_gather) against the NCCL
all_gather path.  Needs at least 2 GPUs (skipped on the single-GPU test box); run by hand with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_peer_gather.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_gather_equals_nccl_gather():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "peer_gather_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "equal to NCCL gather: True" in res.stdout
