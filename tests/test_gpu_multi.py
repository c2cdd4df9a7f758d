"""Real multi-GPU check (skipped with < 2 GPUs): band-sharded frames over NCCL / peer stores equal the single-GPU frame."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_band_sharded_frame_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29431", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "multi-GPU OK" in p.stdout
