"""BASELINE configs[4] / SURVEY §8e: ONE map tiled over several GPUs (giant.lsd_tiled: row bands of the stencil stage, NCCL
broadcasts of the bands, max-all-reduce of maxGrad, region stages on rank 0) equals the single-GPU result bit for bit.
Needs >= 2 GPUs: launched as a 2-rank torchrun from inside the test (skipped on a one-GPU box; tools/giant_check.py is the
same check as a script, its log is kept under profiles/)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tiled_map_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", os.path.join(ROOT, "tools", "giant_check.py"), "--size", "4096"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "TILED == SINGLE" in r.stdout
