"""Multi-GPU path (image tiles per rank + one ncclReduce inside rl_render); needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_nccl_reduce_equals_single_gpu():
    world = min(_ngpu(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "tests", "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"MGPU_OK {world}" in r.stdout
