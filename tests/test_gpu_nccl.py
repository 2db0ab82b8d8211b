"""The N>1 path under NCCL on real GPUs (the CPU suite covers the same logic over gloo, tests/test_host_logic.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_run_with_nccl_gather_equals_one_rank():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box (run with `gpurun --gpus 2`); the gather is covered over gloo on the CPU")
    world = 2 if n < 4 else 4
    env = dict(os.environ, LP_NCCL_FRAMES="256")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"NCCL_OK world={world} frames=256" in out.stdout, out.stdout[-2000:]
