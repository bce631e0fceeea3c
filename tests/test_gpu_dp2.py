"""Data-parallel train_step on real GPUs (skipped where the box has fewer): scripts/dp_equivalence_gpu.py under torchrun
with 2, 4 and 8 ranks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_step_equals_whole_batch_step(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs {} GPUs".format(world))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29541 + world), os.path.join(ROOT, "scripts", "dp_equivalence_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
