"""Multi-GPU correctness of the real model on the CUDA path: N NCCL ranks, each running its shard of
molecules, all-reduce of the flat gradient buffer == the single-GPU gradient of the whole batch.  Needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a 1-GPU box.  The host-side logic
alone is covered on CPU with gloo (tests/test_parallel_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", [2])
def test_nccl_sharded_gradients_equal_single_gpu(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_GRAD_CHECK" in r.stdout
