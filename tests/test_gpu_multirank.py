"""2-rank NCCL test of the sharded sampling path and the DDP gradient all-reduce on real GPUs
(skipped on a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent


def test_sharded_sampling_and_gradient_allreduce_nccl_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(HERE / "multirank_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0 and "MULTIRANK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
