"""NCCL parity under pytest: launches tests/multi_gpu_check.py with torch.distributed.run on min(device_count, 4) GPUs
(K1 genome blocks, K2 OR-merge by bit range, K3 / K4 bit-range shards with all_reduce, K4 with sharded hashing and
all_to_all -- all against the oracle).  Skipped on a one-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
def test_nccl_parity_on_all_local_gpus():
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "multi-GPU parity ok: world=%d" % n in out.stdout
