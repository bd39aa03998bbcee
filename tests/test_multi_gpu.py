"""N > 1 on real GPUs (skipped on a single-GPU box): uid shards over two ranks, the NCCL all-reduce of the volumes on the library's
collective stream, and a failing rank joining the collective instead of leaving its peers inside it (tools/check_nccl_volumes.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_volumes_and_failed_rank_flag():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "check_nccl_volumes.py")], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "nccl volumes ok: 2 ranks" in r.stdout
