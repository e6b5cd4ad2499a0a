"""Multi-GPU parity under pytest: launches tests/dist_check_gpu.py with torchrun on 2 (and 4 / 8 when present) GPUs of
the box - the NCCL all-to-all / merge / all-gather path must reproduce the single-GPU search exactly.  Skipped on a
single-GPU box (the gloo tests in test_sharded_cpu.py cover the host logic there)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_nccl_equals_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check_gpu.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "dist check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
