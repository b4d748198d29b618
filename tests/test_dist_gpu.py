"""Multi-GPU parity on real GPUs as a `-m gpu` test: spawns tests/dist_gpu_check.py under torch.distributed.run with one
rank per GPU (NCCL + the NVLink peer-memory halo / all-reduce kernels of csrc/halo.cu) when at least two GPUs are visible;
skipped on single-GPU boxes (the sharding logic itself runs on one GPU in tests/test_full_size.py and on the CPU with gloo
in tests/test_dist_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_callbacks_over_nccl_and_peer_memory(world):
    n = _ngpu()
    if n < world:
        pytest.skip(f"needs {world} GPUs, {n} visible")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    err = "\n".join(l for l in r.stderr.splitlines() if not l.startswith(("W1", "W0", "*", "Setting OMP")))
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + err[-6000:]
    assert r.stdout.count(" OK ") >= 5, r.stdout[-3000:]
