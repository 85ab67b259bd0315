"""Multi-GPU parity of cuda.jl_b200/multigpu.py on REAL GPUs (NCCL + NVLink peer stores): launches
tests/mgpu_check.py under torchrun on min(#GPUs, 8) ranks and requires its "ALL OK".  Self-skips on boxes with fewer
than two GPUs (the driver's single-GPU box); run it with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`.
mgpu_check.py compares every sharded result on the host with the CPU oracle: bit-exact for integer reductions / scans,
sort! and sortperm (stable across ranks), stated tolerance for Float32 sums and scans."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("log2n", [18, 24])
def test_sharded_paths_against_the_oracle(log2n):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = min(n, 8)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py"), str(log2n), "--no-timing"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    tail = (res.stdout + res.stderr)[-4000:]
    assert res.returncode == 0 and "ALL OK" in res.stdout, tail
