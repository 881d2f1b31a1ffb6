"""Multi-GPU DEVICE path against the single-domain oracle: torchrun with one rank per GPU pushes a small
2-D P1 slab problem through the fused (overlapped) NCCL exchange and a small 3-D P2 z-slab problem through the exchange
that follows the general H1 gather; every rank compares its owned columns and b with orc.example200_assemble on the
global mesh (tests/mg_parity.py).  Needs >= 2 GPUs; bench.py runs the same check at every N > 1
(`checks.mg_parity_max_rel`), which is where the driver's 2/4/8-GPU boxes exercise it."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("nranks", [2, 4])
def test_device_exchange_matches_single_domain_oracle(nranks):
    import torch

    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs, found {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "bench.py"), "--gpus", str(nranks), "--parity-only"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    par = json.loads(line)["mg_parity"]
    assert par["max_rel"] <= 1e-12, par
