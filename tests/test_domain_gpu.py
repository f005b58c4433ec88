"""Spatially sharded E+F (xequinet_b200/domain.py) on 2 GPUs over NCCL against the single-GPU result and the
fp64 oracle.  Needs >= 2 GPUs (skipped on single-GPU boxes; the host logic is covered by test_domain_gloo.py)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_energy_forces_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", str(ROOT / "tests" / "domain_gpu_worker.py"), "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "SHARDED OK" in r.stdout and "SHARDED GRAPH OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
