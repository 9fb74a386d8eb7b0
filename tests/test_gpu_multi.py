"""Sharded (i_dw columns) operator / Lanczos / GF chains on >= 2 GPUs, one process per GPU over NCCL.
Skipped on a single-GPU box; the 1-GPU tests cover everything except the exchange step."""
import os
import subprocess
import sys

import pytest

import edgpu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nproc", [2, 3])
def test_sharded_parity_under_torchrun(nproc):
    if edgpu.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + nproc), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
