"""Multi-GPU parity on real devices: torchrun over NCCL, every rank checks the all-gathered event and
segment tables against the oracle run on the whole trace (straddling events at every chunk boundary)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_pipeline_matches_oracle(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "dist_worker.py"), "24"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
