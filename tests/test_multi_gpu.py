"""Multi-GPU (NCCL) check of the sharded paths; needs >= 2 GPUs on the box, skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_matches_single_gpu():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    err = "\n".join(l for l in out.stderr.splitlines() if "Error" in l or "assert" in l or "File \"/root" in l or "Traceback" in l)
    assert out.returncode == 0 and "MGPU_OK 2" in out.stdout, out.stdout[-2000:] + err[-6000:]
