"""Multi-GPU (one process per GPU, NCCL) parity; skipped on boxes with a single GPU."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_search_and_peer_gather_over_nccl(libmrag):
    n = torch.cuda.device_count()          # 2, 4 or 8: one rank per GPU of the box
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "tests" / "mgpu_worker.py")],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and f"MGPU_OK {n}" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
