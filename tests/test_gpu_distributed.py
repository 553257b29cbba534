"""Multi-GPU parity on a real box: row shards + NCCL all-gather + merge kernel must reproduce the CPU oracle's answer for the
whole corpus on every rank (tests/scripts/dist_check.py under torchrun).  Skipped on single-GPU boxes; the CPU suite covers
the same exchange over gloo (tests/test_distributed_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_search_matches_the_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(root, "tests", "scripts", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("bit-exact vs oracle") == 16 and "MISMATCH" not in r.stdout          # 2 ranks x 8 checks
