"""Multi-GPU test of the fused epilogue + all-gather over NVLink peer memory (needs >= 2 GPUs; skipped otherwise).

Spawns tools/p2p_check.py under torchrun: every rank must end with a periodogram bit-identical to the
NCCL all-gather path and the same global peak.
"""
import ast
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.timeout(600)
def test_fused_p2p_gather_matches_nccl_path():
    n = min(_ngpu(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29597", os.path.join(ROOT, "tools", "p2p_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = ast.literal_eval(line)
    assert res["world"] == n and res["parity_all_ranks"] is True
