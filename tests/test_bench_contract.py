"""bench.py output contract, checked on the CPU-only reference arm (no GPU needed): exactly ONE JSON line on
stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.timeout(300)
@pytest.mark.parametrize("workload", ["gls_c2", "sl"])
def test_reference_arm_prints_one_json_line(workload):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", workload, "--no-configs"], capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    from oracle import refload
    # the unmodified reference files when they are loadable (here: /root/reference or the staged oracle/_ref), else the port
    want_kind = "reference" if refload.available() else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["config"]["workload"] and d["config"]["l2"] and d["config"]["n_gpus"] == 1
