"""CPU checks of the measurement plumbing: the reference arm of bench.py (oracle port on the host cores) prints the
contract's JSON line, and the device-side synthetic generators (torch) agree with the numpy definitions."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--sample-reads", "20000"], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "Gbases/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("build Gbases/s") and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_torch_generators_match_numpy():
    import torch

    from ggcat_b200 import synth

    g = synth.genome_codes(0xC4, 70001)
    gt = synth.genome_codes_torch(0xC4, 70001, "cpu")
    assert np.array_equal(g, gt.numpy())
    for err in (0.0, 0.01):
        a = synth.reads_to_ascii_batch(synth.simulate_reads(g, 2500, 150, err, 0xC5, first_read=12345))[0]
        b = synth.simulate_reads_torch(gt, 2500, 150, err, 0xC5, first_read=12345)
        assert np.array_equal(a, b.numpy())
