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


def test_sampled_superkmers_keep_the_units_merge_unit_needs():
    """bench.multi_gpu_parity pre-filters the oracle's super-k-mers to the sampled units: merge_unit must see the same rows."""
    import numpy as np

    import bench
    from oracle import oracle as O

    data, offsets = bench.make_reads(0, 2, 3000)
    d2, _ = bench.make_reads(1, 2, 3000)          # second slice from the cached genome
    assert data.size == d2.size and not np.array_equal(data, d2)
    reads = O.Reads(data, offsets)
    b1, b2 = 3, 2
    sk, _ = O.bucketing(reads, bench.K, bench.M, b1, b2)
    units = [0, 5, 17, (8 << b2) | 1, 33]
    small = bench.sampled_superkmers(sk, units, b2)
    assert 0 < small.size < sk.size
    for u in units:
        a, _, ta = O.merge_unit(reads, sk, u >> b2, u & ((1 << b2) - 1), bench.K, bench.S)
        b, _, tb = O.merge_unit(reads, small, u >> b2, u & ((1 << b2) - 1), bench.K, bench.S)
        assert ta == tb and np.array_equal(a, b)


def test_parity_against_oracle_accepts_oracle_tables_and_flags_a_corrupted_one():
    """The host half of the bench's N > 1 parity object (rank 0: oracle on the union of all ranks' reads), fed with unit
    tables made by the oracle itself: everything matches; one flipped count is reported."""
    import numpy as np

    import bench
    from oracle import oracle as O

    world, n_reads, b1, b2 = 2, 2500, 3, 2
    parts = bench.make_reads_all(world, n_reads)
    data = np.concatenate([d for d, _ in parts])
    reads = O.Reads(data, np.arange(data.size // bench.READ_LEN + 1, dtype=np.uint64) * np.uint64(bench.READ_LEN))
    sk, _ = O.bucketing(reads, bench.K, bench.M, b1, b2)

    def table(u):
        ref, _, _ = O.merge_unit(reads, sk, u >> b2, u & ((1 << b2) - 1), bench.K, bench.S)
        ref = ref[ref["kept"] == 1]
        cf = (ref["multiplicity"].astype(np.uint32) & np.uint32(0x3FFFFFFF)) | (ref["flags"].astype(np.uint32) << np.uint32(30))
        return np.array(ref["key_lo"]), cf

    gathered = []
    for r, units in enumerate([[1, 6, 9], [17, 20, 30]]):
        rows = [table(u) for u in units]
        gathered.append({"rank": r, "units": units, "tables": {"device": rows, "host": [(k.copy(), c.copy()) for k, c in rows]}})
    res = bench.parity_against_oracle(gathered, world, n_reads, b1, b2)
    assert res["ok"] and res["units_checked"] == 12 and res["paths"] == ["device", "host"] and res["entries_checked"] > 0
    victim = next(i for i, (k, c) in enumerate(gathered[1]["tables"]["host"]) if c.size)
    gathered[1]["tables"]["host"][victim][1][0] ^= np.uint32(1)
    res = bench.parity_against_oracle(gathered, world, n_reads, b1, b2)
    assert not res["ok"] and res["mismatches"] == [{"rank": 1, "unit": gathered[1]["units"][victim], "path": "host"}]
