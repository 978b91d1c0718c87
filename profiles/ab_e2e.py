"""A/B driver for the host path (e2e): C2 through ggcat_b200_push_reads + ggcat_b200_merge_bucket_range with pinned host
buffers, one subprocess per environment setting; prints the median e2e ms and a digest of the host table.
Usage (under gpurun): python profiles/ab_e2e.py "GGCAT_B200_PART_SCHEME=0" "GGCAT_B200_PART_SCHEME=1 GGCAT_B200_PART_KMERS=25165824" ..."""
import hashlib
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def one():
    import numpy as np
    import torch

    import bench
    import ggcat_b200 as G

    n_reads = bench.READS_PER_GPU
    data, offsets = bench.make_reads(0, 1, n_reads)
    b1, b2 = G.bucket_counts(int(n_reads * (bench.READ_LEN + 15)))
    h_data = torch.from_numpy(data).pin_memory()
    h_off = torch.from_numpy(offsets.view(np.int64)).pin_memory()
    ctx = G.GGCATB200(G.Params(k=bench.K, m=bench.M, min_multiplicity=bench.S, buckets_count_log=b1, second_buckets_count_log=b2))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts, parts = [], []
    digest = None
    for i in range(14):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.reset()
        ctx.push_reads_ptr(h_data.data_ptr(), h_off.data_ptr(), n_reads)
        t1 = time.perf_counter()
        ctx.finish_bucketing()
        tab = ctx.merge_bucket_range(0, (1 << b1) + 1, copy=False)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if i >= 4:
            ts.append(1e3 * (t2 - t0)); parts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
        if i == 13:
            h = hashlib.sha256()
            h.update(tab.keys_lo.tobytes()); h.update(tab.count_flags.tobytes()); h.update(tab.unit_offsets.tobytes())
            digest = h.hexdigest()[:16]
        tab.release()
    ts.sort()
    print(json.dumps({"env": os.environ.get("AB_LABEL", ""), "e2e_ms_median": round(ts[len(ts) // 2], 3), "e2e_ms_min": round(ts[0], 3),
                      "push_ms": round(sorted(p[0] for p in parts)[len(parts) // 2], 3),
                      "merge_ms": round(sorted(p[1] for p in parts)[len(parts) // 2], 3), "digest": digest}), flush=True)
    ctx.close()


if __name__ == "__main__":
    if sys.argv[1:] == ["--one"]:
        one()
    else:
        for setting in sys.argv[1:]:
            env = dict(os.environ)
            env["AB_LABEL"] = setting
            for kv in setting.split():
                if "=" in kv:
                    k, v = kv.split("=", 1)
                    env[k] = v
            r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True, timeout=300)
            out = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(out[-1] if out else json.dumps({"env": setting, "error": (r.stderr or r.stdout)[-400:]}), flush=True)
