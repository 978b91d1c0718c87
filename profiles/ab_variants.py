"""A/B driver for kernel variants: runs the C2 step with every library given (builds of the same sources with different
-D switches, e.g. ggcat_b200/variants/*.so), one subprocess each, and prints per-family kernel times, the step time and a
digest of the resulting table (all variants must agree).
Usage (under gpurun): python profiles/ab_variants.py lib1.so lib2.so ...   |   python profiles/ab_variants.py --one"""
import hashlib
import json
import os
import subprocess
import sys
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
    ctx = G.GGCATB200(G.Params(k=bench.K, m=bench.M, min_multiplicity=bench.S, buckets_count_log=b1, second_buckets_count_log=b2))
    d_data = torch.from_numpy(data).cuda()
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ext = torch.cuda.ExternalStream(ctx.stream_ptr)

    def step():
        ctx.reset()
        ctx.push_reads_device(d_data.data_ptr(), d_off.data_ptr(), n_reads, int(data.size))
        st = ctx.finish_bucketing()
        return st, ctx.merge_bucket_range_device(0, (1 << b1) + 1)

    for _ in range(3):
        step()
    ms = []
    for _ in range(10):
        with torch.cuda.stream(ext):
            flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ext)
        st, res = step()
        b.record(ext)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ctx.set_timing(True)
    ctx.kernel_times(reset=True)
    for _ in range(3):
        with torch.cuda.stream(ext):
            flush.fill_(1)
        step()
    kt = {k: round(v[0] / 3, 4) for k, v in ctx.kernel_times(reset=True).items() if v[1]}
    ctx.set_timing(False)
    tab = ctx.read_device_table()
    h = hashlib.sha256()
    h.update(tab.keys_lo.tobytes()); h.update(tab.count_flags.tobytes()); h.update(tab.unit_offsets.tobytes())
    print(json.dumps({"lib": os.environ.get("GGCAT_B200_LIB", "default"), "step_ms_median": round(sorted(ms)[len(ms) // 2], 4),
                      "step_ms_min": round(min(ms), 4), "kernels": kt, "superkmers": int(st.n_superkmers), "entries": int(res[0]),
                      "digest": h.hexdigest()[:16]}), flush=True)
    ctx.close()


if __name__ == "__main__":
    if sys.argv[1:] == ["--one"]:
        one()
    else:
        for lib in sys.argv[1:]:
            env = dict(os.environ)
            if lib != "default":
                env["GGCAT_B200_LIB"] = str(Path(lib).resolve())
            r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True, timeout=300)
            out = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(out[-1] if out else json.dumps({"lib": lib, "error": (r.stderr or r.stdout)[-400:]}), flush=True)
