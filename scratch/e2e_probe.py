"""Times the host-buffer path (push_reads + merge_bucket_range) under different batch / part settings."""
import os, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
import ggcat_b200 as G

n_reads = bench.READS_PER_GPU
data, offsets = bench.make_reads(0, 1, n_reads)
b1, b2 = G.bucket_counts(int(bench.READS_PER_GPU * (bench.READ_LEN + 15)))
h_data = torch.from_numpy(data).pin_memory()
h_off = torch.from_numpy(offsets.view(np.int64)).pin_memory()
for hb, pk in [(1 << 30, 1 << 40), (40 << 20, 1 << 40), (1 << 30, 36 << 20), (40 << 20, 36 << 20), (20 << 20, 18 << 20), (80 << 20, 72 << 20)]:
    os.environ["GGCAT_B200_HOST_BATCH"] = str(hb)
    os.environ["GGCAT_B200_PART_KMERS"] = str(pk)
    ctx = G.GGCATB200(G.Params(k=bench.K, m=bench.M, min_multiplicity=bench.S, buckets_count_log=b1, second_buckets_count_log=b2))
    ts = []
    for i in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.reset()
        ctx.push_reads_ptr(h_data.data_ptr(), h_off.data_ptr(), n_reads)
        t1 = time.perf_counter()
        ctx.finish_bucketing()
        t2 = time.perf_counter()
        tab = ctx.merge_bucket_range(0, (1 << b1) + 1, copy=False)
        t3 = time.perf_counter()
        tab.release()
        ts.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
    a = np.array(ts[3:]).mean(0) * 1e3
    print(f"host_batch={hb>>20}M part_kmers={pk>>20}M chunks={ctx.n_chunks()}: push {a[0]:.2f} finish {a[1]:.2f} merge {a[2]:.2f} total {a[3]:.2f} ms", flush=True)
    ctx.close()
