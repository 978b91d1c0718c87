"""Minimal driver for ncu captures: a few C2 steps (phase 1 + phase 2) with inputs resident in HBM.
Usage (under gpurun):  ncu ... python profiles/prof_driver.py [steps] [reads]
The last step is bracketed by cudaProfilerStart/Stop, so `ncu --profile-from-start off` captures exactly one warm step
(bench.py's traffic pass does that)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import ggcat_b200 as G  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else bench.READS_PER_GPU
data, offsets = bench.make_reads(0, 1, n_reads)
b1, b2 = G.bucket_counts(int(bench.READS_PER_GPU * (bench.READ_LEN + 15)))
ctx = G.GGCATB200(G.Params(k=bench.K, m=bench.M, min_multiplicity=bench.S, buckets_count_log=b1, second_buckets_count_log=b2))
d_data = torch.from_numpy(data).cuda()
d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
torch.cuda.synchronize()
for i in range(steps):
    if i == steps - 1:
        torch.cuda.profiler.start()   # ncu --profile-from-start off: only the last (warm) step is captured
    ctx.reset()
    ctx.push_reads_device(d_data.data_ptr(), d_off.data_ptr(), n_reads, int(data.size))
    st = ctx.finish_bucketing()
    res = ctx.merge_bucket_range_device(0, (1 << b1) + 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("steps", steps, "superkmers", st.n_superkmers, "kmers", st.n_kmers, "kept/unique/total", res)
ctx.close()
