import os, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench, ggcat_b200 as G
n_reads = bench.READS_PER_GPU
data, offsets = bench.make_reads(0, 1, n_reads)
b1, b2 = G.bucket_counts(int(n_reads * (bench.READ_LEN + 15)))
h_data = torch.from_numpy(data).pin_memory(); h_off = torch.from_numpy(offsets.view(np.int64)).pin_memory()
ctx = G.GGCATB200(G.Params(k=bench.K, m=bench.M, min_multiplicity=bench.S, buckets_count_log=b1, second_buckets_count_log=b2))
for i in range(5):
    if i == 4: os.environ["GGCAT_B200_TRACE"] = "1"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.reset(); ctx.push_reads_ptr(h_data.data_ptr(), h_off.data_ptr(), n_reads); t1 = time.perf_counter()
    ctx.finish_bucketing(); t2 = time.perf_counter()
    tab = ctx.merge_bucket_range(0, (1 << b1) + 1, copy=False); t3 = time.perf_counter(); tab.release()
    print(f"push {1e3*(t1-t0):.2f} finish {1e3*(t2-t1):.2f} merge {1e3*(t3-t2):.2f}")
