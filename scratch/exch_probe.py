"""2-GPU timing of the exchange step pieces (run under torchrun)."""
import os, sys, time
from pathlib import Path
import numpy as np
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
import ggcat_b200 as G
from ggcat_b200 import dist as gdist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
if os.environ.get("P2PCH"):
    os.environ["NCCL_MIN_P2P_NCHANNELS"] = os.environ["P2PCH"]; os.environ["NCCL_MAX_P2P_NCHANNELS"] = "32"
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n_reads = bench.READS_PER_GPU
data, offsets = bench.make_reads(rank, world, n_reads)
b1, b2 = G.bucket_counts(int(n_reads * world * (bench.READ_LEN + 15)))
ctx = G.GGCATB200(G.Params(k=bench.K, m=bench.M, min_multiplicity=bench.S, buckets_count_log=b1, second_buckets_count_log=b2, device=lr))
d_data = torch.from_numpy(data).cuda(); d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
owner = gdist.OwnerMap(b1, b2, world)
def sync():
    torch.cuda.synchronize(); ctx.synchronize()
for it in range(6):
    dist.barrier(); sync()
    t0 = time.perf_counter()
    ctx.reset(); ctx.push_reads_device(d_data.data_ptr(), d_off.data_ptr(), n_reads, int(data.size)); ctx.finish_bucketing(); sync()
    t1 = time.perf_counter()
    gdist.exchange_and_import(ctx, owner, rank, world); sync()
    t2 = time.perf_counter()
    fb, cnt = owner.bucket_range(rank)
    ctx.merge_bucket_range_device(fb, cnt); sync()
    t3 = time.perf_counter()
    if rank == 0 and it >= 2:
        print(f"bucketing {1e3*(t1-t0):.2f} exchange {1e3*(t2-t1):.2f} merge {1e3*(t3-t2):.2f} ms", flush=True)
ctx.close(); dist.destroy_process_group()
