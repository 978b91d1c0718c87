"""Kernel timeline of one warm bench step (device path), N ranks: GGCAT_B200_TRACE prints every timed launch.
   torchrun ... profiles/trace_step.py   (or plain python for N=1)"""
import os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench, ggcat_b200 as G
from ggcat_b200 import dist as gdist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n_reads = bench.READS_PER_GPU
data, offsets = bench.make_reads(rank, world, n_reads)
b1, b2 = G.bucket_counts(int(n_reads * world * (bench.READ_LEN + 15)))
ctx = G.GGCATB200(G.Params(k=bench.K, m=bench.M, min_multiplicity=bench.S, buckets_count_log=b1, second_buckets_count_log=b2, device=lr))
d_data = torch.from_numpy(data).cuda(); d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
owner = gdist.OwnerMap(b1, b2, world)
if world > 1:
    gdist.peer_setup(ctx, rank, world, arena_bytes=max(6 * int(data.size), 64 << 20))
npush = int(os.environ.get("NPUSH", "1" if world == 1 else "2"))
per = (n_reads + npush - 1) // npush
pushes = []
for r0 in range(0, n_reads, per):
    r1 = min(n_reads, r0 + per)
    off = (d_off[r0:r1 + 1] - r0 * 150).contiguous()
    pushes.append((d_data.data_ptr() + r0 * 150, off, r1 - r0, (r1 - r0) * 150))
fb, cnt = owner.bucket_range(rank)
def step():
    ctx.reset()
    for ptr, off, nr, nb in pushes:
        ctx.push_reads_device(ptr, off.data_ptr(), nr, nb)
    ctx.finish_bucketing()
    if world > 1:
        ctx.peer_exchange()
    return ctx.merge_bucket_range_device(fb, cnt)
for _ in range(4):
    step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
ctx.set_timing(True); ctx.kernel_times(reset=True)
if rank == 0 and os.environ.get("HOST_TRACE"):
    os.environ["GGCAT_B200_TRACE"] = "host"
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); t1 = time.perf_counter()
os.environ.pop("GGCAT_B200_TRACE", None)
if rank == 0:
    os.environ["GGCAT_B200_TRACE"] = "1"
kt = ctx.kernel_times(reset=True)
if rank == 0:
    print("host wall of the step: %.3f ms" % (1e3 * (t1 - t0)), flush=True)
if world > 1: dist.barrier()
ctx.close()
