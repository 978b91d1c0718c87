"""Exploration: device-timed passes of the other BASELINE configs (not a bench line).
    python scratch/workloads.py c4 20000000      # C4-shaped slice: R error-free reads of a 5R bp genome, k=31, 1024 buckets
    python scratch/workloads.py c5 10000000      # k=63 m=14 rabin-karp128
    python scratch/workloads.py c3 20            # coloured, G genomes x 5 Mbp
"""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ggcat_b200 as G
from ggcat_b200 import synth

wl = sys.argv[1]
R = int(sys.argv[2])
dev = torch.device("cuda", 0)
L = 150
colors = None
if wl in ("c4", "c5", "c2"):
    err = {"c4": 0.0, "c5": 0.0, "c2": 0.01}[wl]
    k, m, s = (63, 14, 2) if wl == "c5" else (31, 12, 2)
    ht = G.api.HASH_RK128 if wl == "c5" else G.api.HASH_AUTO
    t0 = time.time()
    genome = synth.genome_codes_torch(0xC4, 5 * R, dev)
    data = synth.simulate_reads_torch(genome, R, L, err, 0xC5, 0)
    del genome
    torch.cuda.synchronize()
    print(f"generated {R} reads in {time.time()-t0:.1f}s", flush=True)
    n_rec = R
    offsets = torch.arange(R + 1, dtype=torch.int64, device=dev) * L
    b1, b2 = G.bucket_counts(int(R * (L + 15)))
else:
    k, m, s, ht = 31, 12, 1, G.api.HASH_AUTO
    d, o, c = synth.config_c3(n_genomes=R)
    data = torch.from_numpy(d).to(dev); offsets = torch.from_numpy(o.view(np.int64)).to(dev)
    colors = torch.from_numpy(c.view(np.int32)).to(dev)
    n_rec = R
    b1, b2 = G.bucket_counts(int(d.size))
print(f"workload {wl}: {data.numel()/1e9:.3f} Gbases, k={k} m={m} s={s} buckets {1<<b1}x{1<<b2}", flush=True)
ctx = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2, hash_type=ht,
                           colors=colors is not None))
ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
BATCH = 4_000_000 if colors is None else 40


def step():
    ctx.reset()
    for r0 in range(0, n_rec, BATCH):
        r1 = min(n_rec, r0 + BATCH)
        if colors is None:
            off = offsets[r0:r1 + 1] - r0 * L
            ctx.push_reads_device(data.data_ptr() + r0 * L, off.data_ptr(), r1 - r0, (r1 - r0) * L)
        else:
            b0 = int(offsets[r0]); b1_ = int(offsets[r1])
            off = offsets[r0:r1 + 1] - b0
            ctx.push_reads_device(data.data_ptr() + b0, off.data_ptr(), r1 - r0, b1_ - b0, colors.data_ptr() + 4 * r0)
    st = ctx.finish_bucketing()
    res = ctx.merge_bucket_range_device(0, (1 << b1) + 1)
    return st, res


for it in range(3):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(ext)
    st, res = step()
    b.record(ext)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = a.elapsed_time(b)
    print(f"iter {it}: device {ms:.2f} ms ({data.numel()/ms/1e6:.2f} Gbases/s), wall {wall*1e3:.1f} ms; sk={st.n_superkmers} kmers={st.n_kmers} "
          f"kept={res[0]} unique={res[1]} total={res[2]}; mem {torch.cuda.memory_allocated()/1e9:.1f} GB torch", flush=True)
ctx.set_timing(True)
ctx.kernel_times(reset=True)
step()
kt = ctx.kernel_times(reset=True)
print({k_: (round(v[0], 3), v[1]) for k_, v in kt.items() if v[1]}, flush=True)
free, total = torch.cuda.mem_get_info()
print(f"device memory in use {(total-free)/1e9:.1f} GB")
