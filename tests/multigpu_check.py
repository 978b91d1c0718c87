"""N-GPU parity check of the sharded path (run under torchrun on a GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
Every rank buckets its own read slice, the all-to-all routes super-k-mers to bucket owners, owners merge,
and each owner compares its tables with the oracle run on the UNION of all ranks' reads."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ggcat_b200 as G  # noqa: E402
from ggcat_b200 import dist as gdist, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    k, m, s, b1, b2 = 31, 12, 2, 5, 3
    n_reads = 20000
    g = synth.genome_codes(0xC2, 100_000 * world)
    slices = [synth.reads_to_ascii_batch(synth.simulate_reads(g, n_reads, 150, 0.01, 0xC3, first_read=r * n_reads)) for r in range(world)]
    data, offsets = slices[rank]
    ctx = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2, device=lr))
    owner = gdist.OwnerMap(b1, b2, world)
    transport = os.environ.get("GGCAT_B200_EXCHANGE", "peer")
    if transport == "peer":
        gdist.peer_setup(ctx, rank, world, arena_bytes=64 << 20)
    half = offsets.size // 2
    for it in range(3):  # several rounds: reset + buffer recycling with imported chunks, arena reuse (flow control)
        ctx.reset()
        if it == 1:      # two pushes -> two local chunks -> two slices per destination
            ctx.push_reads(data[: int(offsets[half])], offsets[: half + 1])
            ctx.push_reads(data[int(offsets[half]):], offsets[half:] - offsets[half])
        else:
            ctx.push_reads(data, offsets)
        ctx.finish_bucketing()
        gdist.exchange_and_import(ctx, owner, rank, world)
        fb, nb = owner.bucket_range(rank)
        tab = ctx.merge_bucket_range(fb, nb)
    all_data = np.concatenate([d for d, _ in slices])
    all_off = np.arange(world * n_reads + 1, dtype=np.uint64) * np.uint64(150)
    reads = O.Reads(all_data, all_off)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    checked = 0
    for u in range(fb << b2, (fb + nb) << b2):
        ref, _, _ = O.merge_unit(reads, sk, u >> b2, u & ((1 << b2) - 1), k, s)
        ref = ref[ref["kept"] == 1]
        sl = tab.unit_slice(u)
        assert np.array_equal(tab.keys_lo[sl], ref["key_lo"]), f"rank {rank} unit {u} keys"
        assert np.array_equal(tab.multiplicity[sl].astype(np.uint64), ref["multiplicity"]), f"rank {rank} unit {u} counts"
        assert np.array_equal(tab.flags[sl], ref["flags"]), f"rank {rank} unit {u} flags"
        checked += len(ref)
    t = torch.tensor([checked], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"multigpu_check ok: world={world} transport={transport} entries checked={int(t.item())}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
