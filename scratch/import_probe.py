"""Single-GPU reproduction of the owner-side import path: context A buckets, its chunk slices are copied into
torch tensors and imported into context B (as the all-to-all would), B merges; tables must equal A's."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ggcat_b200 as G
from ggcat_b200 import dist as gdist, synth, _lib

k, m, s, b1, b2 = 31, 12, 2, 5, 3
g = synth.genome_codes(0xC2, 100_000)
data, offsets = synth.reads_to_ascii_batch(synth.simulate_reads(g, 20000, 150, 0.01, 0xC3))
A = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2))
half = 10000
A.push_reads(data[: half * 150], offsets[: half + 1])
A.push_reads(data[half * 150:], offsets[half:] - offsets[half])
A.finish_bucketing()
world = 2
owner = gdist.OwnerMap(b1, b2, world)
dev = torch.device("cuda", 0)
n_units_total = ((1 << b1) + 1) << b2
for rank in range(world):
    B = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2))
    fu, nu = owner.unit_range(rank)
    keep = []
    for c in range(A.n_chunks()):
        sl = A.export_chunk_slice(c, fu, nu)
        desc = gdist._view(sl.d_descriptors, int(sl.n_superkmers) * 16, torch.uint8, dev).clone()
        pay = torch.zeros(int(sl.n_words) + 8, dtype=torch.int32, device=dev)
        pay[: int(sl.n_words)] = gdist._view(sl.d_payload, int(sl.n_words), torch.int32, dev)
        uc = gdist._view(sl.d_unit_counts, nu, torch.int32, dev).clone()
        uw = gdist._view(sl.d_unit_words, nu, torch.int32, dev).clone()
        uk = gdist._view(sl.d_unit_kmers, nu, torch.int32, dev).clone()
        torch.cuda.synchronize()
        s2 = _lib.ChunkSliceC(n_superkmers=sl.n_superkmers, n_words=sl.n_words, word_bias=sl.word_bias, d_descriptors=desc.data_ptr(),
                              d_payload=pay.data_ptr(), d_unit_counts=uc.data_ptr(), d_unit_words=uw.data_ptr(), d_unit_kmers=uk.data_ptr())
        B.import_chunk_slice(fu, nu, s2, keepalive=(desc, pay, uc, uw, uk))
    B.finish_bucketing()
    fb, nb = owner.bucket_range(rank)
    tb = B.merge_bucket_range(fb, nb)
    ta = A.merge_bucket_range(fb, nb)
    assert np.array_equal(ta.keys_lo, tb.keys_lo) and np.array_equal(ta.count_flags, tb.count_flags) and np.array_equal(ta.unit_offsets, tb.unit_offsets)
    print("rank", rank, "ok", tb.n_entries)
    B.close()
A.close()
