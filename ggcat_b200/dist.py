"""Multi-GPU plumbing: bucket -> owner map and the one exchange step of the path.

The path shards by first-level minimizer bucket: every k-mer occurrence of a
(k-1)-mer group lands in one bucket that depends only on the minimizer value
(crates/hashes/src/cn_nthash.rs:135-142), boundary k-mers are deliberately stored on both sides
(crates/assembler_minimizer_bucketing/src/lib.rs:252-265), so merging needs no cross-bucket traffic.
The only exchange is the reference's "write bucket files / read bucket files" shuffle
(crates/minimizer_bucketing/src/lib.rs:340-351 -> crates/kmers_transform/src/lib.rs:294-371),
done here as one all-to-all of unit-sorted bucket chunks over NCCL (gloo on CPU in the tests).

A chunk is five flat arrays (see include/ggcat_b200.h, ggcat_b200_chunk_slice): 16-byte descriptors,
payload words, and per-unit super-k-mer / word / k-mer counts.  Because chunks are unit-sorted and
owners hold contiguous unit ranges, every array is already the concatenation of the per-destination
slices in rank order: the all-to-all needs no packing pass.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


class OwnerMap:
    """Contiguous bucket ranges per rank; the duplicates bucket (index 1 << b1) goes to the last rank."""

    def __init__(self, buckets_count_log: int, second_buckets_count_log: int, world: int):
        self.b1, self.b2, self.world = buckets_count_log, second_buckets_count_log, world
        nb = 1 << buckets_count_log
        if world > nb:
            raise ValueError(f"{world} ranks for {nb} buckets")
        # owner(b) = b * world >> b1  <=>  rank r owns [ceil(r*nb/world), ceil((r+1)*nb/world))
        self.starts = [-(-r * nb // world) for r in range(world)] + [nb + 1]
        self.starts[world] = nb + 1  # last rank also owns the duplicates bucket

    def bucket_range(self, rank: int) -> tuple[int, int]:
        return self.starts[rank], self.starts[rank + 1] - self.starts[rank]

    def unit_range(self, rank: int) -> tuple[int, int]:
        fb, cnt = self.bucket_range(rank)
        return fb << self.b2, cnt << self.b2

    def owner_of_bucket(self, bucket: int) -> int:
        if bucket >= (1 << self.b1):
            return self.world - 1
        return (bucket * self.world) >> self.b1


@dataclass
class ChunkArrays:
    """One bucket chunk as torch tensors (CPU or CUDA): the unit of exchange."""
    desc: torch.Tensor        # uint8 [n_sk * 16]
    payload: torch.Tensor     # int32 [n_words]
    unit_cnt: torch.Tensor    # int32 [n_units]
    unit_words: torch.Tensor  # int32 [n_units]
    unit_kmers: torch.Tensor  # int32 [n_units]


@dataclass
class ReceivedSlice:
    src: int
    n_sk: int
    n_words: int
    word_bias: int
    desc: torch.Tensor
    payload: torch.Tensor
    unit_cnt: torch.Tensor
    unit_words: torch.Tensor
    unit_kmers: torch.Tensor


def plan_splits(unit_cnt: np.ndarray, unit_words: np.ndarray, owner: OwnerMap):
    """Per-destination (n_sk, n_words, word_bias) from the per-unit counts of a whole chunk."""
    cs = np.concatenate([[0], np.cumsum(unit_cnt.astype(np.int64))])
    ws = np.concatenate([[0], np.cumsum(unit_words.astype(np.int64))])
    out = np.zeros((owner.world, 3), np.int64)
    for r in range(owner.world):
        fu, nu = owner.unit_range(r)
        out[r] = (cs[fu + nu] - cs[fu], ws[fu + nu] - ws[fu], ws[fu])
    return out


def exchange_chunk(chunk: ChunkArrays, owner: OwnerMap, rank: int, group=None) -> list[ReceivedSlice]:
    """All-to-all of one chunk.  Returns the slices this rank owns, one per source rank."""
    world = owner.world
    dev = chunk.desc.device
    plan = plan_splits(chunk.unit_cnt.cpu().numpy(), chunk.unit_words.cpu().numpy(), owner)
    send_meta = torch.from_numpy(plan).to(dev)
    recv_meta = torch.empty_like(send_meta)
    dist.all_to_all_single(recv_meta, send_meta, group=group)
    rm = recv_meta.cpu().numpy()
    my_fu, my_nu = owner.unit_range(rank)
    unit_splits_in = [owner.unit_range(r)[1] for r in range(world)]

    def a2a(t: torch.Tensor, in_splits, out_splits, pad: int = 0):
        n_out = int(sum(out_splits))
        buf = torch.zeros(n_out + pad, dtype=t.dtype, device=dev)  # pad: the merge kernel reads 2 words past a payload
        dist.all_to_all_single(buf[:n_out], t[: int(sum(in_splits))], output_split_sizes=[int(x) for x in out_splits],
                               input_split_sizes=[int(x) for x in in_splits], group=group)
        return buf

    desc = a2a(chunk.desc, plan[:, 0] * 16, rm[:, 0] * 16)
    payload = a2a(chunk.payload, plan[:, 1], rm[:, 1], pad=8)
    ucnt = a2a(chunk.unit_cnt, unit_splits_in, [my_nu] * world)
    uwords = a2a(chunk.unit_words, unit_splits_in, [my_nu] * world)
    ukmers = a2a(chunk.unit_kmers, unit_splits_in, [my_nu] * world)
    out = []
    d0 = w0 = 0
    for src in range(world):
        n_sk, n_words, bias = (int(x) for x in rm[src])
        out.append(ReceivedSlice(src, n_sk, n_words, bias, desc[d0 * 16:(d0 + n_sk) * 16], payload[w0:w0 + n_words],
                                 ucnt[src * my_nu:(src + 1) * my_nu], uwords[src * my_nu:(src + 1) * my_nu],
                                 ukmers[src * my_nu:(src + 1) * my_nu]))
        d0 += n_sk
        w0 += n_words
    return out


class _DevArray:
    """Zero-copy torch view of library-owned device memory via __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int, typestr: str, itemsize: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2}


def _view(ptr: int, n: int, dtype: torch.dtype, device) -> torch.Tensor:
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    ts, isz = {torch.uint8: ("|u1", 1), torch.int32: ("<i4", 4)}[dtype]
    return torch.as_tensor(_DevArray(ptr, n * isz, ts, isz), device=device)


def exchange_and_import(ctx, owner: OwnerMap, rank: int, world: int, stream: Optional[torch.cuda.Stream] = None, group=None):
    """GPU path: route every local chunk of `ctx` to the bucket owners and register what arrives.
    Must be called after ctx.finish_bucketing() (which synchronises the library stream)."""
    from . import _lib  # noqa: F401

    dev = torch.device("cuda", ctx.params.device)
    n_units_total = ((1 << owner.b1) + 1) << owner.b2
    received = []
    for c in range(ctx.n_chunks()):
        s = ctx.export_chunk_slice(c, 0, n_units_total)
        chunk = ChunkArrays(
            desc=_view(s.d_descriptors, int(s.n_superkmers) * 16, torch.uint8, dev),
            payload=_view(s.d_payload, int(s.n_words), torch.int32, dev),
            unit_cnt=_view(s.d_unit_counts, n_units_total, torch.int32, dev),
            unit_words=_view(s.d_unit_words, n_units_total, torch.int32, dev),
            unit_kmers=_view(s.d_unit_kmers, n_units_total, torch.int32, dev),
        )
        received += exchange_chunk(chunk, owner, rank, group)
    torch.cuda.synchronize(dev)  # collectives done before the library stream consumes / recycles buffers
    ctx.drop_local_chunks()
    my_fu, my_nu = owner.unit_range(rank)
    for r in received:
        if r.n_sk == 0:
            continue
        sl = _lib.ChunkSliceC(n_superkmers=r.n_sk, n_words=r.n_words, word_bias=r.word_bias,
                              d_descriptors=r.desc.data_ptr(), d_payload=r.payload.data_ptr(),
                              d_unit_counts=r.unit_cnt.data_ptr(), d_unit_words=r.unit_words.data_ptr(),
                              d_unit_kmers=r.unit_kmers.data_ptr())
        ctx.import_chunk_slice(my_fu, my_nu, sl, keepalive=r)
    return received
