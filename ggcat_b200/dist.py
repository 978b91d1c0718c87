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
    # host copies of the three per-unit arrays (the library keeps them after finish_bucketing); when absent they
    # are read back from the tensors
    h_unit_cnt: Optional[np.ndarray] = None
    h_unit_words: Optional[np.ndarray] = None
    h_unit_kmers: Optional[np.ndarray] = None


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
    h_unit_cnt: Optional[np.ndarray] = None   # host copies (int32) of the three arrays above
    h_unit_words: Optional[np.ndarray] = None
    h_unit_kmers: Optional[np.ndarray] = None


def plan_splits(unit_cnt: np.ndarray, unit_words: np.ndarray, owner: OwnerMap):
    """Per-destination (n_sk, n_words, word_bias) from the per-unit counts of a whole chunk."""
    out = np.zeros((owner.world, 3), np.int64)
    bias = 0
    for r in range(owner.world):  # owners hold contiguous, ascending unit ranges
        fu, nu = owner.unit_range(r)
        w = int(unit_words[fu:fu + nu].sum(dtype=np.int64))
        out[r] = (int(unit_cnt[fu:fu + nu].sum(dtype=np.int64)), w, bias)
        bias += w
    return out


_PINNED: dict = {}
_META_TAIL = 4  # n_sk, n_words, word_bias (lo, hi) appended to the per-unit arrays of every destination


def exchange_chunk(chunk: ChunkArrays, owner: OwnerMap, rank: int, group=None) -> list[ReceivedSlice]:
    """All-to-all of one chunk: three collectives (metadata, descriptors, payload) and one host read-back.
    Returns the slices this rank owns, one per source rank.

    Metadata message to destination r: [unit_cnt | unit_words | unit_kmers] of r's units + (n_sk, n_words, bias);
    it is planned on the host from the host copies of the per-unit arrays, so the sender never reads the device."""
    import os, time
    _tr = os.environ.get("GGCAT_B200_TRACE") == "1"
    _t = [time.perf_counter()]

    def _mark():
        if _tr:
            if chunk.desc.device.type == "cuda":
                torch.cuda.synchronize()
            _t.append(time.perf_counter())

    world = owner.world
    dev = chunk.desc.device
    h_cnt = chunk.h_unit_cnt if chunk.h_unit_cnt is not None else chunk.unit_cnt.cpu().numpy()
    h_words = chunk.h_unit_words if chunk.h_unit_words is not None else chunk.unit_words.cpu().numpy()
    h_kmers = chunk.h_unit_kmers if chunk.h_unit_kmers is not None else chunk.unit_kmers.cpu().numpy()
    plan = plan_splits(h_cnt, h_words, owner)
    my_fu, my_nu = owner.unit_range(rank)
    msgs = []
    for r in range(world):
        fu, nu = owner.unit_range(r)
        tail = np.array([plan[r, 0], plan[r, 1], plan[r, 2] & 0xFFFFFFFF, plan[r, 2] >> 32], np.int64).astype(np.uint32).view(np.int32)
        msgs.append(np.concatenate([np.asarray(h_cnt[fu:fu + nu]).view(np.int32), np.asarray(h_words[fu:fu + nu]).view(np.int32),
                                    np.asarray(h_kmers[fu:fu + nu]).view(np.int32), tail]))
    meta_np = np.concatenate(msgs)
    if dev.type == "cuda":
        # staged through a cached pinned buffer (the read-back below completes the copy before the next call reuses it)
        pin = _PINNED.get(meta_np.size)
        if pin is None:
            pin = _PINNED[meta_np.size] = torch.empty(meta_np.size, dtype=torch.int32).pin_memory()
        pin.numpy()[:] = meta_np
        send_meta = pin.to(dev, non_blocking=True)
    else:
        send_meta = torch.from_numpy(meta_np)
    _mark()
    in_meta = [3 * owner.unit_range(r)[1] + _META_TAIL for r in range(world)]
    per_src = 3 * my_nu + _META_TAIL
    recv_meta = torch.empty(per_src * world, dtype=torch.int32, device=dev)
    dist.all_to_all_single(recv_meta, send_meta, output_split_sizes=[per_src] * world, input_split_sizes=in_meta, group=group)
    rm = recv_meta.cpu().numpy().reshape(world, per_src)   # the one synchronisation of the exchange
    _mark()
    tails = rm[:, 3 * my_nu:].view(np.uint32).astype(np.int64)
    r_sk, r_words, r_bias = tails[:, 0], tails[:, 1], tails[:, 2] | (tails[:, 3] << 32)

    def a2a(t: torch.Tensor, in_splits, out_splits, pad: int = 0):
        n_out = int(sum(out_splits))
        buf = torch.empty(n_out + pad, dtype=t.dtype, device=dev)
        if pad:
            buf[n_out:].zero_()  # the merge kernel reads up to 2 words past a payload
        dist.all_to_all_single(buf[:n_out], t[: int(sum(in_splits))], output_split_sizes=[int(x) for x in out_splits],
                               input_split_sizes=[int(x) for x in in_splits], group=group)
        return buf

    desc = a2a(chunk.desc, plan[:, 0] * 16, r_sk * 16)
    _mark()
    payload = a2a(chunk.payload, plan[:, 1], r_words, pad=8)
    _mark()
    if _tr and rank == 0:
        print("exchange_chunk ms: plan+meta-build %.3f meta-a2a %.3f desc-a2a %.3f (%d B) payload-a2a %.3f (%d B)" % (
            1e3 * (_t[1] - _t[0]), 1e3 * (_t[2] - _t[1]), 1e3 * (_t[3] - _t[2]), desc.numel(), 1e3 * (_t[4] - _t[3]), payload.numel() * 4), flush=True)
    rmd = recv_meta.view(world, per_src)
    out = []
    d0 = w0 = 0
    for src in range(world):
        n_sk, n_words, bias = int(r_sk[src]), int(r_words[src]), int(r_bias[src])
        out.append(ReceivedSlice(src, n_sk, n_words, bias, desc[d0 * 16:(d0 + n_sk) * 16], payload[w0:w0 + n_words],
                                 rmd[src, 0:my_nu], rmd[src, my_nu:2 * my_nu], rmd[src, 2 * my_nu:3 * my_nu],
                                 np.ascontiguousarray(rm[src, 0:my_nu]), np.ascontiguousarray(rm[src, my_nu:2 * my_nu]),
                                 np.ascontiguousarray(rm[src, 2 * my_nu:3 * my_nu])))
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


def _host_u32(ptr: int, n: int) -> np.ndarray:
    import ctypes as C

    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(n,))


def peer_setup(ctx, rank: int, world: int, arena_bytes: int, group=None):
    """Collective, once per context: allocate the NVLink receive arena (ggcat_b200_peer_init), all-gather the
    64-byte CUDA IPC handles with torch.distributed and map every peer's arena (ggcat_b200_peer_connect).
    Afterwards exchange_and_import() is one library call (copy kernel over peer memory, no NCCL)."""
    handle = ctx.peer_init(rank, world, arena_bytes)
    if world == 1:
        ctx.peer_connect([handle])
        return
    dev = torch.device("cuda", ctx.params.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
    allh = torch.empty(64 * world, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, mine, group=group)
    raw = allh.cpu().numpy().tobytes()
    ctx.peer_connect([raw[64 * r:64 * (r + 1)] for r in range(world)])
    dist.barrier(group=group)  # every arena is mapped everywhere before the first push


def exchange_and_import(ctx, owner: OwnerMap, rank: int, world: int, stream: Optional[torch.cuda.Stream] = None, group=None):
    """GPU path: route every local chunk of `ctx` to the bucket owners and register what arrives.
    Must be called after ctx.finish_bucketing() (which synchronises the library stream).
    With a connected peer arena (peer_setup) this is ggcat_b200_peer_exchange; otherwise the NCCL all-to-all."""
    from . import _lib  # noqa: F401

    if getattr(ctx, "peer_connected", False):
        ctx.peer_exchange()
        return []

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
            h_unit_cnt=_host_u32(s.h_unit_counts, n_units_total), h_unit_words=_host_u32(s.h_unit_words, n_units_total),
            h_unit_kmers=_host_u32(s.h_unit_kmers, n_units_total),
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
                              d_unit_kmers=r.unit_kmers.data_ptr(),
                              h_unit_counts=r.h_unit_cnt.ctypes.data, h_unit_words=r.h_unit_words.ctypes.data,
                              h_unit_kmers=r.h_unit_kmers.ctypes.data)
        ctx.import_chunk_slice(my_fu, my_nu, sl, keepalive=r)
    return received
