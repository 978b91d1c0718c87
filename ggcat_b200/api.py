"""Host-side mirror of the reference's phase-1 / phase-2 entry points over the C ABI.

Reference interface mirrored (names, argument meaning, error behaviour):
  * ``minimizer_bucketing(input_blocks, buckets_count, second_buckets_count, k, m, forward_only)``
      <- crates/assembler_minimizer_bucketing/src/lib.rs:279-340 (returns the bucket handle the way the
         reference returns ``Vec<MultiChunkBucket>``; here buckets live in HBM inside the context);
  * ``kmers_merge(buckets, min_multiplicity, ...)``
      <- crates/assembler_kmers_merge/src/lib.rs:159-284 up to the completed k-mer table
         (``FxHashMap<hash, MapEntry>``, crates/structs/src/map_entry.rs) per (bucket, second_bucket).
Errors: the reference panics (crates/logging/src/lib.rs:82-107); here every failure raises
``GgcatB200Error`` carrying the library's message.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _lib

HASH_AUTO, HASH_SEQ, HASH_RK128 = 0, 1, 4  # crates/api/src/utils.rs:4-8


class GgcatB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ggcat_b200 error {code}: {msg}")
        self.code = code


def _check(rc: int):
    if rc != 0:
        raise GgcatB200Error(rc, _lib.load().ggcat_b200_last_error().decode())


def compute_best_m(k: int) -> int:
    """crates/utils/src/lib.rs:29-40"""
    return int(_lib.load().ggcat_b200_compute_best_m(k))


def bucket_counts(estimated_bases: int) -> tuple[int, int]:
    """(buckets_count_log, second_buckets_count_log) per crates/io/src/lib.rs:67-140."""
    a, b = C.c_uint32(0), C.c_uint32(0)
    _lib.load().ggcat_b200_bucket_counts(estimated_bases, C.byref(a), C.byref(b))
    return a.value, b.value


@dataclass
class Params:
    k: int
    m: int = 0
    min_multiplicity: int = 2
    buckets_count_log: int = 9
    second_buckets_count_log: int = 6
    forward_only: bool = False
    hash_type: int = HASH_AUTO
    colors: bool = False
    device: int = 0


@dataclass
class BucketStats:
    total_bases: int
    valid_bases: int
    n_superkmers: int
    n_kmers: int
    payload_words: int
    n_buckets: int
    n_units: int


SUPERKMER_DTYPE = np.dtype(
    [("payload_offset", "<u8"), ("len", "<u4"), ("color", "<u4"), ("bucket", "<u2"), ("minimizer_pos", "<u2"),
     ("second_bucket", "u1"), ("flags", "u1"), ("rc", "u1"), ("pad", "u1")]
)


class KmerTable:
    """Filtered k-mer table of a bucket range (host copy).  Arrays are numpy views copied out of the
    library-owned pinned buffers, so the object stays valid after release."""

    def __init__(self, keys_lo, keys_hi, count_flags, first_unit, unit_offsets, total_kmers, unique_kmers,
                 color_offsets=None, colors=None, src_kmers=None):
        self.keys_lo = keys_lo
        self.keys_hi = keys_hi
        self.count_flags = count_flags
        self.first_unit = first_unit
        self.unit_offsets = unit_offsets
        self.total_kmers = total_kmers
        self.unique_kmers = unique_kmers
        self.color_offsets = color_offsets
        self.colors = colors
        # rabin-karp128 only: (n_entries, words) uint64, the bases of one occurrence of every key (base j at bits 2(j % 32) of
        # word j // 32), oriented so that their forward hash is the key -- the reference's saved_reads contract
        self.src_kmers = src_kmers
        self._release = None

    def release(self):
        """Return zero-copy buffers to the library (no-op for copied tables)."""
        if self._release is not None:
            self._release()
            self._release = None
            self.keys_lo = self.keys_hi = self.count_flags = self.unit_offsets = self.src_kmers = None

    @property
    def n_entries(self) -> int:
        return int(self.keys_lo.size)

    @property
    def multiplicity(self) -> np.ndarray:
        return self.count_flags & np.uint32(0x3FFFFFFF)

    @property
    def flags(self) -> np.ndarray:
        return (self.count_flags >> np.uint32(30)).astype(np.uint8)

    def src_kmer(self, entry: int, k: int) -> bytes:
        """ASCII bases of entry's source k-mer (non-invertible keys)."""
        w = self.src_kmers[entry]
        return bytes(b"ACTG"[(int(w[j >> 5]) >> (2 * (j & 31))) & 3] for j in range(k))

    def colors_of(self, entry: int) -> np.ndarray:
        """Sorted-unique colour ids of one entry (coloured builds)."""
        return self.colors[int(self.color_offsets[entry]):int(self.color_offsets[entry + 1])]

    def unit_slice(self, unit: int) -> slice:
        u = unit - self.first_unit
        return slice(int(self.unit_offsets[u]), int(self.unit_offsets[u + 1]))

    def unit_slice_at(self, i: int) -> slice:
        """Slice of the i-th unit of a table read with read_device_table(units=[...])."""
        return slice(int(self.unit_offsets[i]), int(self.unit_offsets[i + 1]))


class GGCATB200:
    """One context = one build on one GPU (the reference's per-run global state)."""

    def __init__(self, params: Params):
        self._lib = _lib.load()
        pc = _lib.ParamsC(k=params.k, m=params.m, min_multiplicity=params.min_multiplicity,
                          buckets_count_log=params.buckets_count_log,
                          second_buckets_count_log=params.second_buckets_count_log,
                          forward_only=int(params.forward_only), hash_type=params.hash_type,
                          colors=int(params.colors), device=params.device)
        h = C.c_void_p(None)
        _check(self._lib.ggcat_b200_create(C.byref(pc), C.byref(h)))
        self._h = h
        self.params = params
        self.m = params.m or compute_best_m(params.k)
        self._keep = []  # device tensors of imported slices must outlive the merge
        self.peer_connected = False

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.ggcat_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        _check(self._lib.ggcat_b200_reset(self._h))
        self._keep.clear()

    # -- phase 1
    def push_reads(self, data: np.ndarray, offsets: np.ndarray, colors: Optional[np.ndarray] = None):
        data = np.ascontiguousarray(data, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        cp = None
        if colors is not None:
            colors = np.ascontiguousarray(colors, np.uint32)
            cp = colors.ctypes.data
        _check(self._lib.ggcat_b200_push_reads(self._h, data.ctypes.data, offsets.ctypes.data, n, cp))

    def push_reads_ptr(self, data_ptr: int, offsets_ptr: int, n_reads: int, colors_ptr: Optional[int] = None):
        """Host pointers (e.g. pinned buffers) without numpy wrapping."""
        _check(self._lib.ggcat_b200_push_reads(self._h, data_ptr, offsets_ptr, n_reads, colors_ptr))

    def push_reads_packed(self, packed: np.ndarray, offsets: np.ndarray, colors: Optional[np.ndarray] = None):
        """2-bit packed reads: one contiguous stream (base i at bits 2(i % 4) of byte i // 4), offsets in BASES."""
        packed = np.ascontiguousarray(packed, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        cp = None
        if colors is not None:
            colors = np.ascontiguousarray(colors, np.uint32)
            cp = colors.ctypes.data
        _check(self._lib.ggcat_b200_push_reads_packed(self._h, packed.ctypes.data, offsets.ctypes.data, offsets.size - 1, cp))

    def push_reads_packed_ptr(self, packed_ptr: int, offsets_ptr: int, n_reads: int, colors_ptr: Optional[int] = None):
        _check(self._lib.ggcat_b200_push_reads_packed(self._h, packed_ptr, offsets_ptr, n_reads, colors_ptr))

    def push_reads_packed_device(self, d_packed_ptr: int, d_offsets_ptr: int, n_reads: int, n_bases: int,
                                 d_colors_ptr: Optional[int] = None):
        _check(self._lib.ggcat_b200_push_reads_packed_device(self._h, d_packed_ptr, d_offsets_ptr, n_reads, n_bases, d_colors_ptr))

    def push_reads_device(self, d_data_ptr: int, d_offsets_ptr: int, n_reads: int, n_bytes: int,
                          d_colors_ptr: Optional[int] = None):
        _check(self._lib.ggcat_b200_push_reads_device(self._h, d_data_ptr, d_offsets_ptr, n_reads, n_bytes, d_colors_ptr))

    def push_text(self, text, fmt: int = 0, color: int = 0) -> int:
        """Raw FASTA (fmt=0) / FASTQ (fmt=1) text (bytes or uint8 array, whole records): tokenised on the device
        (ggcat_b200_push_text).  Returns the number of records found."""
        a = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, np.uint8)
        n = C.c_uint64(0)
        _check(self._lib.ggcat_b200_push_text(self._h, a.ctypes.data if a.size else None, a.size, fmt, color, C.byref(n)))
        return int(n.value)

    def tokenize(self, text, fmt: int = 0):
        """Test hook: (sequence bytes, offsets) the device tokenizer makes of raw FASTA / FASTQ text."""
        import torch

        a = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, np.uint8)
        d = torch.from_numpy(a.copy()).cuda(self.params.device)
        ps, po, nr, ns = C.c_void_p(0), C.c_void_p(0), C.c_uint64(0), C.c_uint64(0)
        _check(self._lib.ggcat_b200_tokenize_device(self._h, d.data_ptr() if a.size else None, a.size, fmt, C.byref(ps), C.byref(po), C.byref(nr), C.byref(ns)))
        seq = self._dev_bytes(ps.value or 0, int(ns.value))
        off = self._dev_bytes(po.value or 0, (int(nr.value) + 1) * 8).view(np.uint64) if po.value else np.zeros(1, np.uint64)
        return seq, off

    def write_bucket_file(self, bucket: int, path) -> int:
        """One first-level bucket as a reference-format bucket file (ggcat_b200_write_bucket_file); returns its records."""
        n = C.c_uint64(0)
        _check(self._lib.ggcat_b200_write_bucket_file(self._h, bucket, str(path).encode(), C.byref(n)))
        return int(n.value)

    def import_bucket_file(self, bucket: int, path) -> int:
        """Registers the super-k-mers of a reference-format bucket file as a bucket chunk (before finish_bucketing)."""
        n = C.c_uint64(0)
        _check(self._lib.ggcat_b200_import_bucket_file(self._h, bucket, str(path).encode(), C.byref(n)))
        return int(n.value)

    def finish_bucketing(self) -> BucketStats:
        st = _lib.BucketStatsC()
        _check(self._lib.ggcat_b200_finish_bucketing(self._h, C.byref(st)))
        return BucketStats(st.total_bases, st.valid_bases, st.n_superkmers, st.n_kmers, st.payload_words, st.n_buckets,
                           st.n_units)

    def unit_sizes(self):
        nu = ((1 << self.params.buckets_count_log) + 1) << self.params.second_buckets_count_log
        a = np.zeros(nu, np.uint64)
        b = np.zeros(nu, np.uint64)
        _check(self._lib.ggcat_b200_unit_sizes(self._h, a.ctypes.data, b.ctypes.data))
        return a, b

    def dump_superkmers(self, bucket: int):
        """Test hook: (structured array of super-k-mers, payload bytes) of one first-level bucket."""
        n, pb = C.c_uint64(0), C.c_uint64(0)
        _check(self._lib.ggcat_b200_dump_superkmers(self._h, bucket, None, 0, None, 0, C.byref(n), C.byref(pb)))
        out = np.zeros(n.value, SUPERKMER_DTYPE)
        payload = np.zeros(max(pb.value, 1), np.uint8)
        if n.value:
            _check(self._lib.ggcat_b200_dump_superkmers(self._h, bucket, out.ctypes.data, n.value, payload.ctypes.data,
                                                        pb.value, C.byref(n), C.byref(pb)))
        return out, payload[: pb.value]

    # -- phase 2
    def merge_bucket_range(self, first_bucket: int, n_buckets: int, copy: bool = True) -> KmerTable:
        """copy=True: numpy copies, table released immediately.  copy=False: zero-copy views of the library's
        pinned buffers, valid until ``table.release()`` (what a Rust/C host gets from the C ABI)."""
        t = _lib.TableC()
        _check(self._lib.ggcat_b200_merge_bucket_range(self._h, first_bucket, n_buckets, C.byref(t)))

        def arr(ptr, n, dt):
            if not n:
                return np.zeros(0, dt)
            a = np.ctypeslib.as_array(ptr, shape=(n,))
            return a.copy() if copy else a

        try:
            ne = int(t.n_entries)
            keys = arr(t.keys_lo, ne, np.uint64)
            cf = arr(t.count_flags, ne, np.uint32)
            uo = arr(t.unit_offsets, int(t.n_units) + 1, np.uint64)
            hi = arr(t.keys_hi, ne, np.uint64) if t.keys_hi else None
            co = cl = None
            if t.color_offsets:
                co = arr(t.color_offsets, ne + 1, np.uint64)
                cl = arr(t.colors, int(co[-1]), np.uint32)
            src = None
            if t.src_kmers:
                sw = int(t.src_kmer_words)
                src = arr(t.src_kmers, ne * sw, np.uint64).reshape(ne, sw)
            tab = KmerTable(keys, hi, cf, int(t.first_unit), uo, int(t.total_kmers), int(t.unique_kmers), co, cl, src)
        except Exception:
            self._lib.ggcat_b200_release_table(self._h, C.byref(t))
            raise
        if copy:
            self._lib.ggcat_b200_release_table(self._h, C.byref(t))
        else:
            tab._release = lambda: self._lib.ggcat_b200_release_table(self._h, C.byref(t))
        return tab

    def merge_bucket_range_device(self, first_bucket: int, n_buckets: int):
        """Table stays in HBM; returns (n_entries, unique_kmers, total_kmers)."""
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(self._lib.ggcat_b200_merge_bucket_range_device(self._h, first_bucket, n_buckets, C.byref(a), C.byref(b),
                                                              C.byref(c)))
        return a.value, b.value, c.value

    def device_table(self) -> "_lib.TableC":
        """The table left in HBM by the last merge_bucket_range_device (every pointer is device memory)."""
        t = _lib.TableC()
        _check(self._lib.ggcat_b200_device_table(self._h, C.byref(t)))
        return t

    def _dev_bytes(self, addr: int, nbytes: int) -> np.ndarray:
        """Host copy of `nbytes` of device memory at `addr` (verification helpers only)."""
        import torch

        if not nbytes or not addr:
            return np.zeros(0, np.uint8)

        class _V:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (addr, False), "version": 2}

        return torch.as_tensor(_V(), device=torch.device("cuda", self.params.device)).cpu().numpy()

    def read_device_table(self, units: Optional[Sequence[int]] = None) -> KmerTable:
        """Host copy of device_table() -- the whole table, or only the slices of the given (absolute) unit ids, laid out
        back to back in the order given (unit_offsets then has len(units)+1 entries; use .unit_slice_at(i)).
        Verification / tests; a device-side consumer reads the pointers of device_table() directly."""
        t = self.device_table()
        ptr = lambda p: C.cast(p, C.c_void_p).value or 0
        nu = int(t.n_units)
        uo = self._dev_bytes(ptr(t.unit_offsets), (nu + 1) * 8).view(np.uint64)
        wide = bool(t.keys_hi)
        sw = int(t.src_kmer_words) if t.src_kmers else 0
        if units is None:
            ne = int(t.n_entries)
            keys = self._dev_bytes(ptr(t.keys_lo), ne * 8).view(np.uint64)
            hi = self._dev_bytes(ptr(t.keys_hi), ne * 8).view(np.uint64) if wide else None
            cf = self._dev_bytes(ptr(t.count_flags), ne * 4).view(np.uint32)
            src = self._dev_bytes(ptr(t.src_kmers), ne * sw * 8).view(np.uint64).reshape(ne, sw) if sw else None
            return KmerTable(keys, hi, cf, int(t.first_unit), uo, int(t.total_kmers), int(t.unique_kmers), src_kmers=src)
        ks, hs, cs, ss, offs = [], [], [], [], [0]
        for u in units:
            a, b = int(uo[u - int(t.first_unit)]), int(uo[u - int(t.first_unit) + 1])
            ks.append(self._dev_bytes(ptr(t.keys_lo) + a * 8, (b - a) * 8).view(np.uint64))
            if wide:
                hs.append(self._dev_bytes(ptr(t.keys_hi) + a * 8, (b - a) * 8).view(np.uint64))
            cs.append(self._dev_bytes(ptr(t.count_flags) + a * 4, (b - a) * 4).view(np.uint32))
            if sw:
                ss.append(self._dev_bytes(ptr(t.src_kmers) + a * sw * 8, (b - a) * sw * 8).view(np.uint64).reshape(b - a, sw))
            offs.append(offs[-1] + (b - a))
        cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
        tab = KmerTable(cat(ks, np.uint64), cat(hs, np.uint64) if wide else None, cat(cs, np.uint32), 0,
                        np.array(offs, np.uint64), int(t.total_kmers), int(t.unique_kmers),
                        src_kmers=(np.concatenate(ss) if ss else np.zeros((0, sw), np.uint64)) if sw else None)
        tab.n_entries_total = int(t.n_entries)
        return tab

    UNITIG_DTYPE = np.dtype([("word_offset", "<u8"), ("len", "<u4"), ("unit", "<u4"), ("bucket", "<u2"), ("flags", "u1"),
                             ("last_align", "u1"), ("n_kmers", "<u4")])

    def partial_unitigs(self, result_buckets_log: int = 3):
        """Partial unitigs of the table left by the last merge_bucket_range_device, built on the device
        (ggcat_b200_partial_unitigs).  Returns (records: structured array, bases: uint32 words, total k-mers)."""
        u = _lib.UnitigsC()
        _check(self._lib.ggcat_b200_partial_unitigs(self._h, result_buckets_log, C.byref(u)))
        n, nw = int(u.n_unitigs), int(u.n_words)
        recs = np.ctypeslib.as_array(C.cast(u.unitigs, C.POINTER(C.c_uint8)), shape=(n * 24,)).copy().view(self.UNITIG_DTYPE) if n else np.zeros(0, self.UNITIG_DTYPE)
        bases = np.ctypeslib.as_array(u.bases, shape=(nw,)).copy() if nw else np.zeros(0, np.uint32)
        return recs, bases, int(u.n_kmers)

    def maximal_unitigs(self):
        """Maximal unitigs: the partial unitigs of the last partial_unitigs() call joined on the device
        (ggcat_b200_maximal_unitigs).  Same return shape as partial_unitigs()."""
        u = _lib.UnitigsC()
        _check(self._lib.ggcat_b200_maximal_unitigs(self._h, C.byref(u)))
        n, nw = int(u.n_unitigs), int(u.n_words)
        recs = np.ctypeslib.as_array(C.cast(u.unitigs, C.POINTER(C.c_uint8)), shape=(n * 24,)).copy().view(self.UNITIG_DTYPE) if n else np.zeros(0, self.UNITIG_DTYPE)
        bases = np.ctypeslib.as_array(u.bases, shape=(nw,)).copy() if nw else np.zeros(0, np.uint32)
        return recs, bases, int(u.n_kmers)

    # -- multi-GPU plumbing
    def n_chunks(self) -> int:
        return int(self._lib.ggcat_b200_n_chunks(self._h))

    def export_chunk_slice(self, chunk: int, first_unit: int, n_units: int) -> "_lib.ChunkSliceC":
        s = _lib.ChunkSliceC()
        _check(self._lib.ggcat_b200_export_chunk_slice(self._h, chunk, first_unit, n_units, C.byref(s)))
        return s

    def import_chunk_slice(self, first_unit: int, n_units: int, s: "_lib.ChunkSliceC", keepalive=None):
        _check(self._lib.ggcat_b200_import_chunk_slice(self._h, first_unit, n_units, C.byref(s)))
        if keepalive is not None:
            self._keep.append(keepalive)

    def drop_local_chunks(self):
        _check(self._lib.ggcat_b200_drop_local_chunks(self._h))

    # -- multi-GPU exchange over NVLink peer memory (include/ggcat_b200.h, ggcat_b200_peer_*)
    def peer_init(self, rank: int, world: int, arena_bytes: int) -> bytes:
        """Allocates this rank's receive arena; returns its 64-byte CUDA IPC handle."""
        h = (C.c_uint8 * 64)()
        _check(self._lib.ggcat_b200_peer_init(self._h, rank, world, arena_bytes, C.cast(h, C.c_void_p)))
        self.peer_world = world
        return bytes(h)

    def peer_connect(self, handles: Sequence[bytes]):
        """handles: the IPC handles of all ranks, in rank order."""
        buf = (C.c_uint8 * (64 * len(handles))).from_buffer_copy(b"".join(handles))
        _check(self._lib.ggcat_b200_peer_connect(self._h, C.cast(buf, C.c_void_p)))
        self.peer_connected = True

    def peer_exchange(self):
        _check(self._lib.ggcat_b200_peer_exchange(self._h))

    def peer_stats(self) -> tuple[int, int]:
        """(bytes pushed to the other owners, bytes received) in the last peer_exchange."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(self._lib.ggcat_b200_peer_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- measurement
    @property
    def stream_ptr(self) -> int:
        return int(self._lib.ggcat_b200_stream(self._h) or 0)

    def synchronize(self):
        _check(self._lib.ggcat_b200_synchronize(self._h))

    def set_timing(self, on: bool):
        _check(self._lib.ggcat_b200_set_timing(self._h, int(on)))

    def kernel_times(self, reset: bool = True) -> dict:
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        ln = (C.c_uint32 * cap)()
        n = self._lib.ggcat_b200_kernel_times(self._h, names, ms, ln, cap, int(reset))
        if n < 0:
            _check(n)
        return {names[i].decode(): (float(ms[i]), int(ln[i])) for i in range(n)}


# ------------------------------------------------------------------ reference-shaped phase functions
def minimizer_bucketing(input_blocks: Iterable[Sequence], buckets_count_log: int, second_buckets_count_log: int, k: int,
                        m: int = 0, forward_only: bool = False, min_multiplicity: int = 2, colors: bool = False,
                        device: int = 0, hash_type: int = HASH_AUTO) -> tuple[GGCATB200, BucketStats]:
    """Phase 1 ("phase: reads bucketing").  input_blocks: iterable of (data, offsets[, colors]) ASCII batches,
    one per input block/file like the reference's ``Vec<GeneralSequenceBlockData>``; with ``colors`` and no
    explicit colour array, block i gets colour i (file_color = i, lib.rs:300-305)."""
    ctx = GGCATB200(Params(k=k, m=m, min_multiplicity=min_multiplicity, buckets_count_log=buckets_count_log,
                           second_buckets_count_log=second_buckets_count_log, forward_only=forward_only, colors=colors,
                           device=device, hash_type=hash_type))
    for i, blk in enumerate(input_blocks):
        data, offsets = blk[0], blk[1]
        col = blk[2] if len(blk) > 2 else (np.full(len(offsets) - 1, i, np.uint32) if colors else None)
        ctx.push_reads(data, offsets, col)
    return ctx, ctx.finish_bucketing()


def kmers_merge(buckets: GGCATB200, first_bucket: int = 0, n_buckets: Optional[int] = None) -> KmerTable:
    """Phase 2 ("phase: kmers merge") up to the completed, filtered k-mer table."""
    nb = (1 << buckets.params.buckets_count_log) + 1
    if n_buckets is None:
        n_buckets = nb - first_bucket
    return buckets.merge_bucket_range(first_bucket, n_buckets)
