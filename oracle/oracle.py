"""ctypes front end of the CPU oracle (oracle/ggcat_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of ggcat_oracle.c.  Nothing under ggcat_b200/
may import this module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs do.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "ggcat_oracle.c"
_SRC2 = _HERE / "ggcat_unitigs.c"
_LIB = _HERE / "_build" / "libggcat_oracle.so"

HASH_SEQ = 1
HASH_RK128 = 4

SUPERKMER_DTYPE = np.dtype(
    [
        ("read_index", "<u4"),
        ("start", "<u4"),
        ("len", "<u4"),
        ("color", "<u4"),
        ("bucket", "<u2"),
        ("minimizer_pos", "<u2"),
        ("second_bucket", "u1"),
        ("flags", "u1"),
        ("rc", "u1"),
        ("pad", "u1"),
    ]
)

TABLE_DTYPE = np.dtype(
    [
        ("key_lo", "<u8"),
        ("key_hi", "<u8"),
        ("counter", "<u8"),
        ("multiplicity", "<u8"),
        ("color_off", "<u4"),
        ("color_len", "<u4"),
        ("flags", "u1"),
        ("kept", "u1"),
        ("pad", "u1", (6,)),
    ]
)

NAIVE_DTYPE = np.dtype([("key_lo", "<u8"), ("key_hi", "<u8"), ("count", "<u8")])


def build(force: bool = False) -> Path:
    """Compile the C restatement with gcc (no external deps)."""
    if _LIB.exists() and not force and _LIB.stat().st_mtime >= max(_SRC.stat().st_mtime, _SRC2.stat().st_mtime):
        return _LIB
    _LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", str(_LIB), str(_SRC), str(_SRC2)]
    subprocess.check_call(cmd)
    return _LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB))
        _lib.orc_sizeof_superkmer.restype = C.c_size_t
        _lib.orc_sizeof_table_entry.restype = C.c_size_t
        assert _lib.orc_sizeof_superkmer() == SUPERKMER_DTYPE.itemsize
        assert _lib.orc_sizeof_table_entry() == TABLE_DTYPE.itemsize
        for name in (
            "orc_split_segments",
            "orc_compress_from_plain",
            "orc_compress_from_plain_rc",
            "orc_encode_varint",
            "orc_encode_varint_flags",
            "orc_decode_varint_flags",
            "orc_decode_varint",
            "orc_nthash_iter",
            "orc_window_minima",
            "orc_bucketing",
            "orc_superkmer_packed",
            "orc_superkmer_record",
            "orc_kmer_hashes",
            "orc_merge_unit",
            "orc_naive_count",
            "orc_compute_best_m",
        ):
            getattr(_lib, name).restype = C.c_size_t
        _lib.orc_seqhash64_get_bucket.restype = C.c_uint16
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _bytes_arr(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


# ---------------------------------------------------------------------------- primitives
def normalize(seq: bytes) -> bytes:
    a = _bytes_arr(seq)
    lib().orc_normalize(_p(a), C.c_size_t(a.size))
    return a.tobytes()


def split_segments(seq: bytes, k: int):
    a = _bytes_arr(seq)
    cap = a.size // max(k, 1) + 2
    st = np.zeros(cap, np.uint64)
    en = np.zeros(cap, np.uint64)
    n = lib().orc_split_segments(_p(a), C.c_size_t(a.size), C.c_size_t(k), _p(st), _p(en), C.c_size_t(cap))
    return [(int(st[i]), int(en[i])) for i in range(n)]


def compress_from_plain(seq: bytes, rc: bool = False) -> bytes:
    a = _bytes_arr(seq)
    out = np.zeros(a.size // 4 + 8, np.uint8)
    f = lib().orc_compress_from_plain_rc if rc else lib().orc_compress_from_plain
    n = f(_p(a), C.c_size_t(a.size), _p(out))
    return out[:n].tobytes()


def unpack(packed: bytes, start: int, n: int) -> bytes:
    a = _bytes_arr(packed)
    out = np.zeros(n, np.uint8)
    lib().orc_unpack(_p(a), C.c_size_t(start), C.c_size_t(n), _p(out))
    return out.tobytes()


def encode_varint(v: int) -> bytes:
    out = np.zeros(16, np.uint8)
    n = lib().orc_encode_varint(C.c_uint64(v), _p(out))
    return out[:n].tobytes()


def decode_varint(b: bytes):
    a = _bytes_arr(b)
    v = C.c_uint64(0)
    n = lib().orc_decode_varint(_p(a), C.c_size_t(a.size), C.byref(v))
    return v.value, n


def encode_varint_flags(v: int, flags: int, flags_count: int = 2) -> bytes:
    out = np.zeros(16, np.uint8)
    n = lib().orc_encode_varint_flags(C.c_uint64(v), C.c_uint8(flags), C.c_int(flags_count), _p(out))
    return out[:n].tobytes()


def decode_varint_flags(b: bytes, flags_count: int = 2):
    a = _bytes_arr(b)
    v = C.c_uint64(0)
    f = C.c_uint8(0)
    n = lib().orc_decode_varint_flags(_p(a), C.c_size_t(a.size), C.c_int(flags_count), C.byref(v), C.byref(f))
    return v.value, f.value, n


def nthash(seq: bytes, m: int):
    a = _bytes_arr(seq)
    n = a.size - m + 1
    fw = np.zeros(max(n, 0), np.uint64)
    rc = np.zeros(max(n, 0), np.uint64)
    if n > 0:
        lib().orc_nthash_iter(_p(a), C.c_size_t(a.size), C.c_size_t(m), _p(fw), _p(rc))
    return fw, rc


def window_minima(vals: np.ndarray, w: int):
    vals = np.ascontiguousarray(vals, np.uint64)
    n = vals.size
    ov = np.zeros(max(n, 1), np.uint64)
    oi = np.zeros(max(n, 1), np.uint32)
    cnt = lib().orc_window_minima(_p(vals), C.c_size_t(n), C.c_size_t(w), _p(ov), _p(oi))
    return ov[:cnt], oi[:cnt]


def kmer_hashes(seq: bytes, k: int, hash_type: int = HASH_SEQ, forward_only: bool = False):
    a = _bytes_arr(seq)
    n = max(a.size - k + 1, 0)
    lo = np.zeros(n, np.uint64)
    hi = np.zeros(n, np.uint64)
    fw = np.zeros(n, np.uint8)
    if n:
        lib().orc_kmer_hashes(_p(a), C.c_size_t(a.size), C.c_size_t(k), C.c_int(hash_type), C.c_int(forward_only), _p(lo), _p(hi), _p(fw))
    return lo, hi, fw


def rk_constants(k: int):
    out = np.zeros(14, np.uint64)
    lib().orc_rk_constants(_p(out), C.c_size_t(k))
    names = ["MULTIPLIER", "MULT_INV", "MULT_A", "MULT_C", "MULT_G", "MULT_T", "RMMULT"]
    return {n: int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i, n in enumerate(names)}


def bucket_counts(bases_count: int):
    a = C.c_uint(0)
    b = C.c_uint(0)
    lib().orc_bucket_counts(C.c_uint64(bases_count), C.byref(a), C.byref(b))
    return a.value, b.value


def compute_best_m(k: int) -> int:
    return int(lib().orc_compute_best_m(C.c_size_t(k)))


# ---------------------------------------------------------------------------- pipeline
class Reads:
    """Concatenated raw ASCII records + offsets (+ optional per-record colour)."""

    def __init__(self, data: np.ndarray, offsets: np.ndarray, colors: np.ndarray | None = None):
        self.data = np.ascontiguousarray(data, np.uint8)
        self.offsets = np.ascontiguousarray(offsets, np.uint64)
        self.colors = None if colors is None else np.ascontiguousarray(colors, np.uint32)
        assert self.offsets[-1] == self.data.size

    @classmethod
    def from_list(cls, seqs, colors=None):
        lens = np.array([len(s) for s in seqs], np.uint64)
        offs = np.zeros(len(seqs) + 1, np.uint64)
        np.cumsum(lens, out=offs[1:])
        data = np.frombuffer(b"".join(bytes(s) for s in seqs), np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
        return cls(data, offs, None if colors is None else np.array(colors, np.uint32))

    @property
    def n(self):
        return self.offsets.size - 1


def bucketing(reads: Reads, k: int, m: int, b1: int, b2: int, forward_only: bool = False):
    """Phase 1 -> structured array of super-k-mers in emission order, valid base count."""
    L = lib()
    cap = max(1024, int(reads.data.size // 4) + reads.n + 16)
    vb = C.c_uint64(0)
    while True:
        out = np.zeros(cap, SUPERKMER_DTYPE)
        n = L.orc_bucketing(
            _p(reads.data), _p(reads.offsets), C.c_size_t(reads.n),
            _p(reads.colors) if reads.colors is not None else None,
            C.c_size_t(k), C.c_size_t(m), C.c_uint(b1), C.c_uint(b2), C.c_int(forward_only),
            _p(out), C.c_size_t(cap), C.byref(vb),
        )
        if n <= cap:
            return out[:n].copy(), vb.value
        cap = n


def bucketing_parallel(reads: Reads, k: int, m: int, b1: int, b2: int, forward_only: bool = False, n_threads: int = 0):
    """bucketing() over contiguous read ranges in threads (records are independent, crates/minimizer_bucketing/src/lib.rs:
    310-320; ctypes releases the GIL): the same rows in the same order, for checkers that need the super-k-mers of 10^7 reads."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    n_threads = n_threads or (os.cpu_count() or 1)
    n = reads.n
    if n_threads <= 1 or n < 4 * n_threads:
        return bucketing(reads, k, m, b1, b2, forward_only)
    cuts = [n * t // n_threads for t in range(n_threads + 1)]

    def part(t):
        a, b = cuts[t], cuts[t + 1]
        o0, o1 = int(reads.offsets[a]), int(reads.offsets[b])
        sub = Reads(reads.data[o0:o1], reads.offsets[a:b + 1] - reads.offsets[a],
                    reads.colors[a:b] if reads.colors is not None else None)
        sk, vb = bucketing(sub, k, m, b1, b2, forward_only)
        sk["read_index"] += np.uint32(a)
        return sk, vb

    with ThreadPoolExecutor(n_threads) as ex:
        res = list(ex.map(part, range(n_threads)))
    return np.concatenate([r[0] for r in res]), sum(r[1] for r in res)


def superkmer_packed(reads: Reads, sk_row) -> bytes:
    row = np.array([sk_row], SUPERKMER_DTYPE)
    out = np.zeros(int(row["len"][0]) // 4 + 8, np.uint8)
    n = lib().orc_superkmer_packed(_p(reads.data), _p(reads.offsets), _p(row), _p(out))
    return out[:n].tobytes()


def superkmer_record(reads: Reads, sk_row, k: int) -> bytes:
    row = np.array([sk_row], SUPERKMER_DTYPE)
    out = np.zeros(int(row["len"][0]) // 4 + 32, np.uint8)
    n = lib().orc_superkmer_record(_p(reads.data), _p(reads.offsets), _p(row), C.c_size_t(k), _p(out))
    return out[:n].tobytes()


def merge_unit(reads: Reads, sk: np.ndarray, bucket: int, second_bucket: int, k: int, min_multiplicity: int,
               hash_type: int = HASH_SEQ, forward_only: bool = False, with_color: bool = False):
    """Phase 2 for one unit (second_bucket=-1 folds the whole first-level bucket).
    Returns (table sorted by key incl. non-kept entries, colours array, total k-mer occurrences)."""
    L = lib()
    sk = np.ascontiguousarray(sk, SUPERKMER_DTYPE)
    sel = sk["bucket"] == bucket
    if second_bucket >= 0:
        sel &= sk["second_bucket"] == second_bucket
    sub = np.ascontiguousarray(sk[sel])
    total = int((sub["len"].astype(np.int64) - k + 1).sum())
    cap = max(total, 1)
    ccap = max(total, 1) if with_color else 1
    out = np.zeros(cap, TABLE_DTYPE)
    cols = np.zeros(ccap, np.uint32)
    ncol = C.c_uint64(0)
    tk = C.c_uint64(0)
    n = L.orc_merge_unit(
        _p(reads.data), _p(reads.offsets), _p(sub), C.c_size_t(sub.size), C.c_int(bucket), C.c_int(second_bucket),
        C.c_size_t(k), C.c_uint64(min_multiplicity), C.c_int(hash_type), C.c_int(forward_only), C.c_int(with_color),
        _p(out), C.c_size_t(cap), _p(cols), C.c_size_t(ccap), C.byref(ncol), C.byref(tk),
    )
    assert n <= cap and ncol.value <= ccap
    return out[:n].copy(), cols[: ncol.value].copy(), tk.value


def naive_count(reads: Reads, k: int, hash_type: int = HASH_SEQ, forward_only: bool = False):
    L = lib()
    lens = (reads.offsets[1:] - reads.offsets[:-1]).astype(np.int64)
    cap = int(np.maximum(lens - k + 1, 0).sum()) + 1
    out = np.zeros(cap, NAIVE_DTYPE)
    tot = C.c_uint64(0)
    n = L.orc_naive_count(_p(reads.data), _p(reads.offsets), C.c_size_t(reads.n), C.c_size_t(k), C.c_int(hash_type),
                          C.c_int(forward_only), _p(out), C.c_size_t(cap), C.byref(tot))
    return out[:n].copy(), tot.value


class PipelineStats(C.Structure):
    _fields_ = [("n_superkmers", C.c_uint64), ("n_kmers", C.c_uint64), ("n_unique", C.c_uint64), ("n_kept", C.c_uint64),
                ("valid_bases", C.c_uint64), ("checksum", C.c_uint64), ("t_bucketing", C.c_double), ("t_merge", C.c_double),
                ("threads", C.c_int)]


def pipeline(reads: Reads, k: int, m: int, b1: int, b2: int, min_multiplicity: int, hash_type: int = HASH_SEQ,
             forward_only: bool = False, n_threads: int = 0) -> PipelineStats:
    """Whole CPU path (phase 1 + phase 2 over all units) with OpenMP: the cpu_baseline of bench.py."""
    st = PipelineStats()
    L = lib()
    L.orc_pipeline.restype = C.c_int
    rc = L.orc_pipeline(_p(reads.data), _p(reads.offsets), C.c_size_t(reads.n), C.c_size_t(k), C.c_size_t(m), C.c_uint(b1),
                        C.c_uint(b2), C.c_int(forward_only), C.c_uint64(min_multiplicity), C.c_int(hash_type),
                        C.c_int(n_threads), C.byref(st))
    assert rc == 0
    return st


def unitigs_from_tables(keys_lo, keys_hi, count_flags, unit_offsets, k: int, forward_only: bool = False):
    """Maximal unitigs from per-unit k-mer tables (the layout of include/ggcat_b200.h): partial unitigs per unit as
    hashmap.rs:442-601, open ends joined.  Returns dict(n_unitigs, n_partial, lengths (sorted), kmers_lo, kmers_hi (sorted))."""
    L = lib()
    keys_lo = np.ascontiguousarray(keys_lo, np.uint64)
    kh = None if keys_hi is None else np.ascontiguousarray(keys_hi, np.uint64)
    cf = np.ascontiguousarray(count_flags, np.uint32)
    uo = np.ascontiguousarray(unit_offsets, np.uint64)
    n = keys_lo.size
    lengths = np.zeros(n + 1, np.uint64)
    klo = np.zeros(n + 1, np.uint64)
    khi = np.zeros(n + 1, np.uint64)
    nu, npart, nk = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    L.orc_unitigs_from_tables.restype = C.c_int
    rc = L.orc_unitigs_from_tables(_p(keys_lo), _p(kh) if kh is not None else None, _p(cf), _p(uo), C.c_size_t(uo.size - 1),
                                   C.c_uint(k), C.c_int(forward_only), C.byref(nu), C.byref(npart), C.byref(nk), _p(lengths),
                                   C.c_size_t(n + 1), _p(klo), _p(khi), C.c_size_t(n + 1))
    if rc != 0:
        raise AssertionError(f"unitig join inconsistent (code {rc})")
    return {"n_unitigs": nu.value, "n_partial": npart.value, "lengths": lengths[: nu.value].copy(),
            "kmers_lo": klo[: nk.value].copy(), "kmers_hi": khi[: nk.value].copy()}


def table_checksum(keys_lo, mult, flags) -> int:
    """Same order-free checksum orc_pipeline accumulates over kept entries (64-bit keys)."""
    with np.errstate(over="ignore"):
        k = np.asarray(keys_lo, np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        v = k + np.asarray(mult, np.uint64) + (np.asarray(flags, np.uint64) << np.uint64(40))
        return int(v.sum(dtype=np.uint64))


def partial_unitigs(keys_lo, keys_hi, count_flags, unit_offsets, first_unit: int, k: int, result_bits: int = 3,
                    forward_only: bool = False):
    """Partial unitigs of every unit of a table + their routing (oracle/ggcat_unitigs.c orc_partial_unitigs: compute_unitigs
    hashmap.rs:442-601, output_sequence final_executor.rs:103-245).  Returns a list of
    (unit, sequence as ASCII bytes, flags, bucket, last_align) in the reference's emission order."""
    L = lib()
    keys_lo = np.ascontiguousarray(keys_lo, np.uint64)
    kh = None if keys_hi is None else np.ascontiguousarray(keys_hi, np.uint64)
    cf = np.ascontiguousarray(count_flags, np.uint32)
    uo = np.ascontiguousarray(unit_offsets, np.uint64)
    n = keys_lo.size
    cap_u, cap_b = n + 1, n * (k + 1) + k + 16
    o_unit = np.zeros(cap_u, np.uint32); o_len = np.zeros(cap_u, np.uint32); o_fl = np.zeros(cap_u, np.uint8)
    o_bk = np.zeros(cap_u, np.uint16); o_al = np.zeros(cap_u, np.uint8); o_off = np.zeros(cap_u, np.uint64)
    bases = np.zeros(cap_b, np.uint8)
    tot = C.c_uint64(0)
    L.orc_partial_unitigs.restype = C.c_size_t
    nu = L.orc_partial_unitigs(_p(keys_lo), _p(kh) if kh is not None else None, _p(cf), _p(uo), C.c_size_t(uo.size - 1),
                               C.c_uint(first_unit), C.c_uint(k), C.c_int(forward_only), C.c_uint(result_bits), _p(o_unit), _p(o_len),
                               _p(o_fl), _p(o_bk), _p(o_al), _p(o_off), C.c_size_t(cap_u), _p(bases), C.c_size_t(cap_b), C.byref(tot))
    assert nu <= cap_u and tot.value <= cap_b
    letters = np.frombuffer(b"ACTG", np.uint8)
    out = []
    for i in range(nu):
        a = int(o_off[i])
        out.append((int(o_unit[i]), letters[bases[a:a + int(o_len[i])]].tobytes(), int(o_fl[i]), int(o_bk[i]), int(o_al[i])))
    return out
