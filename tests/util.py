"""Shared helpers for the tests (inputs, oracle digests)."""
import gzip
import hashlib
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def read_fasta_records(path):
    """Minimal FASTA reader (host I/O is outside the hot path): list of raw sequence bytes."""
    op = gzip.open if str(path).endswith(".gz") else open
    recs, cur = [], []
    with op(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if cur:
                    recs.append(b"".join(cur))
                cur = []
            elif line and not line.startswith(b";"):
                cur.append(line)
    if cur:
        recs.append(b"".join(cur))
    return recs


def c1_records():
    recs = []
    for f in ("sal1.fa.gz", "sal2.fa.gz", "sal3.fa.gz"):
        recs += read_fasta_records(GOLDEN / f)
    return recs


def table_digest(keys_lo, keys_hi, mult, flags) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(keys_lo, "<u8").tobytes())
    if keys_hi is not None:
        h.update(np.ascontiguousarray(keys_hi, "<u8").tobytes())
    h.update(np.ascontiguousarray(mult, "<u8").tobytes())
    h.update(np.ascontiguousarray(flags, "u1").tobytes())
    return h.hexdigest()


def rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(list(alphabet), n).tolist())


def revcomp(s: bytes) -> bytes:
    t = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}
    return bytes(t.get(c, 78) for c in s[::-1])


def bincode_varint(buf, p):
    b = buf[p]
    if b < 251:
        return b, p + 1
    n = {251: 2, 252: 4, 253: 8}[b]
    return int.from_bytes(buf[p + 1:p + 1 + n], "little"), p + 1 + n


def parse_bucket_file(buf: bytes, k: int):
    """Independent reader of a PLAIN SingleGrouped bucket file, following
    parallel-processor-rs/src/buckets/readers/binary_reader.rs:120-190 and crates/io/src/varint.rs:58-84:
    -> {sub_bucket: [record bytes incl. the varint_flags header]} in file order."""
    assert buf[:16] == b"PLAIN_INTR_BKT_M"
    index_offset = int.from_bytes(buf[16:24], "little")
    fmt = buf[24:56]
    assert fmt[0] == 1 and not any(fmt[1:]), "data_format_info = bincode(MinimizerBucketMode::SingleGrouped)"
    n, p = bincode_varint(buf, index_offset)
    cps = []
    for _ in range(n):
        off, p = bincode_varint(buf, p)
        opt = buf[p]; p += 1
        data = None
        if opt == 1:
            ln, p = bincode_varint(buf, p)
            sub, q = bincode_varint(buf, p)
            cnt, q = bincode_varint(buf, q)
            assert q == p + ln
            data = (sub, cnt); p += ln
        cps.append((off, data))
    assert p == len(buf), "the checkpoint index is the last thing in the file"
    assert cps[0] == (56, None), "first checkpoint: right behind the header, no data"
    assert [c[0] for c in cps] == sorted(c[0] for c in cps)
    klog = (k - 1).bit_length()
    out = {}
    for i, (off, data) in enumerate(cps):
        if data is None:
            continue
        end = cps[i + 1][0] if i + 1 < len(cps) else index_offset
        recs, q = [], off
        while q < end:
            start = q
            f = buf[q]; q += 1
            v, sh, nxt = f & 31, 5, bool(f & 32)
            while nxt:
                b = buf[q]; q += 1
                nxt = bool(b & 128); v |= (b & 127) << sh; sh += 7
            ln = (v >> klog) + k
            q += (ln + 3) // 4
            recs.append(bytes(buf[start:q]))
        assert q == end and len(recs) == data[1]
        assert data[0] not in out, "one chunk per sub-bucket"
        out[data[0]] = recs
    return out
