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
