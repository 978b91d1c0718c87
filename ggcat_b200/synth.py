"""Deterministic synthetic genomes / reads for the BASELINE configs.

Counter-based SplitMix64 so every element is a pure function of (seed, index): the same
definition is evaluated by numpy here and could be evaluated per-thread on the device.
Base codes follow the reference alphabet A=0 C=1 T=2 G=3 (crates/utils/src/lib.rs:17,44-46).
"""
from __future__ import annotations

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_LETTERS = np.frombuffer(b"ACTG", dtype=np.uint8)


def splitmix64(seed: int, start: int, n: int) -> np.ndarray:
    """n outputs of SplitMix64 for counters start+1 .. start+n (wrapping u64 arithmetic)."""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def genome_codes(seed: int, length: int) -> np.ndarray:
    """Uniform iid 2-bit codes, 32 bases per SplitMix64 output."""
    nw = (length + 31) // 32
    out = np.empty(nw * 32, np.uint8)
    step = 1 << 20
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    for w0 in range(0, nw, step):
        w = splitmix64(seed, w0, min(step, nw - w0))
        out[w0 * 32:(w0 + w.size) * 32] = ((w[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)
    return out[:length]


def codes_to_ascii(codes: np.ndarray) -> np.ndarray:
    return _LETTERS[codes]


def simulate_reads(genome: np.ndarray, n_reads: int, read_len: int, err_rate: float, seed: int,
                   first_read: int = 0) -> np.ndarray:
    """(n_reads, read_len) uint8 codes.  Read r (global index first_read + r): start uniform in
    [0, G-read_len], strand uniform, iid substitution errors (uniform over the 3 other bases)."""
    G = genome.size
    out = np.empty((n_reads, read_len), np.uint8)
    thr = np.uint64(int(err_rate * (1 << 24)))
    step = 1 << 18
    ar = np.arange(read_len, dtype=np.int64)[None, :]
    for r0 in range(0, n_reads, step):
        nr = min(step, n_reads - r0)
        r = splitmix64(seed, first_read + r0, nr)
        start = ((r >> np.uint64(1)) % np.uint64(G - read_len + 1)).astype(np.int64)
        strand = (r & np.uint64(1)).astype(bool)
        b = genome[start[:, None] + ar]
        b[strand] = b[strand, ::-1] ^ 2
        if err_rate > 0:
            e = splitmix64(seed ^ 0x5EED0E44, (first_read + r0) * read_len, nr * read_len).reshape(nr, read_len)
            hit = (e & np.uint64(0xFFFFFF)) < thr
            delta = (((e >> np.uint64(24)) % np.uint64(3)) + np.uint64(1)).astype(np.uint8)
            b = np.where(hit, (b + delta) & 3, b).astype(np.uint8)
        out[r0:r0 + nr] = b
    return out


def reads_to_ascii_batch(codes2d: np.ndarray):
    """-> (concatenated ASCII bytes, offsets[n+1] uint64)"""
    n, L = codes2d.shape
    data = _LETTERS[codes2d].reshape(-1)
    offsets = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    return np.ascontiguousarray(data), offsets


def pack_2bit(ascii_bases: np.ndarray) -> np.ndarray:
    """ACGT bytes -> the reference's 2-bit stream (code = (c >> 1) & 3, base i at bits 2(i % 4) of byte i // 4,
    crates/utils/src/lib.rs:44-46 + crates/io/src/compressed_read.rs:610-618).  Host-side helper for
    ggcat_b200_push_reads_packed; the input must not contain N (split there first)."""
    a = np.ascontiguousarray(ascii_bases, np.uint8)
    codes = (a >> 1) & 3
    pad = (-codes.size) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros(pad, np.uint8)])
    q = codes.reshape(-1, 4)
    return np.ascontiguousarray(q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)


# ---- named BASELINE configs ---------------------------------------------------------------
def config_c2(n_reads: int = 1_000_000, genome_len: int = 5_000_000, read_len: int = 150, err: float = 0.01,
              seed: int = 0xC2, first_read: int = 0):
    """BASELINE configs[1]: synthetic 5 Mbp bacterial genome, 150 bp reads at 30x with 1% errors."""
    g = genome_codes(seed, genome_len)
    r = simulate_reads(g, n_reads, read_len, err, seed + 1, first_read)
    return reads_to_ascii_batch(r)


def config_c3(n_genomes: int = 100, genome_len: int = 5_000_000, n_shared: int = 20, shared_len: int = 200_000,
              mut: float = 0.001, seed: int = 0xC3):
    """BASELINE configs[2]: n genomes sharing mutated segments of an ancestor; colour = genome index.
    Returns (data, offsets, colors)."""
    anc = genome_codes(seed, genome_len)
    seqs = []
    for gidx in range(n_genomes):
        g = genome_codes(seed + 1 + gidx, genome_len)
        # shared segments: evenly spaced windows copied from the ancestor, mutated at `mut`
        stride = genome_len // n_shared
        for s in range(n_shared):
            a = s * stride
            b = min(a + shared_len, genome_len)
            seg = anc[a:b].copy()
            e = splitmix64((seed + 1 + gidx) ^ 0xABCD, a, b - a)
            hit = (e & np.uint64(0xFFFFFF)) < np.uint64(int(mut * (1 << 24)))
            delta = (((e >> np.uint64(24)) % np.uint64(3)) + np.uint64(1)).astype(np.uint8)
            seg = np.where(hit, (seg + delta) & 3, seg).astype(np.uint8)
            g[a:b] = seg
        seqs.append(g)
    data = _LETTERS[np.concatenate(seqs)]
    offsets = np.arange(n_genomes + 1, dtype=np.uint64) * np.uint64(genome_len)
    colors = np.arange(n_genomes, dtype=np.uint32)
    return np.ascontiguousarray(data), offsets, colors


# ---- device-side generation (torch on the GPU) -------------------------------------------------------
# The same counter-based definitions as above evaluated with wrapping int64 arithmetic on the device, for the
# configs whose reads do not fit a host round trip (C4: "reads generated on device per GPU slice").
def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z, s: int):
    return (z >> s) & ((1 << (64 - s)) - 1)


def splitmix64_torch(seed: int, idx):
    """idx: int64 tensor of counter values (numpy version: start+1 .. start+n).  Returns int64 (bit pattern of the u64)."""
    z = idx * _s64(0x9E3779B97F4A7C15) + _s64(seed)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def genome_codes_torch(seed: int, length: int, device):
    import torch

    nw = (length + 31) // 32
    out = torch.empty(nw * 32, dtype=torch.uint8, device=device)
    step = 1 << 22
    shifts = (torch.arange(32, dtype=torch.int64, device=device) * 2)[None, :]
    for w0 in range(0, nw, step):
        n = min(step, nw - w0)
        w = splitmix64_torch(seed, torch.arange(w0 + 1, w0 + n + 1, dtype=torch.int64, device=device))
        out[w0 * 32:(w0 + n) * 32] = ((w[:, None] >> shifts) & 3).to(torch.uint8).reshape(-1)
    return out[:length]


def simulate_reads_torch(genome, n_reads: int, read_len: int, err_rate: float, seed: int, first_read: int = 0, out=None):
    """ASCII reads (n_reads * read_len uint8, device) with the definition of simulate_reads()."""
    import torch

    dev = genome.device
    G = genome.numel()
    if out is None:
        out = torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    letters = torch.tensor(list(b"ACTG"), dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, dtype=torch.int64, device=dev)[None, :]
    thr = int(err_rate * (1 << 24))
    step = 1 << 20
    for r0 in range(0, n_reads, step):
        nr = min(step, n_reads - r0)
        r = splitmix64_torch(seed, torch.arange(first_read + r0 + 1, first_read + r0 + nr + 1, dtype=torch.int64, device=dev))
        start = _lsr(r, 1) % (G - read_len + 1)
        strand = (r & 1).bool()
        b = genome[start[:, None] + ar]
        b = torch.where(strand[:, None], b.flip(1) ^ 2, b)
        if err_rate > 0:
            c0 = (first_read + r0) * read_len
            e = splitmix64_torch(seed ^ 0x5EED0E44, torch.arange(c0 + 1, c0 + nr * read_len + 1, dtype=torch.int64, device=dev)).view(nr, read_len)
            hit = (e & 0xFFFFFF) < thr
            delta = (_lsr(e, 24) % 3 + 1).to(torch.uint8)
            b = torch.where(hit, (b + delta) & 3, b)
        out[r0 * read_len:(r0 + nr) * read_len] = letters[b.long()].reshape(-1)
    return out
