"""Definition-level Python model of SURVEY Appendix A.3 (window minimum + super-k-mer boundaries).

Independent of oracle/ggcat_oracle.c's literal BatchMinQueue restatement: it evaluates every window
from scratch.  Used to cross-check the oracle on small inputs, and it is the position-parallel
formulation the CUDA kernel implements.
"""
MASK64 = (1 << 64) - 1
NT_MULT = 0x397F178C6AE330F9


def _rotl(x, r):
    r %= 64
    return ((x << r) | (x >> (64 - r))) & MASK64 if r else x


def _h(c):  # hashes/src/nthash_base.rs:51-55 on ASCII
    return (((c & 6) + 1) * NT_MULT) & MASK64


def _r(c):  # hashes/src/nthash_base.rs:57-61
    return ((((c & 6) ^ 4) + 1) * NT_MULT) & MASK64


def mmer_items(seg: bytes, m: int):
    out = []
    for i in range(len(seg) - m + 1):
        f = r = 0
        for j in range(m):
            f ^= _rotl(_h(seg[i + j]), m - 1 - j)
            r ^= _rotl(_r(seg[i + j]), j)
        v = ((min(f, r) << 1) & MASK64) | (1 if f != r else 0)
        out.append((v, f < r))
    return out


def windows(seg: bytes, k: int, m: int):
    """[(M_j, argmin_or_None)] for j in 0..L-k+1"""
    items = mmer_items(seg, m)
    w = k - m
    res = []
    for j in range(len(seg) - k + 2):
        vals = [items[i][0] for i in range(j, j + w)]
        V = min(vals)
        dup = vals.count(V) > 1 or (V & 1) == 0
        if dup:
            res.append((V & ~1, None))
        else:
            res.append((V, j + vals.index(V)))
    return res, items


def superkmers(seg: bytes, k: int, m: int, b1: int, b2: int, forward_only: bool = False):
    """List of dicts with start,len,bucket,second_bucket,minimizer_pos,flags,rc (segment coordinates)."""
    L = len(seg)
    assert L >= k
    win, items = windows(seg, k, m)
    n = len(win)
    starts = [0]
    for j in range(1, n):
        Mj, aj = win[j]
        Mp, ap = win[j - 1]
        if Mj != Mp or ((Mj & 1) == 1 and aj != ap):
            starts.append(j)
    out = []
    for t, p in enumerate(starts):
        first = t == 0
        last = t == len(starts) - 1
        q = starts[t + 1] if not last else None
        s = 0 if first else p - 1
        e = L if last else q + k - 1
        M, arg = win[p]
        if (M & 1) == 0:
            bucket, rc, mpos = 1 << b1, False, 0
        else:
            is_fw = items[arg][1]
            rc = (not forward_only) and (not is_fw)
            bucket = (M >> 1) % (1 << b1)
            mpos = (e - arg - m) if rc else (arg - s)
        second = (M >> (b1 + 1)) % (1 << b2)
        flags = ((1 if first else 0) << (1 if rc else 0)) | ((1 if last else 0) << (0 if rc else 1))
        out.append(dict(start=s, len=e - s, bucket=bucket, second_bucket=second, minimizer_pos=mpos & 0xFFFF,
                        flags=flags, rc=int(rc)))
    return out
