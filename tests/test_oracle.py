"""CPU tests that pin the oracle (oracle/ggcat_oracle.c) to the reference's own known-answer tests,
hash property tests and source constants, and cross-check its literal
BatchMinQueue restatement against the definition-level model in tests/model.py."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import model, util

RNG = np.random.default_rng(773)


def rand_seq(n, alphabet=b"ACGT", rng=RNG):
    return bytes(rng.choice(list(alphabet), n).tolist())


def revcomp(s: bytes) -> bytes:
    t = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}
    return bytes(t[c] for c in s[::-1])


# ------------------------------------------------------------------ reference known-answer tests
def test_packing_kat():
    # crates/io/src/compressed_read.rs:1002-1026 (test_compression, test_rc_compression)
    bases = b"ACGTACGCGGTAGCTAAGCATCGATGCCGATCGTGTTTAACCATG"
    expected_rc = b"CATGGTTAAACACGATCGGCATCGATGCTTAGCTACCGCGTACGT"
    p = O.compress_from_plain(bases)
    assert len(p) == (len(bases) + 3) // 4
    assert O.unpack(p, 0, len(bases)) == bases
    prc = O.compress_from_plain(bases, rc=True)
    assert O.unpack(prc, 0, len(bases)) == expected_rc
    # layout: base i at bits 2(i%4) of byte i/4, A0 C1 T2 G3 (compressed_read.rs:610-618, 882-885)
    assert p[0] == (0 | (1 << 2) | (3 << 4) | (2 << 6))
    # unused high bits of the last byte are zero
    assert p[-1] >> (2 * (len(bases) % 4)) == 0


@pytest.mark.parametrize("n", list(range(1, 40)) + [63, 64, 65, 150])
def test_packing_rc_all_lengths(n):
    s = rand_seq(n)
    assert O.unpack(O.compress_from_plain(s, rc=True), 0, n) == revcomp(s)
    assert O.compress_from_plain(revcomp(s)) == O.compress_from_plain(s, rc=True)


def test_varints_roundtrip():
    # crates/io/src/varint.rs:106-135
    for i in list(range(0, 100000, 7)) + [127, 128, 16383, 16384, 2**32, 2**62]:
        e = O.encode_varint(i)
        v, n = O.decode_varint(e)
        assert (v, n) == (i, len(e))
        e = O.encode_varint_flags(i, i % 4)
        v, f, n = O.decode_varint_flags(e)
        assert (v, f, n) == (i, i % 4, len(e))
    # byte-level layout of varint_flags with 2 flag bits
    assert O.encode_varint_flags(5, 3) == bytes([(3 << 6) | 5])
    assert O.encode_varint_flags(32, 1) == bytes([(1 << 6) | (1 << 5) | 0, 1])


def test_normalize_and_split():
    s = b"acgtNNNxACGTACGTACGTnACG"
    n = O.normalize(s)
    assert n == b"ACGTNNNNACGTACGTACGTNACG"
    assert O.split_segments(n, 4) == [(0, 4), (8, 20)]
    assert O.split_segments(n, 3) == [(0, 4), (8, 20), (21, 24)]
    assert O.split_segments(b"NNNN", 2) == []
    assert O.split_segments(b"", 2) == []


# ------------------------------------------------------------------ constants from source
def test_constants():
    c = O.rk_constants(63)
    M = 1 << 128
    assert c["MULTIPLIER"] == 0x3EB9402F3E733993ADD64D3CA00E1B6B  # cn_rkhash.rs:69
    assert c["MULT_INV"] == 0x9CB6FF6F1B1A6D733E0952E899C3943  # cn_rkhash.rs:70
    assert c["MULTIPLIER"] * c["MULT_INV"] % M == 1
    assert c["RMMULT"] == pow(c["MULTIPLIER"], 62, M)  # hashes/src/lib.rs:169-191
    assert c["MULT_A"] == 0x4751137D01D863C5B8C36DE2B7D399DF
    assert c["MULT_T"] == 0x1E62D96A5E1F5ADE2D4E68D8F88110B7
    assert [O.compute_best_m(k) for k in (31, 63, 15, 21, 38)] == [12, 14, 9, 10, 13]  # utils/src/lib.rs:29-40
    # Appendix B: bucket counts for the BASELINE configs (io/src/lib.rs:67-140)
    assert O.bucket_counts(509_594) == (2, 6)
    assert O.bucket_counts(165_000_000) == (9, 6)
    assert O.bucket_counts(508_000_000) == (10, 6)
    assert O.bucket_counts(100_000_000_000) == (10, 6)


# ------------------------------------------------------------------ hash property tests (hashes/src/lib.rs:265-454)
@pytest.mark.parametrize("m", [5, 12, 14, 31, 32])
def test_nthash_properties(m):
    s = rand_seq(300)
    fw, rc = O.nthash(s, m)
    fw2, rc2 = O.nthash(revcomp(s), m)
    # canonical: hashes(rc(s)) == reverse(hashes(s))
    assert (np.minimum(fw, rc) == np.minimum(fw2, rc2)[::-1]).all()
    assert (fw == rc2[::-1]).all() and (rc == fw2[::-1]).all()
    # rolling == from-scratch closed form (A.2)
    items = model.mmer_items(s, m)
    mn = np.minimum(fw, rc)
    for i in (0, 1, 17, len(items) - 1):
        v = ((int(mn[i]) << 1) & model.MASK64) | int(fw[i] != rc[i])
        assert items[i][0] == v and items[i][1] == bool(fw[i] < rc[i])


@pytest.mark.parametrize("k,ht", [(5, O.HASH_SEQ), (31, O.HASH_SEQ), (32, O.HASH_SEQ), (33, O.HASH_SEQ), (63, O.HASH_SEQ),
                                  (64, O.HASH_SEQ), (31, O.HASH_RK128), (63, O.HASH_RK128), (100, O.HASH_RK128)])
def test_kmer_hash_properties(k, ht):
    s = rand_seq(400)
    lo, hi, fw = O.kmer_hashes(s, k, ht)
    lo2, hi2, fw2 = O.kmer_hashes(revcomp(s), k, ht)
    assert (lo == lo2[::-1]).all() and (hi == hi2[::-1]).all()
    # no collisions among distinct canonical k-mers (>= 64-bit hashes)
    canon = {}
    for i in range(len(s) - k + 1):
        km = s[i:i + k]
        c = min(km, revcomp(km))
        key = (int(lo[i]), int(hi[i]))
        assert canon.setdefault(key, c) == c
    if ht == O.HASH_SEQ:
        # invertible: key bytes LE are the packed canonical-by-value k-mer (cn_seqhash_base.rs:224-230)
        for i in (0, 7, len(s) - k):
            key = int(lo[i]) | (int(hi[i]) << 64)
            dec = bytes(b"ACTG"[(key >> (2 * j)) & 3] for j in range(k))
            assert dec in (s[i:i + k], revcomp(s[i:i + k]))
            f = sum(((c >> 1) & 3) << (2 * j) for j, c in enumerate(s[i:i + k]))
            r = sum((((c >> 1) & 3) ^ 2) << (2 * (k - 1 - j)) for j, c in enumerate(s[i:i + k]))
            assert key == min(f, r) and bool(fw[i]) == (f < r)
    # forward-only variant
    flo, fhi, ffw = O.kmer_hashes(s, k, ht, forward_only=True)
    assert ffw.all()
    sel = fw.astype(bool)
    assert (flo[sel] == lo[sel]).all() and (fhi[sel] == hi[sel]).all()


def test_rk128_closed_form():
    k = 17
    s = rand_seq(60)
    lo, hi, fw = O.kmer_hashes(s, k, O.HASH_RK128)
    c = O.rk_constants(k)
    M = 1 << 128
    L = {0: c["MULT_A"], 1: c["MULT_C"], 2: c["MULT_T"], 3: c["MULT_G"]}
    for i in (0, 3, 43):
        codes = [(ch >> 1) & 3 for ch in s[i:i + k]]
        f = sum(L[b] * pow(c["MULTIPLIER"], k - 1 - j, M) for j, b in enumerate(codes)) % M
        r = sum(L[b ^ 2] * pow(c["MULTIPLIER"], j, M) for j, b in enumerate(codes)) % M
        assert (int(lo[i]) | (int(hi[i]) << 64)) == min(f, r)
        assert bool(fw[i]) == (f < r)


# ------------------------------------------------------------------ window minimum: literal vs definition
@pytest.mark.parametrize("w", [2, 3, 5, 13, 19, 49])
def test_window_minima_vs_bruteforce(w):
    rng = np.random.default_rng(w)
    for trial in range(60):
        n = int(rng.integers(w, 6 * w + 10))
        nvals = int(rng.choice([2, 3, 8, 1000]))
        vals = (rng.integers(0, nvals, n).astype(np.uint64) << np.uint64(1)) | rng.choice([0, 1, 1, 1], n).astype(np.uint64)
        ov, oi = O.window_minima(vals, w)
        assert len(ov) == n - w + 1
        for j in range(n - w + 1):
            W = vals[j:j + w].tolist()
            V = min(W)
            dup = W.count(V) > 1 or (V & 1) == 0
            assert int(ov[j]) == ((V & ~1) if dup else V), (trial, j)
            if not dup:
                assert int(oi[j]) == j + W.index(V)


@pytest.mark.parametrize("k,m", [(31, 12), (63, 14), (15, 9), (21, 10), (13, 9), (32, 12)])
@pytest.mark.parametrize("alphabet", [b"ACGT", b"AC", b"A", b"AT"])
def test_superkmers_literal_vs_definition(k, m, alphabet):
    rng = np.random.default_rng(k * 100 + len(alphabet))
    for trial in range(12):
        L = int(rng.integers(k, 4 * k + 40))
        seg = rand_seq(L, alphabet, rng)
        reads = O.Reads.from_list([seg])
        for fo in (False, True):
            sk, vb = O.bucketing(reads, k, m, 4, 6, forward_only=fo)
            ref = model.superkmers(seg, k, m, 4, 6, forward_only=fo)
            assert vb == L
            assert len(sk) == len(ref)
            for a, b in zip(sk, ref):
                for f in ("start", "len", "bucket", "second_bucket", "minimizer_pos", "flags", "rc"):
                    assert int(a[f]) == b[f], (f, a, b)


def test_superkmer_structure():
    k, m = 31, 12
    seg = rand_seq(1000)
    reads = O.Reads.from_list([seg])
    sk, _ = O.bucketing(reads, k, m, 9, 6)
    # consecutive super-k-mers overlap by exactly k bases; first starts at 0, last ends at L
    assert sk["start"][0] == 0 and sk["start"][-1] + sk["len"][-1] == len(seg)
    ends = sk["start"].astype(np.int64) + sk["len"]
    assert (ends[:-1] - sk["start"][1:] == k).all()
    # wire record round-trips (A.4)
    for row in sk[:20]:
        rec = O.superkmer_record(reads, row, k)
        assert rec[0] == row["second_bucket"]
        v, f, n = O.decode_varint_flags(rec[1:])
        assert f == row["flags"]
        assert v & 31 == row["minimizer_pos"] and (v >> 5) + k == row["len"]
        body = rec[1 + n:]
        s = seg[row["start"]:row["start"] + row["len"]]
        assert O.unpack(body, 0, int(row["len"])) == (revcomp(s) if row["rc"] else s)
        # the minimizer m-mer sits at minimizer_pos of the stored orientation
        if row["bucket"] != 512:
            stored = O.unpack(body, 0, int(row["len"]))
            mm = stored[row["minimizer_pos"]:row["minimizer_pos"] + m]
            fw, rc = O.nthash(mm, m)
            assert fw[0] < rc[0]  # stored orientation makes the minimizer forward
            assert ((int(min(fw[0], rc[0])) << 1) >> 1) % 512 == row["bucket"]


# ------------------------------------------------------------------ end-to-end invariants vs naive counter
def _check_tables_against_naive(reads, k, m, b1, b2, ht, fo, s):
    sk, vb = O.bucketing(reads, k, m, b1, b2, forward_only=fo)
    nv, tot = O.naive_count(reads, k, ht, fo)
    truth = {(int(e["key_lo"]), int(e["key_hi"])): int(e["count"]) for e in nv}
    seen = set()
    occ = 0
    for b in np.unique(sk["bucket"]):
        tab, _, tk = O.merge_unit(reads, sk, int(b), -1, k, s, ht, fo)
        keys = list(zip(tab["key_lo"].tolist(), tab["key_hi"].tolist()))
        assert keys == sorted(keys, key=lambda x: (x[1], x[0]))
        for e, key in zip(tab, keys):
            # invariant (ii): multiplicity == true number of occurrences, in every bucket
            assert int(e["multiplicity"]) == truth[key]
            assert bool(e["kept"]) == (truth[key] >= s)
            seen.add(key)
        # per sub-bucket tables fold to the bucket table
        sub_sum = {}
        for sb in np.unique(sk["second_bucket"][sk["bucket"] == b]):
            t2, _, _ = O.merge_unit(reads, sk, int(b), int(sb), k, s, ht, fo)
            for e in t2:
                kk = (int(e["key_lo"]), int(e["key_hi"]))
                c, f = sub_sum.get(kk, (0, 0))
                sub_sum[kk] = (c + int(e["counter"]), f | int(e["flags"]))
                assert int(e["multiplicity"]) == truth[kk]
        assert {kk: (int(e["counter"]), int(e["flags"])) for e, kk in zip(tab, keys)} == sub_sum
        occ += int(tab["counter"].sum())
    assert seen == set(truth)
    # every k-mer occurrence is stored once, boundary copies twice
    n_boundary = len(sk) - len(np.unique(sk["read_index"].astype(np.int64) * (1 << 32) + 0)) if False else None
    assert occ >= tot


@pytest.mark.parametrize("k,m,ht,fo", [(31, 12, O.HASH_SEQ, False), (31, 12, O.HASH_SEQ, True), (15, 9, O.HASH_SEQ, False),
                                       (63, 14, O.HASH_RK128, False), (63, 14, O.HASH_SEQ, False), (33, 12, O.HASH_SEQ, False)])
def test_tables_match_naive_counter(k, m, ht, fo):
    rng = np.random.default_rng(k + 7 * fo)
    g = rand_seq(3000, rng=rng)
    seqs = []
    for _ in range(120):
        a = int(rng.integers(0, len(g) - 150))
        r = bytearray(g[a:a + 150])
        if rng.random() < 0.5:
            r = bytearray(revcomp(bytes(r)))
        if rng.random() < 0.2:
            r[int(rng.integers(0, 150))] = ord("N")
        seqs.append(bytes(r))
    seqs += [b"A" * 200, b"ACACACACAC" * 20, g[:k], g[:k - 1], b"", b"N" * 50, rand_seq(k + 1, rng=rng)]
    reads = O.Reads.from_list(seqs)
    _check_tables_against_naive(reads, k, m, 3, 2, ht, fo, 2)


def test_colors_sets():
    rng = np.random.default_rng(5)
    g = rand_seq(2000, rng=rng)
    seqs = [g[:1500], g[500:2000], g[200:900], rand_seq(800, rng=rng)]
    reads = O.Reads.from_list(seqs, colors=[0, 1, 2, 3])
    k, m = 31, 12
    sk, _ = O.bucketing(reads, k, m, 2, 2)
    truth = {}
    for c, s in enumerate(seqs):
        lo, hi, _ = O.kmer_hashes(s, k)
        for x in lo.tolist():
            truth.setdefault(x, set()).add(c)
    for b in np.unique(sk["bucket"]):
        tab, cols, _ = O.merge_unit(reads, sk, int(b), -1, k, 1, with_color=True)
        for e in tab:
            got = cols[e["color_off"]:e["color_off"] + e["color_len"]].tolist()
            assert got == sorted(truth[int(e["key_lo"])])


def _oracle_tables(reads, k, m, b1, b2, s, fo):
    sk, _ = O.bucketing(reads, k, m, b1, b2, forward_only=fo)
    nu = ((1 << b1) + 1) << b2
    kl, kh, cf, uo = [], [], [], [0]
    for u in range(nu):
        ref, _, _ = O.merge_unit(reads, sk, u >> b2, u & ((1 << b2) - 1), k, s, O.HASH_SEQ, fo)
        ref = ref[ref["kept"] == 1]
        kl.append(ref["key_lo"]); kh.append(ref["key_hi"])
        cf.append(ref["multiplicity"].astype(np.uint32) | (ref["flags"].astype(np.uint32) << 30))
        uo.append(uo[-1] + len(ref))
    return np.concatenate(kl), np.concatenate(kh), np.concatenate(cf), np.array(uo, np.uint64)


@pytest.mark.parametrize("k,m,b1,b2,s,fo", [(31, 12, 3, 2, 1, False), (31, 12, 2, 1, 2, False), (21, 10, 2, 2, 1, True),
                                             (15, 9, 1, 1, 1, False), (41, 13, 2, 2, 1, False)])
def test_unitigs_per_unit_join_equals_global(k, m, b1, b2, s, fo):
    """Consumer restatement (hashmap.rs:162-297,442-601): partial unitigs per unit with contig breaks at flagged
    k-mers, joined at their shared end k-mers, equal the maximal unitigs of the flag-free global k-mer set
    (independent naive counter): same count, same length multiset, same canonical k-mer set."""
    rng = np.random.default_rng(k + s)
    g = util.rand_seq(rng, 6000)
    cyc = util.rand_seq(rng, 300)
    seqs = [g, util.revcomp(g[1000:3000]), g[2500:5000], g[:200] + g[400:900], util.rand_seq(rng, 500),
            g[5000:5600] + g[100:700], cyc + cyc[:k]]
    if s == 2:
        seqs = seqs + seqs[:4]
    reads = O.Reads.from_list(seqs)
    A = O.unitigs_from_tables(*_oracle_tables(reads, k, m, b1, b2, s, fo), k, fo)
    nv, _ = O.naive_count(reads, k, O.HASH_SEQ, fo)
    nv = nv[nv["count"] >= s]
    B = O.unitigs_from_tables(nv["key_lo"], nv["key_hi"], nv["count"].astype(np.uint32), np.array([0, len(nv)], np.uint64), k, fo)
    assert A["n_partial"] > A["n_unitigs"] == B["n_unitigs"] == B["n_partial"]
    assert np.array_equal(A["lengths"], B["lengths"])
    assert np.array_equal(A["kmers_lo"], nv["key_lo"]) and np.array_equal(A["kmers_hi"], nv["key_hi"])
    assert np.array_equal(B["kmers_lo"], nv["key_lo"])


# ------------------------------------------------------------------ the reference's own test harnesses, re-run on the oracle
class Pcg32:
    """PCG XSH-RR 64/32 with a selectable stream (O'Neill 2014), seeded the way rand_core 0.6.4 `seed_from_u64` fills a
    16-byte seed (Cargo.lock:2373-2376) and the PCG reference implementation consumes it (state, sequence).
    The reference's harness uses `pcg_rand::Pcg32::seed_from_u64(773 + k)` (crates/hashes/src/lib.rs:206-210,269);
    pcg_rand 0.13.0 (Cargo.lock:2126-2129) is not vendored under /root/reference, so its exact seed layout cannot be
    confirmed here -- every check below is a PROPERTY that holds for any base sequence, the stream only picks the sample."""
    MUL = 6364136223846793005
    M64 = (1 << 64) - 1

    def __init__(self, seed_u64: int):
        st = seed_u64 & self.M64
        words = []
        for _ in range(4):   # rand_core::SeedableRng::seed_from_u64: PCG32 outputs fill the seed, 4 bytes at a time
            st = (st * self.MUL + 11634580027462260723) & self.M64
            x = (((st >> 18) ^ st) >> 27) & 0xFFFFFFFF
            rot = st >> 59
            words.append(((x >> rot) | (x << ((32 - rot) & 31))) & 0xFFFFFFFF)
        initstate = words[0] | (words[1] << 32)
        initseq = words[2] | (words[3] << 32)
        self.inc = ((initseq << 1) | 1) & self.M64
        self.state = 0
        self.next_u32()
        self.state = (self.state + initstate) & self.M64
        self.next_u32()

    def next_u32(self) -> int:
        old = self.state
        self.state = (old * self.MUL + self.inc) & self.M64
        x = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((x >> rot) | (x << ((32 - rot) & 31))) & 0xFFFFFFFF


def _harness_bases(k: int) -> bytes:
    """generate_bases(k * 10, 773 + k) of crates/hashes/src/lib.rs:238-247 (decompress_base: 0123 -> ACTG)."""
    rng = Pcg32(773 + k)
    return bytes(b"ACTG"[rng.next_u32() % 4] for _ in range(k * 10))


def _check_hash_properties(hashes_of, k: int, canonical: bool, wide_enough: bool):
    """The stream-independent checks of test_hash_function (crates/hashes/src/lib.rs:265-349): no collisions between
    different k-mers (>= 64-bit hashes), hash(x || x) repeats with period |x|, canonical symmetry."""
    s = _harness_bases(k)
    h = hashes_of(s)
    n = len(s) - k + 1
    assert len(h) == n
    if wide_enough:
        seen = {}
        for i, v in enumerate(h):
            km = s[i:i + k]
            c = min(km, revcomp(km)) if canonical else km
            assert seen.setdefault(v, c) == c, f"collision at k={k}"
    d = hashes_of(s + s)
    assert d[:n] == d[n + k - 1:], f"double-hash test failed at k={k}"
    if canonical:
        assert h == hashes_of(revcomp(s))[::-1], f"canonical test failed at k={k}"


@pytest.mark.parametrize("m", [32, 33, 47, 64, 100, 255, 511])      # cn_nthash.rs:245-248 runs 32..512
def test_reference_hash_harness_nthash(m):
    def hashes_of(s):
        fw, rc = O.nthash(s, m)
        return [int(x) for x in np.minimum(fw, rc)]      # to_unextendable (cn_nthash.rs:96-99, before the << 1)

    _check_hash_properties(hashes_of, m, canonical=True, wide_enough=True)


@pytest.mark.parametrize("k,ht,fo", [(k, O.HASH_SEQ, False) for k in (2, 3, 15, 16, 31, 32, 33, 47, 63)]      # cn_seqhash_base.rs:246-252
                         + [(k, O.HASH_SEQ, True) for k in (2, 31, 32, 63)]                                    # fw_seqhash_base.rs:224-230
                         + [(k, O.HASH_RK128, False) for k in (2, 17, 31, 63, 64, 65, 255, 1000, 4095)]        # cn_rkhash_base.rs:303-306
                         + [(k, O.HASH_RK128, True) for k in (2, 63, 300)])                                    # fw_rkhash_base.rs:250-253
def test_reference_hash_harness_kmer_hashes(k, ht, fo):
    def hashes_of(s):
        lo, hi, _ = O.kmer_hashes(s, k, ht, forward_only=fo)
        return [(int(a), int(b)) for a, b in zip(lo, hi)]

    # seq-hash of tiny k is narrower than 64 bits in the reference too (u16 / u32): collisions impossible anyway (invertible)
    _check_hash_properties(hashes_of, k, canonical=not fo, wide_enough=True)


@pytest.mark.parametrize("inject_duplicates", [False, True])
def test_reference_minqueue_invariant_10m(inject_duplicates):
    """crates/hashes/src/rolling/minqueue_testing.rs:261-365 (minqueue_test; stale against today's callback arity, so it
    is restated here): 10 M random u64 with bit 0 set, window 13 -- for every window the reported minimum equals the true
    window minimum (unique-flag masked) and the unique flag is cleared iff the minimum occurs more than once.
    The second variant adds the duplicates the reference test has commented out (`117 << 32 | 1` pushed with p = 0.5), so
    that the cleared-flag branch is exercised too.  (numpy's PCG64 stream, seed 2: the reference's pcg_rand::Pcg64 is
    not vendored; the invariant does not depend on the stream.)"""
    SIZE, W = 10_000_000, 13
    rng = np.random.Generator(np.random.PCG64(2))
    items = rng.integers(0, 1 << 64, SIZE, dtype=np.uint64) | np.uint64(1)
    if inject_duplicates:
        dup = rng.random(SIZE) < 0.5
        items = np.where(dup, np.uint64((117 << 32) | 1), items)
    items = items[::-1].copy()
    ov, _ = O.window_minima(items, W)
    n = SIZE - W + 1
    assert len(ov) == n
    mask = ~np.uint64(1)
    step = 1 << 20
    for a in range(0, n, step):
        b = min(n, a + step)
        win = np.lib.stride_tricks.sliding_window_view(items[a:b + W - 1], W)
        true_min = win.min(axis=1)
        got = ov[a:b]
        assert np.array_equal(got & mask, true_min & mask)
        count = ((win & mask) == (got & mask)[:, None]).sum(axis=1)
        assert np.array_equal(count > 1, (got & np.uint64(1)) == 0)


def test_bucketing_parallel_equals_bucketing():
    rng = np.random.default_rng(5)
    seqs = [util.rand_seq(rng, int(rng.integers(20, 400))) for _ in range(200)] + [b"ACGTNNACGT" * 20, b"", b"A" * 100]
    reads = O.Reads.from_list(seqs, colors=list(range(len(seqs))))
    a, va = O.bucketing(reads, 31, 12, 3, 2)
    b, vb = O.bucketing_parallel(reads, 31, 12, 3, 2, n_threads=5)
    assert va == vb and np.array_equal(a, b)
