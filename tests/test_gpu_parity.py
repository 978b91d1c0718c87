"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(include/ggcat_b200.h) and is compared bit-for-bit with the CPU oracle on the same inputs."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g

    g.build()
    import ggcat_b200 as G

    return G


def _sk_records(reads, sk, k):
    """Canonical comparable form of oracle super-k-mers: sorted list of tuples."""
    out = []
    for row in sk:
        out.append((int(row["bucket"]), int(row["second_bucket"]), int(row["len"]), int(row["flags"]), int(row["rc"]),
                    int(row["minimizer_pos"]), O.superkmer_packed(reads, row)))
    return sorted(out)


def _gpu_records(ctx, n_buckets):
    out = []
    for b in range(n_buckets):
        sk, payload = ctx.dump_superkmers(b)
        pb = payload.tobytes()
        for row in sk:
            nb = (int(row["len"]) + 3) // 4
            off = int(row["payload_offset"])
            out.append((int(row["bucket"]), int(row["second_bucket"]), int(row["len"]), int(row["flags"]), int(row["rc"]),
                        int(row["minimizer_pos"]), pb[off:off + nb]))
    return sorted(out)


def _check_src_kmers(tab, sl, k, hash_type, limit=64):
    """Non-invertible keys (rk128) come with the bases of one occurrence (the reference's saved_reads contract,
    hashmap.rs:96-149): re-hash them on the host -- their FORWARD hash must be the key."""
    if hash_type != O.HASH_RK128:
        assert tab.src_kmers is None
        return
    assert tab.src_kmers is not None and tab.src_kmers.shape == (tab.keys_lo.size, max(2, (k + 31) // 32))
    idx = list(range(sl.start, sl.stop))
    if len(idx) > limit:
        idx = idx[:: max(1, len(idx) // limit)]
    for e in idx:
        seq = tab.src_kmer(e, k)
        lo, hi, _ = O.kmer_hashes(seq, k, O.HASH_RK128, True)
        assert (int(lo[0]), int(hi[0])) == (int(tab.keys_lo[e]), int(tab.keys_hi[e])), f"entry {e}: source bases {seq} do not hash to the key"
        full = sum(int(w) << (64 * q) for q, w in enumerate(tab.src_kmers[e]))
        assert full >> (2 * k) == 0, "bits above the k-mer must be clear"


def _check_tables(G, ctx, reads, sk, k, s, b1, b2, forward_only=False, ranges=None, hash_type=O.HASH_SEQ, colors=False):
    nb = (1 << b1) + 1
    ranges = ranges or [(0, nb)]
    n_checked = 0
    for (fb, cnt) in ranges:
        tab = ctx.merge_bucket_range(fb, cnt)
        assert tab.first_unit == fb << b2
        tot_occ = 0
        uniq = 0
        for u in range(fb << b2, (fb + cnt) << b2):
            ref, rcols, tk = O.merge_unit(reads, sk, u >> b2, u & ((1 << b2) - 1), k, s, hash_type, forward_only, with_color=colors)
            tot_occ += tk
            uniq += len(ref)
            ref = ref[ref["kept"] == 1]
            sl = tab.unit_slice(u)
            assert np.array_equal(tab.keys_lo[sl], ref["key_lo"]), f"unit {u}: keys differ"
            if tab.keys_hi is not None:
                assert np.array_equal(tab.keys_hi[sl], ref["key_hi"]), f"unit {u}: high key words differ"
            else:
                assert not ref["key_hi"].any()
            assert np.array_equal(tab.multiplicity[sl].astype(np.uint64), ref["multiplicity"]), f"unit {u}: counts"
            assert np.array_equal(tab.flags[sl], ref["flags"]), f"unit {u}: flags"
            _check_src_kmers(tab, sl, k, hash_type)
            if colors:
                for j, e in enumerate(range(sl.start, sl.stop)):
                    want = rcols[int(ref["color_off"][j]):int(ref["color_off"][j]) + int(ref["color_len"][j])]
                    assert np.array_equal(tab.colors_of(e), want), f"unit {u} entry {j}: colour sets differ"
            n_checked += len(ref)
        assert tab.total_kmers == tot_occ
        if not colors:
            assert tab.unique_kmers == uniq
        if colors:
            assert int(tab.color_offsets[-1]) == tab.colors.size
    return n_checked


def _mixed_reads(rng, k, n=300):
    g = util.rand_seq(rng, 5000)
    seqs = []
    for _ in range(n):
        L = int(rng.integers(k - 2, 260))
        a = int(rng.integers(0, len(g) - L))
        r = bytearray(g[a:a + L])
        if rng.random() < 0.5:
            r = bytearray(util.revcomp(bytes(r)))
        if rng.random() < 0.15:
            r[int(rng.integers(0, L))] = ord("N")
        if rng.random() < 0.05:
            r[int(rng.integers(0, L))] = ord("x")
        if rng.random() < 0.1:
            r = bytearray(bytes(r).lower())
        seqs.append(bytes(r))
    seqs += [b"A" * 300, b"ACACACACAC" * 30, b"AT" * 100, g[:k], g[:k - 1], b"", b"N" * 50, b"ACGT" * 2,
             util.rand_seq(rng, k + 1), g[:3000], b"N" + g[100:400] + b"NN" + g[500:531] + b"N" + g[600:630] + b"N"]
    return seqs


@pytest.mark.parametrize("k,m,b1,b2,fo,s", [(31, 12, 3, 2, False, 2), (31, 12, 4, 6, True, 1), (15, 9, 2, 1, False, 1),
                                            (21, 10, 3, 3, False, 3), (27, 5, 2, 2, False, 2), (31, 29, 2, 2, False, 1)])
def test_small_mixed_inputs(k, m, b1, b2, fo, s):
    G = _gpu()
    rng = np.random.default_rng(k * 31 + m)
    seqs = _mixed_reads(rng, k)
    reads = O.Reads.from_list(seqs)
    sk, vb = O.bucketing(reads, k, m, b1, b2, forward_only=fo)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, forward_only=fo, min_multiplicity=s)
    try:
        assert st.n_superkmers == len(sk)
        assert st.n_kmers == int((sk["len"].astype(np.int64) - k + 1).sum())
        assert _gpu_records(ctx, (1 << b1) + 1) == _sk_records(reads, sk, k)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2, fo, ranges=[(0, 1), (1, (1 << b1))])
    finally:
        ctx.close()


def test_long_sequences_cross_tiles():
    """Sequences much longer than a 1024-window tile, incl. low-complexity runs longer than a tile."""
    G = _gpu()
    rng = np.random.default_rng(99)
    k, m, b1, b2, s = 31, 12, 3, 2, 1
    seqs = [util.rand_seq(rng, 40000), b"A" * 5000 + util.rand_seq(rng, 3000) + b"CA" * 2000, util.rand_seq(rng, 1023 + k),
            util.rand_seq(rng, 1024 + k - 2), util.rand_seq(rng, 2048), b"G" * 1024, b"T" * 1025]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        assert st.n_superkmers == len(sk)
        assert _gpu_records(ctx, (1 << b1) + 1) == _sk_records(reads, sk, k)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)
    finally:
        ctx.close()


@pytest.mark.parametrize("n_bases", [3000, 8000, 14000, 60000, 400000])
def test_unit_size_paths(n_bases):
    """Units sized for each merge variant: <= 6144 records (512-thread smem), <= 12288 (1024-thread smem),
    larger (global-scratch sort)."""
    G = _gpu()
    rng = np.random.default_rng(n_bases)
    k, m, b1, b2, s = 31, 12, 0, 0, 1
    g = util.rand_seq(rng, n_bases // 3)
    seqs = [g, util.revcomp(g[: len(g) // 2]), g[len(g) // 4:], util.rand_seq(rng, n_bases // 6)]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        assert st.n_superkmers == len(sk)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)
    finally:
        ctx.close()


@pytest.mark.parametrize("n_bases", [5800, 11500, 20000, 45000])
def test_all_distinct_units_retry_paths(n_bases):
    """-s 1 on a single random sequence: nearly every k-mer survives, so the hash kernels cannot sort the
    survivors inside their table and hand the unit to the sort-based kernels (smem and global retry lists)."""
    G = _gpu()
    rng = np.random.default_rng(n_bases + 1)
    k, m, b1, b2, s = 31, 12, 0, 0, 1
    reads = O.Reads.from_list([util.rand_seq(rng, n_bases)])
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)
    finally:
        ctx.close()


def test_multiple_pushes_equal_single_push():
    """Chunk invariance: the table does not depend on how the input was batched."""
    G = _gpu()
    rng = np.random.default_rng(5)
    k, m, b1, b2, s = 31, 12, 3, 2, 2
    seqs = _mixed_reads(rng, k, n=600)
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    half = len(seqs) // 2
    r1 = O.Reads.from_list(seqs[:half])
    r2 = O.Reads.from_list(seqs[half:])
    ctx, st = G.minimizer_bucketing([(r1.data, r1.offsets), (r2.data, r2.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        assert ctx.n_chunks() == 2
        assert st.n_superkmers == len(sk)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)
    finally:
        ctx.close()


@pytest.mark.parametrize("k,ht", [(31, O.HASH_SEQ), (63, O.HASH_RK128)])
def test_pipelined_host_path_small_batches(k, ht, monkeypatch):
    """push_reads splits the host input into double-buffered H2D batches (all accumulated into one bucket chunk) and
    merge_bucket_range appends parts whose D2H overlaps the next merge: force many small batches / parts and require the
    identical table."""
    G = _gpu()
    monkeypatch.setenv("GGCAT_B200_HOST_BATCH", "3000")
    monkeypatch.setenv("GGCAT_B200_PART_KMERS", "2500")
    rng = np.random.default_rng(17 + k)
    m, b1, b2, s = (12 if k == 31 else 14), 3, 2, 2
    seqs = _mixed_reads(rng, k, n=500)
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s, hash_type=ht)
    try:
        assert ctx.n_chunks() == 1      # the H2D batches of one push accumulate into ONE bucket chunk
        assert st.n_superkmers == len(sk)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2, hash_type=ht)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2, hash_type=ht, ranges=[(2, 5)])
    finally:
        ctx.close()


@pytest.mark.parametrize("k,ht,s", [(31, O.HASH_SEQ, 1), (31, O.HASH_SEQ, 2), (63, O.HASH_RK128, 1), (41, O.HASH_SEQ, 2)])
def test_final_table_growth_and_device_parts(k, ht, s, monkeypatch):
    """The final table is sized from an estimate of the survivors and grows when a part does not fit; the
    device-resident merge appends bounded parts.  Force a tiny estimate and tiny parts: identical tables (host path,
    which shares the growth code) and identical counters from merge_bucket_range_device."""
    G = _gpu()
    monkeypatch.setenv("GGCAT_B200_FINAL_EST", "7")
    monkeypatch.setenv("GGCAT_B200_PART_KMERS", "3000")
    monkeypatch.setenv("GGCAT_B200_PART_KMERS_DEV", "3000")
    rng = np.random.default_rng(900 + k + s)
    m, b1, b2 = (12 if k == 31 else 14), 3, 2
    seqs = _mixed_reads(rng, k, n=400) + [util.rand_seq(rng, 40000)]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s, hash_type=ht)
    try:
        n_checked = _check_tables(G, ctx, reads, sk, k, s, b1, b2, hash_type=ht)
        ne, uq, tk = ctx.merge_bucket_range_device(0, (1 << b1) + 1)
        assert ne == n_checked and tk == st.n_kmers
        ne2, _, _ = ctx.merge_bucket_range_device(2, 4)      # a sub-range after a full one: the table restarts at entry 0
        tab = ctx.merge_bucket_range(2, 4)
        assert ne2 == tab.n_entries
    finally:
        ctx.close()


def test_several_large_units_in_one_range():
    """More than one unit above the shared-memory capacity in the same merge call (global-scratch tables, work list
    of large units sorted by size)."""
    G = _gpu()
    rng = np.random.default_rng(4242)
    k, m, b1, b2, s = 31, 12, 1, 1, 2
    g = util.rand_seq(rng, 120000)
    seqs = [g, util.revcomp(g[:90000]), g[20000:], util.rand_seq(rng, 30000)]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        _, km = ctx.unit_sizes()
        assert (km > 12288).sum() >= 3
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)
    finally:
        ctx.close()


@pytest.mark.parametrize("s", [1, 2, 40])
def test_key_partition_overflow_falls_back(s):
    """A big unit whose k-mers are few but very frequent (a tandem repeat): some key partitions exceed their capacity,
    the unit is redone by the global-table kernel; result must be identical.  Also a normal big unit beside it."""
    G = _gpu()
    rng = np.random.default_rng(77)
    k, m, b1, b2 = 31, 12, 1, 0
    motif = util.rand_seq(rng, 211)
    seqs = [motif * 400, util.rand_seq(rng, 70000), util.revcomp(motif * 150)]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        _, km = ctx.unit_sizes()
        assert (km > 12288).sum() >= 2
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)
    finally:
        ctx.close()


@pytest.mark.parametrize("s,ratio", [(1, "0.01"), (2, "0.01"), (1, "0.1")])
def test_key_partitions_sized_by_distinct_keys_split_when_wrong(s, ratio, monkeypatch):
    """Key partitions are sized from the distinct/records ratio of earlier parts.  Pretend a ratio of 1 % on data whose
    k-mers are almost all distinct: partitions of tens of thousands of records overflow the 8192-slot table and
    k_merge_parts must split their key space (twice) -- identical tables and counters."""
    G = _gpu()
    monkeypatch.setenv("GGCAT_B200_DISTINCT_RATIO", ratio)
    rng = np.random.default_rng(99 + s)
    k, m, b1, b2 = 31, 12, 1, 0
    g = util.rand_seq(rng, 90000)
    seqs = [g, util.rand_seq(rng, 60000), util.revcomp(g[10000:40000])]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        _, km = ctx.unit_sizes()
        assert (km > 30000).sum() >= 2
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)   # second merge: the ratio now comes from the first one
    finally:
        ctx.close()


@pytest.mark.parametrize("k,ht", [(31, O.HASH_SEQ), (63, O.HASH_RK128)])
def test_owner_side_import_on_one_gpu(k, ht):
    """The multi-GPU owner path on one device: context A buckets two pushes, its per-owner chunk slices are copied
    (as the all-to-all would deliver them) and imported into a fresh context per owner, which merges its bucket
    range; tables must equal the oracle's."""
    G = _gpu()
    import torch

    from ggcat_b200 import _lib, dist as gdist

    rng = np.random.default_rng(k)
    m, b1, b2, s = (12 if k == 31 else 14), 3, 2, 2
    seqs = _mixed_reads(rng, k, n=600) + [util.rand_seq(rng, 50000)]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    half = len(seqs) // 2
    r1, r2 = O.Reads.from_list(seqs[:half]), O.Reads.from_list(seqs[half:])
    A, _ = G.minimizer_bucketing([(r1.data, r1.offsets), (r2.data, r2.offsets)], b1, b2, k, m, min_multiplicity=s, hash_type=ht)
    dev = torch.device("cuda", 0)
    world = 3
    owner = gdist.OwnerMap(b1, b2, world)
    try:
        for rank in range(world):
            B = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2, hash_type=ht))
            try:
                fu, nu = owner.unit_range(rank)
                for c in range(A.n_chunks()):
                    sl = A.export_chunk_slice(c, fu, nu)
                    desc = gdist._view(sl.d_descriptors, int(sl.n_superkmers) * 16, torch.uint8, dev).clone()
                    pay = torch.zeros(int(sl.n_words) + 8, dtype=torch.int32, device=dev)
                    pay[: int(sl.n_words)] = gdist._view(sl.d_payload, int(sl.n_words), torch.int32, dev)
                    uc = gdist._view(sl.d_unit_counts, nu, torch.int32, dev).clone()
                    uw = gdist._view(sl.d_unit_words, nu, torch.int32, dev).clone()
                    uk = gdist._view(sl.d_unit_kmers, nu, torch.int32, dev).clone()
                    torch.cuda.synchronize()
                    s2 = _lib.ChunkSliceC(n_superkmers=sl.n_superkmers, n_words=sl.n_words, word_bias=sl.word_bias,
                                          d_descriptors=desc.data_ptr(), d_payload=pay.data_ptr(), d_unit_counts=uc.data_ptr(),
                                          d_unit_words=uw.data_ptr(), d_unit_kmers=uk.data_ptr())
                    B.import_chunk_slice(fu, nu, s2, keepalive=(desc, pay, uc, uw, uk))
                B.finish_bucketing()
                fb, nb = owner.bucket_range(rank)
                _check_tables(G, B, reads, sk, k, s, b1, b2, ranges=[(fb, nb)], hash_type=ht)
            finally:
                B.close()
    finally:
        A.close()


def test_c1_example_inputs(golden_dir):
    """BASELINE configs[0]: sal1+sal2+sal3, k=31 -s 1, 4(+1) x 64 buckets -- full parity + committed digests."""
    G = _gpu()
    gold = json.loads((golden_dir / "c1_golden.json").read_text())
    recs = util.c1_records()
    reads = O.Reads.from_list(recs)
    k, m, b1, b2, s = gold["k"], gold["m"], gold["b1"], gold["b2"], gold["s"]
    assert G.bucket_counts(509_594) == (b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        assert st.n_superkmers == gold["n_superkmers"]
        sk, _ = O.bucketing(reads, k, m, b1, b2)
        assert _gpu_records(ctx, (1 << b1) + 1) == _sk_records(reads, sk, k)
        for b in range((1 << b1) + 1):
            tab = ctx.merge_bucket_range(b, 1)
            # fold sub-buckets into the bucket table: sum counters, OR flags, halve if flags == 3
            keys, inv = np.unique(tab.keys_lo, return_inverse=True)
            cnt = np.zeros(keys.size, np.uint64)
            # a unit's raw MapEntry counter is its multiplicity un-halved
            raw = tab.multiplicity.astype(np.uint64) << (tab.flags == 3).astype(np.uint64)
            np.add.at(cnt, inv, raw)
            fl = np.zeros(keys.size, np.uint8)
            np.bitwise_or.at(fl, inv, tab.flags)
            mult = cnt >> (fl == 3).astype(np.uint64)
            g = gold["buckets"][str(b)]
            assert tab.total_kmers == g["n_kmer_occurrences"]
            assert keys.size == g["n_entries"]
            assert util.table_digest(keys, None, mult, fl) == g["digest"], f"bucket {b}"
    finally:
        ctx.close()


def test_c2_full_size_properties():
    """BASELINE configs[1] at full size (1 M x 150 bp): size-independent properties + sampled unit parity."""
    G = _gpu()
    from ggcat_b200 import synth

    n_reads = int(os.environ.get("GGCAT_TEST_C2_READS", "1000000"))
    data, offsets = synth.config_c2(n_reads=n_reads)
    k, m, s = 31, 12, 2
    b1, b2 = G.bucket_counts(int(data.size * 1.1))
    ctx, st = G.minimizer_bucketing([(data, offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        n_kmers_input = n_reads * (150 - k + 1)
        # every k-mer occurrence is stored once per (k-1)-mer group side: boundary k-mers twice
        assert st.n_kmers == n_kmers_input + (st.n_superkmers - n_reads)
        sizes_sk, sizes_km = ctx.unit_sizes()
        assert int(sizes_sk.sum()) == st.n_superkmers and int(sizes_km.sum()) == st.n_kmers
        nb = (1 << b1) + 1
        t1 = ctx.merge_bucket_range(0, nb)
        assert t1.total_kmers == st.n_kmers
        # sortedness inside every unit, multiplicity filter respected
        assert (t1.multiplicity >= s).all()
        uo = t1.unit_offsets
        d = np.diff(t1.keys_lo.astype(np.uint64))
        bad = np.nonzero(t1.keys_lo[1:] <= t1.keys_lo[:-1])[0] + 1
        assert np.isin(bad, uo).all(), "keys not strictly increasing inside a unit"
        # idempotence: a second merge gives the identical table
        t2 = ctx.merge_bucket_range(0, nb)
        assert np.array_equal(t1.keys_lo, t2.keys_lo) and np.array_equal(t1.count_flags, t2.count_flags)
        assert np.array_equal(t1.unit_offsets, t2.unit_offsets)
        # a k-mer has the same multiplicity wherever it appears (checksum over duplicates across units)
        order = np.argsort(t1.keys_lo, kind="stable")
        ks, ms = t1.keys_lo[order], t1.multiplicity[order]
        same = ks[1:] == ks[:-1]
        assert (ms[1:][same] == ms[:-1][same]).all()
        # sampled units against the oracle
        reads = O.Reads(data, offsets)
        sk, _ = O.bucketing(reads, k, m, b1, b2)
        assert st.n_superkmers == len(sk)
        rng = np.random.default_rng(1)
        for u in rng.choice(nb << b2, 24, replace=False):
            ref, _, _ = O.merge_unit(reads, sk, int(u) >> b2, int(u) & ((1 << b2) - 1), k, s)
            ref = ref[ref["kept"] == 1]
            sl = t1.unit_slice(int(u))
            assert np.array_equal(t1.keys_lo[sl], ref["key_lo"])
            assert np.array_equal(t1.multiplicity[sl].astype(np.uint64), ref["multiplicity"])
            assert np.array_equal(t1.flags[sl], ref["flags"])
    finally:
        ctx.close()


@pytest.mark.parametrize("k,m,b1,b2,fo,s,ht", [
    (63, 14, 3, 2, False, 2, O.HASH_RK128),    # BASELINE configs[4] parameters
    (63, 14, 2, 1, True, 1, O.HASH_RK128),
    (31, 12, 2, 2, False, 2, O.HASH_RK128),    # -w rabin-karp128 at small k
    (63, 14, 3, 2, False, 2, O.HASH_SEQ),      # the reference's automatic choice for 32 < k <= 64: seq-hash u128
    (33, 12, 2, 2, False, 1, O.HASH_SEQ),
    (32, 12, 2, 1, False, 2, O.HASH_SEQ),      # even k: 64-bit key needs the wide path (no room for flag bits)
    (64, 14, 1, 1, True, 1, O.HASH_SEQ),
    (47, 13, 2, 3, True, 3, O.HASH_SEQ),
    (65, 16, 2, 1, False, 1, O.HASH_RK128),    # k > 64: the reference switches to rabin-karp128 (crates/api/src/utils.rs:17-26)
    (67, 17, 2, 2, False, 2, O.HASH_RK128),    # k - 2 > 64: window validity spans two bitmap words
    (96, 24, 2, 2, False, 2, O.HASH_RK128),
    (101, 12, 1, 2, False, 1, O.HASH_RK128),   # k - m = 89 m-mers per window
    (127, 32, 2, 1, True, 1, O.HASH_RK128),
    (128, 32, 1, 1, False, 1, O.HASH_RK128),
])
def test_wide_keys(k, m, b1, b2, fo, s, ht):
    """128-bit key path: seq-hash u128 and rabin-karp128, forward-only and canonical."""
    G = _gpu()
    rng = np.random.default_rng(k * 131 + m + ht)
    seqs = _mixed_reads(rng, k, n=250)
    if k % 2 == 0 and not fo:
        # even k: self-complementary k-mers get special trimming in the reference (final_executor.rs:322-352);
        # table semantics are unaffected, inputs here simply may contain them
        pass
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2, forward_only=fo)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, forward_only=fo, min_multiplicity=s, hash_type=ht)
    try:
        assert st.n_superkmers == len(sk)
        assert _gpu_records(ctx, (1 << b1) + 1) == _sk_records(reads, sk, k)
        n = _check_tables(G, ctx, reads, sk, k, s, b1, b2, fo, ranges=[(0, 1 << b1), (1 << b1, 1)], hash_type=ht)
        assert n > 0
    finally:
        ctx.close()


@pytest.mark.parametrize("n_bases,ht", [(9000, O.HASH_RK128), (30000, O.HASH_SEQ), (120000, O.HASH_RK128)])
def test_wide_unit_size_paths(n_bases, ht):
    """Wide path unit classes: <= 3072 records (512-thread table), <= 6144 (1024-thread table), global-scratch table;
    survivor counts above and below the shared-memory sort capacity (2048)."""
    G = _gpu()
    rng = np.random.default_rng(n_bases)
    k, m, b1, b2 = 63, 14, 0, 0
    g = util.rand_seq(rng, n_bases // 3)
    seqs = [g, util.revcomp(g[: len(g) // 2]), g[len(g) // 4:], util.rand_seq(rng, n_bases // 6)]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    for s in (1, 2):
        ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s, hash_type=ht)
        try:
            assert st.n_superkmers == len(sk)
            _check_tables(G, ctx, reads, sk, k, s, b1, b2, hash_type=ht)
        finally:
            ctx.close()


@pytest.mark.parametrize("k,ht,s", [(63, O.HASH_RK128, 1), (63, O.HASH_RK128, 2), (41, O.HASH_SEQ, 2), (64, O.HASH_SEQ, 40),
                                    (97, O.HASH_RK128, 1), (128, O.HASH_RK128, 2)])
def test_wide_key_partitions_and_overflow_fallback(k, ht, s):
    """Wide path, units above the shared-table capacity: key partitions in HBM (k_partition_units128 ->
    k_merge_hash128<SRC_RECORDS>, survivors of all partitions contiguous in the unit's static region).  A tandem
    repeat makes some partitions overflow: that unit is redone by the global-table kernel.  Identical tables."""
    G = _gpu()
    rng = np.random.default_rng(177 + k)
    m, b1, b2 = 14, 1, 0
    motif = util.rand_seq(rng, 211)
    seqs = [motif * 400, util.rand_seq(rng, 70000), util.revcomp(motif * 150), util.rand_seq(rng, 40000)]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s, hash_type=ht)
    try:
        _, km = ctx.unit_sizes()
        assert (km > 6144).sum() >= 2
        _check_tables(G, ctx, reads, sk, k, s, b1, b2, hash_type=ht)
    finally:
        ctx.close()


@pytest.mark.parametrize("k,ht,ratio", [(63, O.HASH_RK128, "0.02"), (41, O.HASH_SEQ, "0.05"), (63, O.HASH_RK128, "0.5"), (97, O.HASH_RK128, "0.01")])
def test_wide_tables_sized_by_distinct_keys_retry_when_wrong(k, ht, ratio, monkeypatch):
    """Wide path: shared tables are sized by the distinct keys a unit is expected to hold.  Here the expectation is
    forced far too low on all-distinct random sequence, so the bounded probing gives up and the units come back through
    the retry list (global-table kernel); and forced moderately low, so that some units fit and some do not.  Same tables."""
    G = _gpu()
    monkeypatch.setenv("GGCAT_B200_DISTINCT_RATIO", ratio)
    monkeypatch.setenv("GGCAT_B200_WIDE_BY_KEYS", "1")      # opt-in routing (the default routes the wide path by records)
    rng = np.random.default_rng(1234 + k)
    m, b1, b2, s = 14, 1, 1, 1
    seqs = [util.rand_seq(rng, 30000), util.rand_seq(rng, 9000), util.rand_seq(rng, 2500)] + [util.rand_seq(rng, 400) for _ in range(30)]
    seqs += [seqs[1][:5000]] * 3          # a repeated stretch: some units really are below the expectation
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s, hash_type=ht)
    try:
        _, km = ctx.unit_sizes()
        assert (km > 6144).sum() >= 1 and (km < 3000).sum() >= 1
        _check_tables(G, ctx, reads, sk, k, s, b1, b2, hash_type=ht)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2, hash_type=ht)      # second merge of the same context
    finally:
        ctx.close()


def test_colored_big_units():
    """-c with units above the shared-table capacity (partition path of the wide kernels in MODE_COLOR)."""
    G = _gpu()
    rng = np.random.default_rng(4711)
    k, m, b1, b2, s = 31, 12, 1, 0, 1
    anc = util.rand_seq(rng, 30000)
    seqs, cols = [], []
    for c in range(4):
        g = bytearray(anc)
        for _ in range(40):
            g[int(rng.integers(0, len(g)))] = ord("ACGT"[int(rng.integers(0, 4))])
        seqs.append(bytes(g)); cols.append(c)
    reads = O.Reads.from_list(seqs, colors=cols)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets, reads.colors)], b1, b2, k, m, min_multiplicity=s, colors=True)
    try:
        _, km = ctx.unit_sizes()
        assert (km > 6144).sum() >= 2
        n = _check_tables(G, ctx, reads, sk, k, s, b1, b2, colors=True)
        assert n > 0
    finally:
        ctx.close()


@pytest.mark.parametrize("k,m,b1,b2,s,fo", [(31, 12, 2, 2, 1, False), (31, 12, 1, 1, 2, False), (21, 10, 2, 1, 1, True),
                                             (41, 13, 1, 1, 1, False)])
def test_colored_build(k, m, b1, b2, s, fo):
    """-c: per k-mer the sorted-unique set of colours of all its occurrences (BASELINE configs[2] semantics at small
    scale: genomes sharing mutated segments, colour = input index)."""
    G = _gpu()
    rng = np.random.default_rng(k + 7 * s)
    anc = bytearray(util.rand_seq(rng, 6000))
    seqs, cols = [], []
    for c in range(9):
        g = bytearray(util.rand_seq(rng, 6000))
        for a in range(0, 6000, 1500):
            seg = bytearray(anc[a:a + 900])
            for _ in range(3):
                seg[int(rng.integers(0, len(seg)))] = ord("ACGT"[int(rng.integers(0, 4))])
            g[a:a + 900] = seg
        if c % 3 == 0:
            g = bytearray(util.revcomp(bytes(g)))
        if c == 4:
            g[100] = ord("N")
        seqs.append(bytes(g))
        cols.append(c if c != 7 else 1000003)   # colour ids need not be dense
        if c == 2:                               # two records of the same colour
            seqs.append(bytes(g[:2000]))
            cols.append(c)
    reads = O.Reads.from_list(seqs, colors=cols)
    sk, _ = O.bucketing(reads, k, m, b1, b2, forward_only=fo)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets, reads.colors)], b1, b2, k, m, forward_only=fo,
                                    min_multiplicity=s, colors=True)
    try:
        assert st.n_superkmers == len(sk)
        n = _check_tables(G, ctx, reads, sk, k, s, b1, b2, fo, colors=True)
        assert n > 0
    finally:
        ctx.close()


@pytest.mark.parametrize("k,m,b1,b2,s,fo", [(31, 12, 3, 2, 1, False), (31, 12, 2, 2, 2, False), (21, 10, 2, 2, 1, True),
                                             (41, 13, 2, 1, 1, False)])
def test_maximal_unitigs_from_gpu_tables(k, m, b1, b2, s, fo):
    """North-star check 2: maximal unitigs built by the reference's consumer logic (oracle restatement of
    hashmap.rs compute_unitigs per unit + join of open ends) from the GPU tables equal the unitigs of the
    independently counted global k-mer set: same unitig count, length multiset and canonical k-mer set."""
    G = _gpu()
    rng = np.random.default_rng(k * 3 + s)
    g = util.rand_seq(rng, 20000)
    cyc = util.rand_seq(rng, 400)
    seqs = [g, util.revcomp(g[3000:9000]), g[8000:15000], g[:700] + g[1500:3000], util.rand_seq(rng, 900),
            g[15000:16000] + g[100:1100], cyc + cyc[:k], b"ACACACACAC" * 12, g[4000:4100] + b"N" + g[4101:4300]]
    if s == 2:
        seqs = seqs + seqs[:5]
    reads = O.Reads.from_list(seqs)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, forward_only=fo, min_multiplicity=s)
    try:
        tab = ctx.merge_bucket_range(0, (1 << b1) + 1)
    finally:
        ctx.close()
    A = O.unitigs_from_tables(tab.keys_lo, tab.keys_hi, tab.count_flags, tab.unit_offsets, k, fo)
    nv, _ = O.naive_count(reads, k, O.HASH_SEQ, fo)
    nv = nv[nv["count"] >= s]
    B = O.unitigs_from_tables(nv["key_lo"], nv["key_hi"], nv["count"].astype(np.uint32), np.array([0, len(nv)], np.uint64), k, fo)
    assert A["n_unitigs"] == B["n_unitigs"] and A["n_partial"] > B["n_partial"]
    assert np.array_equal(A["lengths"], B["lengths"])
    assert np.array_equal(A["kmers_lo"], nv["key_lo"]) and np.array_equal(A["kmers_hi"], nv["key_hi"])


def test_c1_maximal_unitigs(golden_dir):
    """BASELINE configs[0] end to end: unitigs of sal1+sal2+sal3 (k=31 -s 1) from the GPU tables vs the global set."""
    G = _gpu()
    reads = O.Reads.from_list(util.c1_records())
    k, m, b1, b2, s = 31, 12, 2, 6, 1
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        tab = ctx.merge_bucket_range(0, (1 << b1) + 1)
    finally:
        ctx.close()
    A = O.unitigs_from_tables(tab.keys_lo, tab.keys_hi, tab.count_flags, tab.unit_offsets, k)
    nv, _ = O.naive_count(reads, k)
    B = O.unitigs_from_tables(nv["key_lo"], nv["key_hi"], nv["count"].astype(np.uint32), np.array([0, len(nv)], np.uint64), k)
    gold = json.loads((golden_dir / "c1_golden.json").read_text())
    assert A["n_unitigs"] == B["n_unitigs"] == gold["n_unitigs"]
    assert np.array_equal(A["lengths"], B["lengths"])
    assert np.array_equal(A["kmers_lo"], nv["key_lo"])


def test_c5_shape_rk128_properties():
    """BASELINE configs[4] at reduced genome size (same read shape and parameters: 150 bp reads at 30x, k=63 m=14 -s 2,
    rabin-karp128): size-independent properties on the whole table + sampled units against the oracle."""
    G = _gpu()
    from ggcat_b200 import synth

    k, m, s = 63, 14, 2
    g = synth.genome_codes(0xC5, 400_000)
    r = synth.simulate_reads(g, 80_000, 150, 0.0, 0xC5 + 1)
    data, offsets = synth.reads_to_ascii_batch(r)
    b1, b2 = 5, 4
    ctx, st = G.minimizer_bucketing([(data, offsets)], b1, b2, k, m, min_multiplicity=s, hash_type=O.HASH_RK128)
    try:
        n_reads = offsets.size - 1
        assert st.n_kmers == n_reads * (150 - k + 1) + (st.n_superkmers - n_reads)
        nb = (1 << b1) + 1
        t1 = ctx.merge_bucket_range(0, nb)
        assert t1.keys_hi is not None and t1.total_kmers == st.n_kmers
        assert (t1.multiplicity >= s).all()
        # strictly increasing u128 keys inside every unit
        hi, lo = t1.keys_hi, t1.keys_lo
        le = (hi[1:] < hi[:-1]) | ((hi[1:] == hi[:-1]) & (lo[1:] <= lo[:-1]))
        assert np.isin(np.nonzero(le)[0] + 1, t1.unit_offsets).all()
        t2 = ctx.merge_bucket_range(0, nb)
        assert np.array_equal(t1.keys_lo, t2.keys_lo) and np.array_equal(t1.keys_hi, t2.keys_hi)
        assert np.array_equal(t1.count_flags, t2.count_flags)
        reads = O.Reads(data, offsets)
        sk, _ = O.bucketing(reads, k, m, b1, b2)
        assert st.n_superkmers == len(sk)
        rng = np.random.default_rng(5)
        for u in rng.choice(nb << b2, 12, replace=False):
            ref, _, _ = O.merge_unit(reads, sk, int(u) >> b2, int(u) & ((1 << b2) - 1), k, s, O.HASH_RK128)
            ref = ref[ref["kept"] == 1]
            sl = t1.unit_slice(int(u))
            assert np.array_equal(t1.keys_lo[sl], ref["key_lo"]) and np.array_equal(t1.keys_hi[sl], ref["key_hi"])
            assert np.array_equal(t1.multiplicity[sl].astype(np.uint64), ref["multiplicity"])
            assert np.array_equal(t1.flags[sl], ref["flags"])
    finally:
        ctx.close()


def test_c3_shape_colored_properties():
    """BASELINE configs[2] at reduced size (genomes sharing mutated segments of an ancestor, one colour per genome,
    k=31 -s 1 -c): colour lists sorted-unique, a k-mer's colour set identical wherever it appears, colour sets of
    sampled units equal to the oracle's, and per-colour k-mer totals equal to an independent per-genome count."""
    G = _gpu()
    from ggcat_b200 import synth

    k, m, s = 31, 12, 1
    data, offsets, colors = synth.config_c3(n_genomes=12, genome_len=60_000, n_shared=6, shared_len=6_000, mut=0.002)
    b1, b2 = 4, 3
    ctx, st = G.minimizer_bucketing([(data, offsets, colors)], b1, b2, k, m, min_multiplicity=s, colors=True)
    try:
        nb = (1 << b1) + 1
        tab = ctx.merge_bucket_range(0, nb)
        co = tab.color_offsets
        assert int(co[-1]) == tab.colors.size and (np.diff(co.astype(np.int64)) >= 1).all()
        # sorted-unique colour lists
        inner = np.ones(tab.colors.size, bool)
        inner[co[:-1].astype(np.int64)] = False
        assert (tab.colors[1:][inner[1:]] > tab.colors[:-1][inner[1:]]).all()
        # independent check: the number of distinct canonical k-mers of genome c == entries whose list holds c
        # (boundary k-mers appear in two units: count distinct keys per colour)
        ent = np.repeat(np.arange(tab.n_entries), np.diff(co.astype(np.int64)))
        for c in (0, 5, 11):
            keys_c = np.unique(tab.keys_lo[ent[tab.colors == c]])
            gen = O.Reads(data[int(offsets[c]):int(offsets[c + 1])], np.array([0, int(offsets[c + 1] - offsets[c])], np.uint64))
            nv, _ = O.naive_count(gen, k)
            assert np.array_equal(keys_c, nv["key_lo"])
        reads = O.Reads(data, offsets, colors)
        sk, _ = O.bucketing(reads, k, m, b1, b2)
        rng = np.random.default_rng(3)
        for u in rng.choice(nb << b2, 6, replace=False):
            ref, rcols, _ = O.merge_unit(reads, sk, int(u) >> b2, int(u) & ((1 << b2) - 1), k, s, with_color=True)
            ref = ref[ref["kept"] == 1]
            sl = tab.unit_slice(int(u))
            assert np.array_equal(tab.keys_lo[sl], ref["key_lo"])
            assert np.array_equal(tab.multiplicity[sl].astype(np.uint64), ref["multiplicity"])
            got = tab.colors[int(co[sl.start]):int(co[sl.stop])]
            want = np.concatenate([rcols[int(o):int(o) + int(n)] for o, n in zip(ref["color_off"], ref["color_len"])]) if len(ref) else np.zeros(0, np.uint32)
            assert np.array_equal(got, want)
    finally:
        ctx.close()


@pytest.mark.parametrize("k,m,ht,batch", [(31, 12, O.HASH_SEQ, None), (31, 12, O.HASH_SEQ, 3000), (63, 14, O.HASH_RK128, 5000)])
def test_packed_input_equals_ascii_input(k, m, ht, batch, monkeypatch):
    """ggcat_b200_push_reads_packed (2-bit stream, offsets in bases, records of any length back to back) gives the same
    super-k-mers and tables as the ASCII push of the same reads -- also when the host path cuts the stream into batches
    that start in the middle of a byte -- and the device variant too."""
    G = _gpu()
    import torch

    from ggcat_b200 import synth

    if batch:
        monkeypatch.setenv("GGCAT_B200_HOST_BATCH", str(batch))
    rng = np.random.default_rng(99 + k)
    g = util.rand_seq(rng, 4000)
    seqs = []
    for _ in range(120):
        L = int(rng.integers(k - 3, 333))          # lengths not multiples of 4: records start anywhere inside a byte
        a = int(rng.integers(0, len(g) - L))
        r = g[a:a + L]
        seqs.append(util.revcomp(r) if rng.random() < 0.5 else r)
    seqs += [b"A" * 77, g[:k], b"", g[100:100 + k - 1], b"ACGT" * 40]
    reads = O.Reads.from_list(seqs)
    packed = synth.pack_2bit(reads.data)
    b1, b2, s = 2, 2, 1
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    want = _sk_records(reads, sk, k)
    for mode in ("host", "device"):
        ctx = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2, hash_type=ht))
        try:
            if mode == "host":
                ctx.push_reads_packed(packed, reads.offsets)
            else:
                dp = torch.from_numpy(np.concatenate([packed, np.zeros(8, np.uint8)])).cuda()
                do = torch.from_numpy(reads.offsets.view(np.int64)).cuda()
                ctx.push_reads_packed_device(dp.data_ptr(), do.data_ptr(), len(seqs), int(reads.data.size))
            st = ctx.finish_bucketing()
            assert st.n_superkmers == len(sk)
            assert _gpu_records(ctx, (1 << b1) + 1) == want
            _check_tables(G, ctx, reads, sk, k, s, b1, b2, hash_type=ht)
        finally:
            ctx.close()


def test_error_behaviour():
    G = _gpu()
    with pytest.raises(G.GgcatB200Error):
        G.GGCATB200(G.Params(k=3))
    with pytest.raises(G.GgcatB200Error):
        G.GGCATB200(G.Params(k=129))
    with pytest.raises(G.GgcatB200Error):
        G.GGCATB200(G.Params(k=65, hash_type=O.HASH_SEQ))     # seq-hash keys hold at most 64 bases
    G.GGCATB200(G.Params(k=65)).close()                       # AUTO -> rabin-karp128 above 64 (crates/api/src/utils.rs:17-26)
    with pytest.raises(G.GgcatB200Error):
        G.GGCATB200(G.Params(k=63, colors=True))
    with pytest.raises(G.GgcatB200Error):
        G.GGCATB200(G.Params(k=31, m=31))
    ctx = G.GGCATB200(G.Params(k=31, buckets_count_log=2, second_buckets_count_log=1))
    with pytest.raises(G.GgcatB200Error) as ei:
        ctx.merge_bucket_range(0, 1)  # before finish_bucketing
    assert ei.value.code == -3
    ctx.finish_bucketing()
    t = ctx.merge_bucket_range(0, 5)  # empty build: empty table, not an error
    assert t.n_entries == 0
    with pytest.raises(G.GgcatB200Error):
        ctx.merge_bucket_range(4, 2)
    ctx.close()


def test_valid_bases_matches_sequences_splitter():
    """stats.valid_bases = bases of the N-free segments of length >= k (SequencesSplitter::valid_bases), over several
    pushes, with N runs, short reads and lower-case / non-ACGT bytes."""
    G = _gpu()
    rng = np.random.default_rng(2024)
    k, m, b1, b2 = 31, 12, 3, 2
    seqs = _mixed_reads(rng, k, n=500)
    reads = O.Reads.from_list(seqs)
    sk, vb = O.bucketing(reads, k, m, b1, b2)
    third = len(seqs) // 3
    blocks = [O.Reads.from_list(seqs[:third]), O.Reads.from_list(seqs[third:2 * third]), O.Reads.from_list(seqs[2 * third:])]
    ctx, st = G.minimizer_bucketing([(r.data, r.offsets) for r in blocks], b1, b2, k, m)
    try:
        assert st.n_superkmers == len(sk)
        assert st.valid_bases == vb
        assert st.total_bases == sum(len(s) for s in seqs)
    finally:
        ctx.close()


@pytest.mark.parametrize("k,colors,s", [(64, False, 1), (64, False, 2), (48, True, 1)])
def test_all_ones_key_forward_only_poly_g(k, colors, s):
    """Forward-only seq-hash of k = 64 uses all 128 key bits: the poly-G k-mer IS the all-ones value that marks an empty
    table slot (with colours: k = 48 and colour 0xFFFFFFFF).  It must be counted like any other key, in the shared-table,
    the key-partition and the global-table paths (the slot stayed 'empty' and its count leaked)."""
    G = _gpu()
    rng = np.random.default_rng(64 + k)
    m, b1, b2 = 14, 2, 1
    g = util.rand_seq(rng, 3000)
    seqs = [b"G" * 300, g[:500] + b"G" * 150 + g[500:900], b"G" * 90, g, b"C" * 200, util.rand_seq(rng, 40000), b"G" * 4000]
    cols = np.array([0xFFFFFFFF, 1, 0xFFFFFFFF, 2, 3, 0xFFFFFFFF, 0xFFFFFFFF], np.uint32) if colors else None
    reads = O.Reads.from_list(seqs, colors=cols) if colors else O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2, forward_only=True)
    blk = (reads.data, reads.offsets, cols) if colors else (reads.data, reads.offsets)
    ctx, st = G.minimizer_bucketing([blk], b1, b2, k, m, forward_only=True, min_multiplicity=s, colors=colors)
    try:
        assert st.n_superkmers == len(sk)
        n = _check_tables(G, ctx, reads, sk, k, s, b1, b2, forward_only=True, colors=colors)
        tab = ctx.merge_bucket_range(0, (1 << b1) + 1)
        allones = (tab.keys_lo == np.uint64(0xFFFFFFFFFFFFFFFF)) & (tab.keys_hi == np.uint64((1 << (2 * k - 64)) - 1))
        assert allones.any(), "the poly-G k-mer must be in the table"
        assert n > 0
    finally:
        ctx.close()


def test_push_reads_from_several_threads():
    """push_reads is callable from many host threads (calls are serialised on the context): four threads push
    interleaved shares; the tables equal the oracle's on all reads (bucket contents are order-independent)."""
    G = _gpu()
    import threading

    rng = np.random.default_rng(8)
    k, m, b1, b2, s = 31, 12, 3, 2, 2
    seqs = _mixed_reads(rng, k, n=800)
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2))
    try:
        shares = [O.Reads.from_list(seqs[i::4]) for i in range(4)]
        errs = []

        def work(r):
            try:
                for a in range(0, r.n, 50):
                    b = min(r.n, a + 50)
                    ctx.push_reads(r.data[int(r.offsets[a]):int(r.offsets[b])], r.offsets[a:b + 1] - r.offsets[a])
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ts = [threading.Thread(target=work, args=(r,)) for r in shares]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs
        st = ctx.finish_bucketing()
        assert st.n_superkmers == len(sk)
        _check_tables(G, ctx, reads, sk, k, s, b1, b2)
    finally:
        ctx.close()


# ------------------------------------------------------------------ FASTA / FASTQ text on the device
def _parse_fasta_ref(text: bytes):
    """The reference's process_fasta (crates/io/src/sequences_reader.rs:106-179) restated line by line."""
    recs, cur = [], bytearray()
    for raw in text.split(b"\n"):
        line = raw[:-1] if raw.endswith(b"\r") else raw
        if line[:1] == b">":
            if cur:
                recs.append(bytes(cur))
            cur = bytearray()
        elif line[:1] == b";":
            continue
        else:
            cur += line
    if cur:
        recs.append(bytes(cur))
    return recs


def _parse_fastq_ref(text: bytes):
    """process_fastq (sequences_reader.rs:181-241): strict 4-line records."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines = lines[:-1]
    recs = []
    for i in range(1, len(lines), 4):
        ln = lines[i][:-1] if lines[i].endswith(b"\r") else lines[i]
        if ln:
            recs.append(bytes(ln))
    return recs


def _fasta_text(rng, n_rec, crlf=False):
    eol = b"\r\n" if crlf else b"\n"
    out = bytearray()
    if rng.random() < 0.5:
        out += util.rand_seq(rng, 37) + eol            # bases before the first ident line form a record too
    for r in range(n_rec):
        out += b">rec%d some description" % r + eol
        if rng.random() < 0.1:
            out += b";a comment line" + eol
        L = int(rng.integers(0, 700))
        s = util.rand_seq(rng, L, b"ACGTNacgt")
        w = int(rng.integers(20, 90))
        for a in range(0, L, w):
            out += s[a:a + w] + eol
        if rng.random() < 0.1:
            out += eol                                     # empty line inside / after a record
    if rng.random() < 0.5:
        out = out[:-len(eol)]                              # file without a trailing newline
    return bytes(out)


@pytest.mark.parametrize("seed,crlf", [(1, False), (2, True), (3, False)])
def test_device_tokenizer_fasta_matches_reader_semantics(seed, crlf):
    G = _gpu()
    rng = np.random.default_rng(seed)
    text = _fasta_text(rng, 400 if seed < 3 else 5000, crlf)
    want = _parse_fasta_ref(text)
    ctx = G.GGCATB200(G.Params(k=31, buckets_count_log=2, second_buckets_count_log=1))
    try:
        seq, off = ctx.tokenize(text, 0)
        got = [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(off.size - 1)]
        assert got == want
    finally:
        ctx.close()


def test_device_tokenizer_fastq_and_edge_cases():
    G = _gpu()
    rng = np.random.default_rng(9)
    recs = [util.rand_seq(rng, int(rng.integers(1, 300)), b"ACGTN") for _ in range(3000)]
    fq = b"".join(b"@r%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n" for i, r in enumerate(recs))
    ctx = G.GGCATB200(G.Params(k=31, buckets_count_log=2, second_buckets_count_log=1))
    try:
        for text in (fq, fq[:-1], fq.replace(b"\n", b"\r\n")):
            seq, off = ctx.tokenize(text, 1)
            got = [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(off.size - 1)]
            assert got == _parse_fastq_ref(text)
        for text in (b"", b">only a header\n", b">h1\n>h2\nACGT\n>h3\n", b"ACGT", b"\n\n>x\n\nAC\n\nGT\n"):
            seq, off = ctx.tokenize(text, 0)
            got = [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(off.size - 1)]
            assert got == _parse_fasta_ref(text), text
    finally:
        ctx.close()


def test_push_text_c1_files_equal_tokenised_push(golden_dir):
    """BASELINE configs[0] from the FASTA text itself: the three example files (multi-line FASTA, gunzipped on the host)
    pushed as raw text give the same super-k-mer counts and the same table digests as the tokenised records."""
    import gzip

    G = _gpu()
    k, m, b1, b2, s = 31, 12, 2, 6, 1
    recs = util.c1_records()
    reads = O.Reads.from_list(recs)
    ctx_a, st_a = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    ctx_b = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2))
    try:
        n = 0
        for f in ("sal1.fa.gz", "sal2.fa.gz", "sal3.fa.gz"):
            n += ctx_b.push_text(gzip.open(golden_dir / f, "rb").read(), 0)
        assert n == len(recs)
        st_b = ctx_b.finish_bucketing()
        assert (st_b.n_superkmers, st_b.n_kmers, st_b.valid_bases) == (st_a.n_superkmers, st_a.n_kmers, st_a.valid_bases)
        nb = (1 << b1) + 1
        ta, tb = ctx_a.merge_bucket_range(0, nb), ctx_b.merge_bucket_range(0, nb)
        assert np.array_equal(ta.keys_lo, tb.keys_lo) and np.array_equal(ta.count_flags, tb.count_flags)
        assert np.array_equal(ta.unit_offsets, tb.unit_offsets)
    finally:
        ctx_a.close(); ctx_b.close()


# ------------------------------------------------------------------ the reference's bucket file format
_parse_bucket_file = util.parse_bucket_file


@pytest.mark.parametrize("k,m", [(31, 12), (21, 10)])
def test_bucket_files_reference_format_roundtrip(k, m, tmp_path):
    """GPU phase 1 -> the reference's bucket files (what its phase 2 reads) -> records byte-equal to the oracle's
    serialisation (creads_utils.rs:374-434 restated in oracle/ggcat_oracle.c); then the files are imported into a fresh
    context (a CPU phase 1 feeding the GPU phase 2) and the merged tables are identical."""
    G = _gpu()
    rng = np.random.default_rng(31 + k)
    b1, b2, s = 2, 3, 2
    seqs = _mixed_reads(rng, k, n=500) + [util.rand_seq(rng, 30000)]
    reads = O.Reads.from_list(seqs)
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    nb = (1 << b1) + 1
    ctx2 = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2))
    try:
        total = 0
        for b in range(nb):
            path = tmp_path / f"bucket.{b}"
            n = ctx.write_bucket_file(b, path)
            total += n
            got = _parse_bucket_file(path.read_bytes(), k)
            want = {}
            for row in sk[sk["bucket"] == b]:
                want.setdefault(int(row["second_bucket"]), []).append(O.superkmer_record(reads, row, k)[1:])   # without the leading second-bucket byte
            assert set(got) == set(want)
            for sub in want:
                assert sorted(got[sub]) == sorted(want[sub]), f"bucket {b} sub-bucket {sub}"
            assert ctx2.import_bucket_file(b, path) == n
        assert total == st.n_superkmers
        st2 = ctx2.finish_bucketing()
        assert (st2.n_superkmers, st2.n_kmers) == (st.n_superkmers, st.n_kmers)
        ta, tb = ctx.merge_bucket_range(0, nb), ctx2.merge_bucket_range(0, nb)
        assert np.array_equal(ta.keys_lo, tb.keys_lo) and np.array_equal(ta.count_flags, tb.count_flags)
        assert np.array_equal(ta.unit_offsets, tb.unit_offsets)
        _check_tables(G, ctx2, reads, sk, k, s, b1, b2)
    finally:
        ctx.close(); ctx2.close()


# ------------------------------------------------------------------ partial unitigs on the device
def _unitig_strings(recs, bases):
    letters = np.frombuffer(b"ACTG", np.uint8)
    out = []
    for r in recs:
        n, w0 = int(r["len"]), int(r["word_offset"])
        words = bases[w0:w0 + (n + 15) // 16].astype(np.uint64)
        codes = ((words[:, None] >> (np.arange(16, dtype=np.uint64) * np.uint64(2))[None, :]) & np.uint64(3)).reshape(-1)[:n]
        out.append((int(r["unit"]), letters[codes.astype(np.int64)].tobytes(), int(r["flags"]), int(r["bucket"]), int(r["last_align"])))
    return out


@pytest.mark.parametrize("k,m,b1,b2,s,pbits", [(31, 12, 2, 2, 1, 3), (31, 12, 3, 2, 2, 5), (21, 10, 1, 1, 1, 4), (15, 9, 2, 1, 1, 3),
                                                (27, 11, 0, 0, 1, 6)])
def test_partial_unitigs_on_device_match_compute_unitigs(k, m, b1, b2, s, pbits):
    """Rows a12 / a13 / f1: the device builds, per merge unit, the partial unitigs HashMapUnitigsExtender::compute_unitigs
    would (same sequences in the same orientation, same open / closed ends, circular flag) and routes them as
    output_sequence would (result bucket, should_rc, HASH_ENDING / OTHER_END, last_align) -- compared as a multiset
    per unit with the oracle restatement run on the same GPU table; the partial unitigs cover every table entry once."""
    G = _gpu()
    rng = np.random.default_rng(1000 + k + b1)
    g = util.rand_seq(rng, 6000)
    seqs = _mixed_reads(rng, k, n=400) + [g, g[1000:3000], util.revcomp(g[2500:5200]), util.rand_seq(rng, 2500)]
    circ = util.rand_seq(rng, 300)
    seqs += [circ + circ[:k - 1], (circ + circ[:k - 1])[40:] + circ[:60]]         # a cycle in the de Bruijn graph
    seqs += [b"ACGT" * 40, b"A" * 200, b"AT" * 90]                                # low complexity, self-overlapping k-mers
    reads = O.Reads.from_list(seqs)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        nb = (1 << b1) + 1
        ne, _, _ = ctx.merge_bucket_range_device(0, nb)
        tab = ctx.read_device_table()
        recs, bases, nk = ctx.partial_unitigs(pbits)
        assert nk == ne == int(recs["n_kmers"].sum()), "every kept k-mer lies on exactly one partial unitig"
        assert (recs["len"] == recs["n_kmers"] + k - 1).all()
        got = _unitig_strings(recs, bases)
        want = O.partial_unitigs(tab.keys_lo, None, tab.count_flags, tab.unit_offsets, tab.first_unit, k, pbits)
        assert len(got) == len(want)
        # cycles: the reference opens them at their smallest k-mer -- so does the device; everything is compared verbatim
        assert sorted(got) == sorted(want)
        assert any(f & 4 for _, _, f, _, _ in want), "the input holds a circular unitig"
        assert any(f & 3 for _, _, f, _, _ in want) or b1 == 0
    finally:
        ctx.close()


def test_partial_unitigs_c1_join_to_maximal_unitigs(golden_dir):
    """BASELINE configs[0]: device-built partial unitigs of the three example files, joined at their open ends (the semantic
    join of oracle/ggcat_unitigs.c on the same table), give the maximal unitigs of the global flag-free build."""
    G = _gpu()
    k, m, b1, b2, s = 31, 12, 2, 6, 1
    recs_in = util.c1_records()
    reads = O.Reads.from_list(recs_in)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        nb = (1 << b1) + 1
        ne, _, _ = ctx.merge_bucket_range_device(0, nb)
        tab = ctx.read_device_table()
        recs, bases, nk = ctx.partial_unitigs(3)
        assert nk == ne
        got = _unitig_strings(recs, bases)
        want = O.partial_unitigs(tab.keys_lo, None, tab.count_flags, tab.unit_offsets, tab.first_unit, k, 3)
        assert sorted(got) == sorted(want)
        # open ends pair up: every open end k-mer (canonical) occurs exactly twice over all partial unitigs
        ends = {}
        for _, sq, f, _, _ in got:
            for bit, km in ((1, sq[:k]), (2, sq[-k:])):
                if f & bit:
                    c = min(km, util.revcomp(km))
                    ends[c] = ends.get(c, 0) + 1
        assert ends and set(ends.values()) == {2}
    finally:
        ctx.close()


def _canon_kmers_of(strings, k):
    """Sorted canonical k-mer keys (seq-hash u64: base i at bits 2i, A0 C1 T2 G3) of a list of sequences, with duplicates."""
    code = np.zeros(256, np.uint64)
    for ch, v in ((b"A", 0), (b"C", 1), (b"T", 2), (b"G", 3)):
        code[ch[0]] = v
    out = []
    for sq in strings:
        c = code[np.frombuffer(sq, np.uint8)]
        n = len(sq) - k + 1
        if n <= 0:
            continue
        win = np.lib.stride_tricks.sliding_window_view(c, k)
        sh = (np.arange(k, dtype=np.uint64) * np.uint64(2))[None, :]
        fw = (win << sh).sum(axis=1, dtype=np.uint64)
        rc = ((win ^ np.uint64(2)) << sh[:, ::-1]).sum(axis=1, dtype=np.uint64)
        out.append(np.minimum(fw, rc))
    return np.sort(np.concatenate(out)) if out else np.zeros(0, np.uint64)


@pytest.mark.parametrize("k,m,b1,b2,s", [(31, 12, 3, 2, 1), (31, 12, 2, 2, 2), (21, 10, 2, 2, 1), (27, 11, 4, 3, 1)])
def test_maximal_unitigs_joined_on_device(k, m, b1, b2, s):
    """SURVEY 8(f)-2 + north-star check 2 through the PRODUCT path: reads -> tables -> partial unitigs -> maximal unitigs,
    all on the device; the result equals the unitigs of the independently counted global k-mer set (oracle join of the
    flag-free global build): same unitig count, same length multiset, same canonical k-mer set, every k-mer once."""
    G = _gpu()
    rng = np.random.default_rng(k * 3 + s)
    g = util.rand_seq(rng, 20000)
    cyc = util.rand_seq(rng, 400)
    seqs = [g, util.revcomp(g[3000:9000]), g[8000:15000], g[:700] + g[1500:3000], util.rand_seq(rng, 900),
            g[15000:16000] + g[100:1100], cyc + cyc[:k], b"ACACACACAC" * 12, g[4000:4100] + b"N" + g[4101:4300]]
    if s == 2:
        seqs = seqs + seqs[:5]
    reads = O.Reads.from_list(seqs)
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        ne, _, _ = ctx.merge_bucket_range_device(0, (1 << b1) + 1)
        precs, _, pk = ctx.partial_unitigs(4)
        assert pk == ne
        recs, bases, nk = ctx.maximal_unitigs()
    finally:
        ctx.close()
    nv, _ = O.naive_count(reads, k)
    nv = nv[nv["count"] >= s]
    B = O.unitigs_from_tables(nv["key_lo"], nv["key_hi"], nv["count"].astype(np.uint32), np.array([0, len(nv)], np.uint64), k)
    assert len(recs) == B["n_unitigs"] and len(precs) > len(recs)
    assert np.array_equal(np.sort(recs["len"].astype(np.uint64)), B["lengths"])
    strings = [sq for _, sq, _, _, _ in _unitig_strings(recs, bases)]
    km = _canon_kmers_of(strings, k)
    assert nk == len(nv) == km.size
    assert np.array_equal(km, nv["key_lo"]), "the maximal unitigs hold every kept k-mer exactly once"
    assert (recs["flags"] & 4).any(), "the circular unitig is flagged"


def test_c1_maximal_unitigs_on_device(golden_dir):
    """BASELINE configs[0] end to end on the device: maximal unitigs of sal1+sal2+sal3 (k=31 -s 1) vs the golden unitig
    count and the global k-mer set."""
    G = _gpu()
    k, m, b1, b2, s = 31, 12, 2, 6, 1
    reads = O.Reads.from_list(util.c1_records())
    ctx, st = G.minimizer_bucketing([(reads.data, reads.offsets)], b1, b2, k, m, min_multiplicity=s)
    try:
        ne, _, _ = ctx.merge_bucket_range_device(0, (1 << b1) + 1)
        ctx.partial_unitigs(3)
        recs, bases, nk = ctx.maximal_unitigs()
    finally:
        ctx.close()
    gold = json.loads((golden_dir / "c1_golden.json").read_text())
    nv, _ = O.naive_count(reads, k)
    assert nk == len(nv)
    if "n_unitigs" in gold:
        assert len(recs) == gold["n_unitigs"]
    km = _canon_kmers_of([sq for _, sq, _, _, _ in _unitig_strings(recs, bases)], k)
    assert np.array_equal(km, nv["key_lo"])
