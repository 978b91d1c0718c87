"""GPU parity at the sizes BASELINE.json names (run on the B200 box: pytest -m gpu).

  C3  coloured build over 100 synthetic 5 Mbp genomes (k=31 -s 1 -c): sampled units against the oracle incl. colour sets,
      per-colour naive counts of whole genomes, size-independent properties of the whole table
  C5  >= 10 M reads x 150 bp of the 250 Mbp genome, k=63 m=14 -s 2 rabin-karp128: sampled units against the oracle
  C4  a human-scale-SHAPED slice (error-free 30x): few buckets so that every unit exceeds 1 M k-mer records, i.e. the
      key-partition path with hundreds of partitions per unit that the full 93 Gbases set takes; sampled units + totals
Sizes can be scaled down with GGCAT_TEST_SCALE (default 1.0) when iterating.
Everything goes through the C ABI; the oracle (oracle/ggcat_oracle.c) runs phase 1 on the whole input on the host cores
and phase 2 on the sampled units only."""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
SCALE = float(os.environ.get("GGCAT_TEST_SCALE", "1.0"))


def _gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g

    g.build()
    import ggcat_b200 as G

    return G


def _check_units(tab_units, units, reads, sk, k, s, b2, hash_type=O.HASH_SEQ, wide=False):
    n = 0
    for i, u in enumerate(units):
        ref, _, _ = O.merge_unit(reads, sk, int(u) >> b2, int(u) & ((1 << b2) - 1), k, s, hash_type)
        ref = ref[ref["kept"] == 1]
        sl = tab_units.unit_slice_at(i)
        assert np.array_equal(tab_units.keys_lo[sl], ref["key_lo"]), f"unit {u}: keys differ"
        if wide:
            assert np.array_equal(tab_units.keys_hi[sl], ref["key_hi"]), f"unit {u}: high key words differ"
        assert np.array_equal(tab_units.multiplicity[sl].astype(np.uint64), ref["multiplicity"]), f"unit {u}: counts"
        assert np.array_equal(tab_units.flags[sl], ref["flags"]), f"unit {u}: flags"
        if hash_type == O.HASH_RK128:   # saved-reads contract: the source bases of a sample of entries re-hash to their keys
            assert tab_units.src_kmers is not None
            for e in list(range(sl.start, sl.stop))[:: max(1, (sl.stop - sl.start) // 16)]:
                lo, hi, _ = O.kmer_hashes(tab_units.src_kmer(e, k), k, O.HASH_RK128, True)
                assert (int(lo[0]), int(hi[0])) == (int(tab_units.keys_lo[e]), int(tab_units.keys_hi[e])), f"unit {u} entry {e}: source bases"
        n += len(ref)
    return n


def test_c4_shape_units_above_one_million_records():
    """BASELINE configs[3] shape: error-free 150 bp reads at 30x, k=31 -s 2, with so few buckets that every unit holds
    more than 1 M k-mer records (the full 93 Gbases set has 1.1 M per unit with its 1024 x 64 units): the key-partition
    path with hundreds of partitions per unit, sized from the distinct/records ratio of the first merge."""
    G = _gpu()
    import torch

    from ggcat_b200 import synth

    k, m, s, b1, b2 = 31, 12, 2, 3, 2
    n_reads = int(360_000 * SCALE)
    dev = torch.device("cuda", 0)
    genome = synth.genome_codes_torch(0xC4, 5 * n_reads, dev)
    d_data = synth.simulate_reads_torch(genome, n_reads, 150, 0.0, 0xC4 + 1)
    d_off = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * 150
    data = d_data.cpu().numpy()
    offsets = d_off.cpu().numpy().view(np.uint64)
    ctx = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2))
    try:
        ctx.push_reads_device(d_data.data_ptr(), d_off.data_ptr(), n_reads, int(d_data.numel()))
        st = ctx.finish_bucketing()
        _, km = ctx.unit_sizes()
        nb = (1 << b1) + 1
        if SCALE >= 1.0:
            assert (km[: (nb - 1) << b2] > 1_000_000).all(), "every ordinary unit must exceed 1 M records"
        reads = O.Reads(data, offsets)
        sk, _ = O.bucketing(reads, k, m, b1, b2)
        assert st.n_superkmers == len(sk)
        rng = np.random.default_rng(4)
        units = sorted(int(u) for u in rng.choice((nb - 1) << b2, 5, replace=False)) + [(nb - 1) << b2]   # + a duplicates-bucket unit
        for it in range(2):     # second merge: partitions sized by the distinct/records ratio the first one measured
            ne, uq, tk = ctx.merge_bucket_range_device(0, nb)
            assert tk == st.n_kmers
            tab = ctx.read_device_table(units)
            assert tab.n_entries_total == ne
            _check_units(tab, units, reads, sk, k, s, b2)
        # whole table through the host path: identical entry count, sorted inside every unit
        t = ctx.merge_bucket_range(0, nb)
        assert t.n_entries == ne and t.total_kmers == st.n_kmers
        bad = np.nonzero(t.keys_lo[1:] <= t.keys_lo[:-1])[0] + 1
        assert np.isin(bad, t.unit_offsets).all()
    finally:
        ctx.close()


def test_c5_ten_million_reads_rk128_sampled_units():
    """BASELINE configs[4]: reads of the synthetic 250 Mbp genome, k=63 m=14 -s 2, rabin-karp128 (128-bit keys),
    10 M x 150 bp reads (1.5 Gbases; the full set is 50 M reads of the same generator), 1024(+1) x 64 units."""
    G = _gpu()
    import torch

    from ggcat_b200 import synth

    k, m, s = 63, 14, 2
    n_reads = int(10_000_000 * SCALE)
    genome_len = int(250_000_000 * SCALE)
    dev = torch.device("cuda", 0)
    genome = synth.genome_codes_torch(0xC5, genome_len, dev)
    d_data = synth.simulate_reads_torch(genome, n_reads, 150, 0.0, 0xC5 + 1)
    del genome
    d_off = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * 150
    b1, b2 = G.bucket_counts(int(50_000_000 * 165 * SCALE))    # bucket counts of the whole C5 input (1024 x 64)
    ctx = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2,
                               hash_type=O.HASH_RK128))
    try:
        per_push = 4_000_000
        for r0 in range(0, n_reads, per_push):
            r1 = min(n_reads, r0 + per_push)
            off = (d_off[r0:r1 + 1] - r0 * 150).contiguous()
            ctx.push_reads_device(d_data.data_ptr() + r0 * 150, off.data_ptr(), r1 - r0, (r1 - r0) * 150)
            ctx.synchronize()
        st = ctx.finish_bucketing()
        assert st.n_kmers == n_reads * (150 - k + 1) + (st.n_superkmers - n_reads)
        nb = (1 << b1) + 1
        ne, uq, tk = ctx.merge_bucket_range_device(0, nb)
        assert tk == st.n_kmers
        data = d_data.cpu().numpy()
        offsets = d_off.cpu().numpy().view(np.uint64)
        del d_data
        reads = O.Reads(data, offsets)
        sk, _ = O.bucketing(reads, k, m, b1, b2)
        assert st.n_superkmers == len(sk)
        rng = np.random.default_rng(55)
        units = sorted(int(u) for u in rng.choice(nb << b2, 16, replace=False))
        tab = ctx.read_device_table(units)
        assert tab.keys_hi is not None and tab.n_entries_total == ne
        n = _check_units(tab, units, reads, sk, k, s, b2, hash_type=O.HASH_RK128, wide=True)
        assert n > 0
    finally:
        ctx.close()


def test_c3_hundred_genomes_colored():
    """BASELINE configs[2]: coloured build (-c) over 100 synthetic 5 Mbp genomes with shared / mutated segments, k=31
    -s 1, 1024(+1) x 64 units, colour = genome index."""
    G = _gpu()
    from ggcat_b200 import synth

    k, m, s = 31, 12, 1
    n_genomes = max(4, int(100 * SCALE))
    data, offsets, colors = synth.config_c3(n_genomes=n_genomes)
    b1, b2 = G.bucket_counts(int(data.size * 1.016))     # FASTA bytes incl. line breaks (1024 x 64)
    ctx = G.GGCATB200(G.Params(k=k, m=m, min_multiplicity=s, buckets_count_log=b1, second_buckets_count_log=b2, colors=True))
    try:
        # a few genomes per push (one colour per record)
        per = 20
        for g0 in range(0, n_genomes, per):
            g1 = min(n_genomes, g0 + per)
            a, b = int(offsets[g0]), int(offsets[g1])
            ctx.push_reads(data[a:b], offsets[g0:g1 + 1] - offsets[g0], colors[g0:g1])
        st = ctx.finish_bucketing()
        nb = (1 << b1) + 1
        rng = np.random.default_rng(33)
        buckets = sorted(int(b) for b in rng.choice(nb - 1, 3, replace=False)) + [nb - 1]
        reads = O.Reads(data, offsets, colors)
        sk, _ = O.bucketing(reads, k, m, b1, b2)
        assert st.n_superkmers == len(sk)
        seen_per_color = {c: [] for c in (0, n_genomes - 1)}
        n_checked = 0
        for fb in buckets:     # whole first-level buckets through the host path (64 units each), every unit checked
            tab = ctx.merge_bucket_range(fb, 1)
            co = tab.color_offsets
            assert int(co[-1]) == tab.colors.size and (np.diff(co.astype(np.int64)) >= 1).all()
            inner = np.ones(tab.colors.size, bool)
            inner[co[:-1].astype(np.int64)] = False
            assert (tab.colors[1:][inner[1:]] > tab.colors[:-1][inner[1:]]).all(), "colour lists must be sorted-unique"
            for u in range(fb << b2, (fb + 1) << b2):
                ref, rcols, _ = O.merge_unit(reads, sk, u >> b2, u & ((1 << b2) - 1), k, s, with_color=True)
                ref = ref[ref["kept"] == 1]
                sl = tab.unit_slice(u)
                assert np.array_equal(tab.keys_lo[sl], ref["key_lo"]), f"unit {u}: keys"
                assert np.array_equal(tab.multiplicity[sl].astype(np.uint64), ref["multiplicity"]), f"unit {u}: counts"
                assert np.array_equal(tab.flags[sl], ref["flags"]), f"unit {u}: flags"
                got = tab.colors[int(co[sl.start]):int(co[sl.stop])]
                want = (np.concatenate([rcols[int(o):int(o) + int(n)] for o, n in zip(ref["color_off"], ref["color_len"])])
                        if len(ref) else np.zeros(0, np.uint32))
                assert np.array_equal(got, want), f"unit {u}: colour sets"
                n_checked += len(ref)
        assert n_checked > 0
        # independent per-colour check on two whole genomes: the distinct canonical k-mers of genome c == the keys whose
        # colour list holds c, over the WHOLE table
        tab = ctx.merge_bucket_range(0, nb)
        co = tab.color_offsets
        ent = np.repeat(np.arange(tab.n_entries), np.diff(co.astype(np.int64)))
        for c in seen_per_color:
            keys_c = np.unique(tab.keys_lo[ent[tab.colors == c]])
            gen = O.Reads(data[int(offsets[c]):int(offsets[c + 1])], np.array([0, int(offsets[c + 1] - offsets[c])], np.uint64))
            nv, _ = O.naive_count(gen, k)
            assert np.array_equal(keys_c, nv["key_lo"]), f"colour {c}: k-mer set differs from the naive count of its genome"
    finally:
        ctx.close()
