"""world_size-2 gloo tests (CPU) of the N>1 host logic: owner map, split planning, the all-to-all of
unit-sorted chunks and the word-bias rebasing.  Chunks are built from the oracle's super-k-mers in the
exact layout the CUDA path produces (include/ggcat_b200.h ggcat_b200_chunk_slice)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ggcat_b200.dist import ChunkArrays, OwnerMap, exchange_chunk, plan_splits
from oracle import oracle as O
from tests import util

K, M, B1, B2 = 31, 12, 3, 2


def _reads_for_rank(rank):
    rng = np.random.default_rng(100 + rank)
    g = util.rand_seq(np.random.default_rng(7), 6000)  # shared genome
    seqs = []
    for _ in range(150):
        a = int(rng.integers(0, len(g) - 150))
        r = g[a:a + 150]
        seqs.append(util.revcomp(r) if rng.random() < 0.5 else r)
    seqs.append(b"A" * 100)
    return O.Reads.from_list(seqs)


def chunk_from_oracle(reads, sk, b2, n_units):
    """Oracle super-k-mers -> (desc u32[n,4], payload u32[], unit_cnt, unit_words, unit_kmers), unit-sorted."""
    unit = (sk["bucket"].astype(np.int64) << b2) | sk["second_bucket"]
    order = np.argsort(unit, kind="stable")
    sk, unit = sk[order], unit[order]
    desc = np.zeros((len(sk), 4), np.uint32)
    words = []
    woff = 0
    for i, row in enumerate(sk):
        pb = O.superkmer_packed(reads, row)
        nw = (int(row["len"]) + 15) // 16
        pb = pb + b"\0" * (nw * 4 - len(pb))
        words.append(np.frombuffer(pb, "<u4"))
        meta = int(row["minimizer_pos"]) | (int(row["flags"]) << 16) | (int(row["rc"]) << 18) | (int(row["second_bucket"]) << 19)
        desc[i] = (woff, int(row["len"]), meta, 0)
        woff += nw
    payload = np.concatenate(words) if words else np.zeros(0, np.uint32)
    unit_cnt = np.bincount(unit, minlength=n_units).astype(np.int32)
    nwords = (sk["len"].astype(np.int64) + 15) // 16
    unit_words = np.bincount(unit, weights=nwords, minlength=n_units).astype(np.int32)
    unit_kmers = np.bincount(unit, weights=sk["len"].astype(np.int64) - K + 1, minlength=n_units).astype(np.int32)
    return desc, payload, unit_cnt, unit_words, unit_kmers


def _records_from_slice(desc_bytes, payload, unit_cnt, word_bias, first_unit, b2):
    desc = np.frombuffer(desc_bytes.numpy().tobytes(), "<u4").reshape(-1, 4)
    pw = payload.numpy().view(np.uint32)
    units = np.repeat(np.arange(first_unit, first_unit + len(unit_cnt)), unit_cnt.numpy())
    out = []
    for d, u in zip(desc, units):
        woff, ln, meta = int(d[0]) - word_bias, int(d[1]), int(d[2])
        nb = (ln + 3) // 4
        body = pw[woff:woff + (ln + 15) // 16].tobytes()[:nb]
        out.append((int(u) >> b2, (meta >> 19) & 0xFF, ln, (meta >> 16) & 3, (meta >> 18) & 1, meta & 0xFFFF, body))
    return sorted(out)


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owner = OwnerMap(B1, B2, world)
        n_units = ((1 << B1) + 1) << B2
        reads = _reads_for_rank(rank)
        sk, _ = O.bucketing(reads, K, M, B1, B2)
        desc, payload, ucnt, uwords, ukmers = chunk_from_oracle(reads, sk, B2, n_units)
        chunk = ChunkArrays(torch.from_numpy(desc.view(np.uint8).reshape(-1).copy()), torch.from_numpy(payload.view(np.int32).copy()),
                            torch.from_numpy(ucnt), torch.from_numpy(uwords), torch.from_numpy(ukmers))
        got = exchange_chunk(chunk, owner, rank)
        assert len(got) == world
        fu, nu = owner.unit_range(rank)
        fb, nb = owner.bucket_range(rank)
        for sl in got:
            # what the source rank's oracle says belongs to my units
            sreads = _reads_for_rank(sl.src)
            ssk, _ = O.bucketing(sreads, K, M, B1, B2)
            mine = ssk[(ssk["bucket"] >= fb) & (ssk["bucket"] < fb + nb)]
            exp = sorted((int(r["bucket"]), int(r["second_bucket"]), int(r["len"]), int(r["flags"]), int(r["rc"]),
                          int(r["minimizer_pos"]), O.superkmer_packed(sreads, r)) for r in mine)
            assert sl.n_sk == len(mine)
            assert int(sl.unit_cnt.sum()) == sl.n_sk and int(sl.unit_words.sum()) == sl.n_words
            assert int(sl.unit_kmers.sum()) == int((mine["len"].astype(np.int64) - K + 1).sum())
            rec = _records_from_slice(sl.desc, sl.payload[:sl.n_words], sl.unit_cnt, sl.word_bias, fu, B2)
            assert rec == exp, f"rank {rank} <- {sl.src}"
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_owner_map():
    for b1 in (2, 9, 10):
        for world in (1, 2, 3, 4, 8):
            if world > (1 << b1):
                continue
            om = OwnerMap(b1, 6, world)
            seen = []
            for r in range(world):
                fb, nb = om.bucket_range(r)
                seen += list(range(fb, fb + nb))
                for b in range(fb, fb + nb):
                    assert om.owner_of_bucket(b) == r
                fu, nu = om.unit_range(r)
                assert fu == fb << 6 and nu == nb << 6
            assert seen == list(range((1 << b1) + 1))  # every bucket incl. the duplicates bucket, exactly once
            assert om.owner_of_bucket(1 << b1) == world - 1


def test_plan_splits():
    om = OwnerMap(2, 1, 2)  # 5 buckets x 2 = 10 units; rank0 units 0..3, rank1 units 4..9
    cnt = np.arange(10, dtype=np.int32)
    words = np.arange(10, dtype=np.int32) * 3
    p = plan_splits(cnt, words, om)
    assert p.tolist() == [[0 + 1 + 2 + 3, 3 * 6, 0], [sum(range(4, 10)), 3 * sum(range(4, 10)), 18]]


def test_exchange_world2_gloo():
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port), nprocs=world, join=True)
