"""CPU tests of the product's bucket-FILE writer / parser (ggcat_b200/csrc/wire.hpp, host-only C++ inside the CUDA library):
the reference's PLAIN SingleGrouped container (parallel-processor-rs/src/buckets/writers/mod.rs:15-70,
lock_free_binary_writer.rs, readers/binary_reader.rs:120-190) and record serializer (crates/io/src/concurrent/temp_reads/
creads_utils.rs:374-434), against the oracle's record restatement and an independent Python reader.  No GPU involved; the
GPU -> file -> GPU round trip is tests/test_gpu_parity.py::test_bucket_files_reference_format_roundtrip."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from tests import util

HERE = Path(__file__).resolve().parent


@pytest.fixture(scope="module")
def wire(tmp_path_factory):
    so = tmp_path_factory.mktemp("wire") / "libwire_harness.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(HERE / "wire_harness.cpp")])
    lib = C.CDLL(str(so))
    lib.wire_build.restype = C.c_int
    lib.wire_parse.restype = C.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _records_of_bucket(reads, sk, bucket, k):
    rows = sk[sk["bucket"] == bucket]
    rows = rows[np.argsort(rows["second_bucket"], kind="stable")]
    packed = [O.superkmer_packed(reads, r) for r in rows]
    off = np.zeros(len(rows) + 1, np.uint64)
    off[1:] = np.cumsum([len(p) for p in packed])
    return rows, np.frombuffer(b"".join(packed) + b"\0" * 8, np.uint8).copy(), off


def _build(wire, k, rows, packed, off):
    n = len(rows)
    ln = np.ascontiguousarray(rows["len"], np.uint32); mp = np.ascontiguousarray(rows["minimizer_pos"], np.uint32)
    fl = np.ascontiguousarray(rows["flags"], np.uint32); sb = np.ascontiguousarray(rows["second_bucket"], np.uint32)
    out = np.zeros(int(off[-1]) + 16 * n + 4096, np.uint8)
    size = C.c_uint64(0)
    rc = wire.wire_build(C.c_uint32(k), C.c_uint32(n), _p(ln), _p(mp), _p(fl), _p(sb), _p(packed), _p(off), _p(out),
                         C.c_uint64(out.size), C.byref(size))
    assert rc == 0
    return out[: size.value].tobytes()


def _parse(wire, buf, k, cap):
    a = np.frombuffer(buf, np.uint8).copy()
    ln, mp, fl, sb = (np.zeros(cap, np.uint32) for _ in range(4))
    bo = np.zeros(cap, np.uint64)
    n = C.c_uint64(0)
    err = C.create_string_buffer(256)
    rc = wire.wire_parse(_p(a), C.c_uint64(a.size), C.c_uint32(k), _p(ln), _p(mp), _p(fl), _p(sb), _p(bo), C.c_uint64(cap), C.byref(n),
                         err, C.c_uint32(256))
    return rc, int(n.value), (ln, mp, fl, sb, bo), err.value.decode()


@pytest.mark.parametrize("k,m", [(31, 12), (21, 10), (63, 14), (97, 24)])
def test_bucket_file_bytes_follow_the_reference_format(wire, k, m):
    rng = np.random.default_rng(7 + k)
    g = util.rand_seq(rng, 6000)
    seqs = [g[a:a + int(rng.integers(k, 400))] for a in rng.integers(0, 5000, 150)] + [g, b"A" * 300, util.revcomp(g[100:900])]
    reads = O.Reads.from_list(seqs)
    b1, b2 = 2, 3
    sk, _ = O.bucketing(reads, k, m, b1, b2)
    total = 0
    for bucket in range((1 << b1) + 1):
        rows, packed, off = _records_of_bucket(reads, sk, bucket, k)
        buf = _build(wire, k, rows, packed, off)
        # an independent reader of the container (header, checkpoint index, one chunk per sub-bucket) ...
        got = util.parse_bucket_file(buf, k)
        want = {}
        for r in rows:      # ... and the oracle's restatement of the record serializer, minus its leading second-bucket byte
            want.setdefault(int(r["second_bucket"]), []).append(O.superkmer_record(reads, r, k)[1:])
        assert got == want
        # the product's own parser returns the same records
        rc, n, (ln, mp, fl, sb, bo), err = _parse(wire, buf, k, len(rows) + 1)
        assert rc == 0 and n == len(rows), err
        assert np.array_equal(ln[:n], rows["len"]) and np.array_equal(mp[:n], rows["minimizer_pos"])
        assert np.array_equal(fl[:n], rows["flags"]) and np.array_equal(sb[:n], rows["second_bucket"])
        for i in range(n):
            nb = (int(ln[i]) + 3) // 4
            assert buf[int(bo[i]):int(bo[i]) + nb] == packed[int(off[i]):int(off[i]) + nb].tobytes()
        total += n
    assert total == len(sk)


def test_bucket_file_parser_rejects_what_it_cannot_read(wire):
    k, m = 31, 12
    reads = O.Reads.from_list([util.rand_seq(np.random.default_rng(3), 2000)])
    sk, _ = O.bucketing(reads, k, m, 1, 2)
    rows, packed, off = _records_of_bucket(reads, sk, int(sk["bucket"][0]), k)
    buf = bytearray(_build(wire, k, rows, packed, off))
    for mutate, what in [(lambda b: b.__setitem__(slice(0, 16), b"CPLZ4_INTR_BKT_M"), "not a PLAIN"),
                         (lambda b: b.__setitem__(24, 0), "SingleGrouped"),
                         (lambda b: b.__setitem__(slice(16, 24), (len(b) + 5).to_bytes(8, "little")), "index offset")]:
        bad = bytearray(buf)
        mutate(bad)
        rc, _, _, err = _parse(wire, bytes(bad), k, len(rows) + 1)
        assert rc == -2 and what in err
    rc, n, _, _ = _parse(wire, bytes(buf), k, len(rows) + 1)
    assert rc == 0 and n == len(rows)
    # an empty bucket is a header, the first checkpoint and an index of one entry
    empty = _build(wire, k, rows[:0], packed, off[:1])
    assert util.parse_bucket_file(empty, k) == {}
    assert _parse(wire, empty, k, 1)[:2] == (0, 0)
