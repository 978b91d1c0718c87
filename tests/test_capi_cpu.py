"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, host-only helpers work, and compute entry points fail loudly without a CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from ggcat_b200 import _lib

    return _lib.load()


def test_header_symbols_exported(lib):
    from ggcat_b200 import _lib

    header = (ROOT / "include" / "ggcat_b200.h").read_text()
    declared = set(re.findall(r"\b(ggcat_b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ggcat_b200.h but not exported"
    assert lib.ggcat_b200_abi_version() == int(re.search(r"#define GGCAT_B200_ABI_VERSION (\d+)", header).group(1))


def test_struct_layouts():
    from ggcat_b200 import _lib, api

    assert C.sizeof(_lib.ParamsC) == 64
    assert C.sizeof(_lib.SuperkmerC) == 24 == api.SUPERKMER_DTYPE.itemsize
    assert C.sizeof(_lib.BucketStatsC) == 48
    assert C.sizeof(_lib.ChunkSliceC) == 88


def test_host_helpers_match_oracle(lib):
    from oracle import oracle as O
    import ggcat_b200 as G

    for k in list(range(4, 70)) + [100, 255]:
        assert G.compute_best_m(k) == O.compute_best_m(k)
    for n in [0, 1000, 509_594, 165_000_000, 508_000_000, 8_000_000_000, 100_000_000_000, 10**13]:
        assert G.bucket_counts(n) == O.bucket_counts(n)


def test_no_cpu_fallback(lib):
    """Without a GPU the product path must refuse to run, not silently compute on the CPU."""
    import torch
    import ggcat_b200 as G

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(G.GgcatB200Error) as ei:
        G.GGCATB200(G.Params(k=31))
    assert ei.value.code == -2  # GGCAT_B200_ERR_CUDA


def test_product_never_imports_oracle():
    for p in (ROOT / "ggcat_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            assert "oracle" not in p.read_text().replace("no oracle", ""), f"{p} references the oracle"


def test_owner_range_matches_owner_map(lib):
    """ggcat_b200_owner_range (the C side of the peer exchange) and dist.OwnerMap (the NCCL/gloo side) agree."""
    from ggcat_b200.dist import OwnerMap

    for b1 in (2, 5, 9, 10, 13):
        for world in (1, 2, 3, 4, 8):
            if world > (1 << b1):
                continue
            om = OwnerMap(b1, 6, world)
            covered = 0
            for r in range(world):
                fb, nb = C.c_uint32(0), C.c_uint32(0)
                assert lib.ggcat_b200_owner_range(b1, r, world, C.byref(fb), C.byref(nb)) == 0
                assert (fb.value, nb.value) == om.bucket_range(r)
                covered += nb.value
            assert covered == (1 << b1) + 1
    assert lib.ggcat_b200_owner_range(2, 0, 8, None, None) < 0     # more ranks than buckets
    assert lib.ggcat_b200_owner_range(5, 4, 4, None, None) < 0     # rank out of range
