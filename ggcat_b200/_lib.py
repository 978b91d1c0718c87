"""ctypes binding of include/ggcat_b200.h.  Fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libggcat_b200.so"


class ParamsC(C.Structure):
    _fields_ = [
        ("k", C.c_uint32), ("m", C.c_uint32), ("min_multiplicity", C.c_uint32), ("buckets_count_log", C.c_uint32),
        ("second_buckets_count_log", C.c_uint32), ("forward_only", C.c_uint32), ("hash_type", C.c_uint32),
        ("colors", C.c_uint32), ("device", C.c_int32), ("reserved", C.c_uint32 * 7),
    ]


class BucketStatsC(C.Structure):
    _fields_ = [
        ("total_bases", C.c_uint64), ("valid_bases", C.c_uint64), ("n_superkmers", C.c_uint64), ("n_kmers", C.c_uint64),
        ("payload_words", C.c_uint64), ("n_buckets", C.c_uint32), ("n_units", C.c_uint32),
    ]


class SuperkmerC(C.Structure):
    _fields_ = [
        ("payload_offset", C.c_uint64), ("len", C.c_uint32), ("color", C.c_uint32), ("bucket", C.c_uint16),
        ("minimizer_pos", C.c_uint16), ("second_bucket", C.c_uint8), ("flags", C.c_uint8), ("rc", C.c_uint8),
        ("pad", C.c_uint8),
    ]


class TableC(C.Structure):
    _fields_ = [
        ("n_entries", C.c_uint64), ("keys_lo", C.POINTER(C.c_uint64)), ("keys_hi", C.POINTER(C.c_uint64)),
        ("count_flags", C.POINTER(C.c_uint32)), ("first_unit", C.c_uint32), ("n_units", C.c_uint32),
        ("unit_offsets", C.POINTER(C.c_uint64)), ("color_offsets", C.POINTER(C.c_uint64)),
        ("colors", C.POINTER(C.c_uint32)), ("total_kmers", C.c_uint64), ("unique_kmers", C.c_uint64),
        ("src_kmers", C.POINTER(C.c_uint64)), ("src_kmer_words", C.c_uint32), ("reserved0", C.c_uint32),
        ("opaque", C.c_void_p),
    ]


class UnitigC(C.Structure):
    _fields_ = [("word_offset", C.c_uint64), ("len", C.c_uint32), ("unit", C.c_uint32), ("bucket", C.c_uint16),
                ("flags", C.c_uint8), ("last_align", C.c_uint8), ("n_kmers", C.c_uint32)]


class UnitigsC(C.Structure):
    _fields_ = [("n_unitigs", C.c_uint64), ("n_words", C.c_uint64), ("n_kmers", C.c_uint64), ("unitigs", C.POINTER(UnitigC)),
                ("bases", C.POINTER(C.c_uint32)), ("d_unitigs", C.c_void_p), ("d_bases", C.c_void_p)]


class ChunkSliceC(C.Structure):
    _fields_ = [
        ("n_superkmers", C.c_uint64), ("n_words", C.c_uint64), ("word_bias", C.c_uint64),
        ("d_descriptors", C.c_void_p), ("d_payload", C.c_void_p), ("d_unit_counts", C.c_void_p),
        ("d_unit_words", C.c_void_p), ("d_unit_kmers", C.c_void_p),
        ("h_unit_counts", C.c_void_p), ("h_unit_words", C.c_void_p), ("h_unit_kmers", C.c_void_p),
    ]


# every symbol include/ggcat_b200.h declares: (restype, argtypes)
_vp, _u32, _u64, _i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
SYMBOLS = {
    "ggcat_b200_last_error": (C.c_char_p, []),
    "ggcat_b200_abi_version": (_u32, []),
    "ggcat_b200_compute_best_m": (_u32, [_u32]),
    "ggcat_b200_bucket_counts": (None, [_u64, C.POINTER(_u32), C.POINTER(_u32)]),
    "ggcat_b200_create": (_i32, [C.POINTER(ParamsC), C.POINTER(_vp)]),
    "ggcat_b200_destroy": (None, [_vp]),
    "ggcat_b200_host_alloc": (_vp, [_u64]),
    "ggcat_b200_host_free": (None, [_vp]),
    "ggcat_b200_push_reads": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "ggcat_b200_push_reads_device": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp]),
    "ggcat_b200_push_reads_packed": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "ggcat_b200_push_reads_packed_device": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp]),
    "ggcat_b200_push_text": (_i32, [_vp, _vp, _u64, _i32, _u32, C.POINTER(_u64)]),
    "ggcat_b200_push_text_device": (_i32, [_vp, _vp, _u64, _i32, _u32, C.POINTER(_u64)]),
    "ggcat_b200_tokenize_device": (_i32, [_vp, _vp, _u64, _i32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_u64), C.POINTER(_u64)]),
    "ggcat_b200_write_bucket_file": (_i32, [_vp, _u32, C.c_char_p, C.POINTER(_u64)]),
    "ggcat_b200_import_bucket_file": (_i32, [_vp, _u32, C.c_char_p, C.POINTER(_u64)]),
    "ggcat_b200_finish_bucketing": (_i32, [_vp, C.POINTER(BucketStatsC)]),
    "ggcat_b200_unit_sizes": (_i32, [_vp, _vp, _vp]),
    "ggcat_b200_dump_superkmers": (_i32, [_vp, _u32, _vp, _u64, _vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "ggcat_b200_merge_bucket_range": (_i32, [_vp, _u32, _u32, C.POINTER(TableC)]),
    "ggcat_b200_release_table": (_i32, [_vp, C.POINTER(TableC)]),
    "ggcat_b200_merge_bucket_range_device": (_i32, [_vp, _u32, _u32, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "ggcat_b200_device_table": (_i32, [_vp, C.POINTER(TableC)]),
    "ggcat_b200_partial_unitigs": (_i32, [_vp, _u32, C.POINTER(UnitigsC)]),
    "ggcat_b200_maximal_unitigs": (_i32, [_vp, C.POINTER(UnitigsC)]),
    "ggcat_b200_reset": (_i32, [_vp]),
    "ggcat_b200_n_chunks": (_u32, [_vp]),
    "ggcat_b200_export_chunk_slice": (_i32, [_vp, _u32, _u32, _u32, C.POINTER(ChunkSliceC)]),
    "ggcat_b200_import_chunk_slice": (_i32, [_vp, _u32, _u32, C.POINTER(ChunkSliceC)]),
    "ggcat_b200_drop_local_chunks": (_i32, [_vp]),
    "ggcat_b200_owner_range": (_i32, [_u32, _u32, _u32, C.POINTER(_u32), C.POINTER(_u32)]),
    "ggcat_b200_peer_init": (_i32, [_vp, _u32, _u32, _u64, _vp]),
    "ggcat_b200_peer_connect": (_i32, [_vp, _vp]),
    "ggcat_b200_peer_exchange": (_i32, [_vp]),
    "ggcat_b200_peer_stats": (_i32, [_vp, _vp, _vp]),
    "ggcat_b200_stream": (_vp, [_vp]),
    "ggcat_b200_synchronize": (_i32, [_vp]),
    "ggcat_b200_set_timing": (_i32, [_vp, _i32]),
    "ggcat_b200_kernel_times": (_i32, [_vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(_u32), _u32, _i32]),
}

_lib = None


def load() -> C.CDLL:
    """Load libggcat_b200.so (built by __graft_entry__.build()).  No fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). ggcat_b200 has no CPU fallback."
        )
    # GGCAT_B200_LIB: another build of the same sources (kernel A/B experiments: profiles/ab_variants.py)
    lib = C.CDLL(os.environ.get("GGCAT_B200_LIB") or str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
