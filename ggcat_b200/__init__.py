"""ggcat_b200 -- B200-native k-mer counting front end for GGCAT (minimizer bucketing + k-mer merge).

The compute path is the CUDA library ggcat_b200/libggcat_b200.so behind the C ABI in
include/ggcat_b200.h; this package is the thin host-side mirror of the reference's phase
functions (see ggcat_b200.api).  There is no CPU fallback: importing works everywhere, computing
requires the built library and a CUDA device.
"""
from .api import (  # noqa: F401
    GGCATB200,
    BucketStats,
    KmerTable,
    Params,
    GgcatB200Error,
    bucket_counts,
    compute_best_m,
    kmers_merge,
    minimizer_bucketing,
)

__all__ = ["GGCATB200", "BucketStats", "KmerTable", "Params", "GgcatB200Error", "bucket_counts", "compute_best_m",
           "kmers_merge", "minimizer_bucketing"]
