"""Regenerates tests/golden/c1_golden.json from the oracle (run from the repo root)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402
from tests import util  # noqa: E402


def main():
    recs = util.c1_records()
    reads = O.Reads.from_list(recs)
    k, m, b1, b2, s = 31, 12, 2, 6, 1
    sk, vb = O.bucketing(reads, k, m, b1, b2)
    out = {"k": k, "m": m, "b1": b1, "b2": b2, "s": s, "records": len(recs), "bases": int(reads.data.size),
           "valid_bases": int(vb), "n_superkmers": int(len(sk)), "buckets": {}}
    nv, tot = O.naive_count(reads, k)
    out["total_kmers"] = int(tot)
    out["distinct_kmers"] = int(len(nv))
    for b in range((1 << b1) + 1):
        tab, _, tk = O.merge_unit(reads, sk, b, -1, k, s)
        tab = tab[tab["kept"] == 1]
        out["buckets"][str(b)] = {
            "n_superkmers": int((sk["bucket"] == b).sum()),
            "n_kmer_occurrences": int(tk),
            "n_entries": int(len(tab)),
            "digest": util.table_digest(tab["key_lo"], None, tab["multiplicity"], tab["flags"]),
        }
    # maximal unitigs of the whole input (consumer restatement, oracle/ggcat_unitigs.c) from the global k-mer set
    U = O.unitigs_from_tables(nv["key_lo"], nv["key_hi"], nv["count"].astype(np.uint32), np.array([0, len(nv)], np.uint64), k)
    out["n_unitigs"] = int(U["n_unitigs"])
    out["unitig_bases"] = int(U["lengths"].sum())
    (ROOT / "tests" / "golden" / "c1_golden.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
