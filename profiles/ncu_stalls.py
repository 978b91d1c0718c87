"""Stall reasons (warps per issue-active cycle) of every kernel in an .ncu-rep, as a markdown table, plus a JSON of the
dominant limiter per kernel family (bench.py reads profiles/r02_limiters.json for roofline.limiter).
Usage: python profiles/ncu_stalls.py gpurun_out/x.ncu-rep [limiters.json]"""
import csv
import json
import re
import subprocess
import sys

FAMILY = [(r"k_pack|k_mark", "k_pack+k_mark"), (r"k_windows", "k_windows"), (r"k_emit", "k_emit"), (r"k_scatter", "k_scatter"),
          (r"k_merge_tier", "k_merge_hash<smem>"), (r"k_finish", "k_gather_units")]


def main(rep, out_json=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    iname = hdr.index("Kernel Name")
    stall = [(i, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
             if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    extra = [hdr.index(x) for x in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
                                    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") if x in hdr]
    names = [r[iname].split("(")[0][-44:] for r in rows[2:]]
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---:|" * len(names))
    for i, label in stall:
        vals = [float(r[i] or 0) for r in rows[2:]]
        if max(vals) < 0.05:
            continue
        print(f"| stall: {label} | " + " | ".join(f"{v:.2f}" for v in vals) + " |")
    for i in extra:
        print(f"| {hdr[i]} | " + " | ".join(f"{float(r[i] or 0):.2f}" for r in rows[2:]) + " |")
    if out_json:
        lim = {}
        ia, idr = extra[0], extra[1]
        for r in rows[2:]:
            fam = next((f for pat, f in FAMILY if re.search(pat, r[iname])), None)
            if fam is None or fam in lim:
                continue
            top = max(((float(r[i] or 0), label) for i, label in stall if label not in ("selected",)), default=(0, "-"))
            lim[fam] = {"kind": "issue" if float(r[ia]) > 45 else "latency", "issue_active_pct": round(float(r[ia]), 1),
                        "dram_pct_of_peak": round(float(r[idr]), 1), "top_stall": top[1], "top_stall_warps_per_issue": round(top[0], 2),
                        "source": "profiles/r02z_ncu_full.md (ncu --set full, first launch of the family)"}
        json.dump(lim, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
