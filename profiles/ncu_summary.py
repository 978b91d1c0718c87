"""Print the judged metrics of every kernel in an .ncu-rep (ncu --set full) as a markdown table.
Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_ncu_full.md"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def main(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    iname = hdr.index("Kernel Name")
    cols = [hdr.index(w) for w in WANT if w in hdr]
    print("| metric | unit | " + " | ".join(f"`{r[iname].split('(')[0][-40:]}`" for r in rows[2:]) + " |")
    print("|---|---|" + "---:|" * (len(rows) - 2))
    for c in cols:
        print(f"| {hdr[c]} | {units[c]} | " + " | ".join(r[c] for r in rows[2:]) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
