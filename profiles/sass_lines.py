"""Join an ncu SASS-page CSV with nvdisasm -g line info and aggregate samples / instructions per source line.

usage: python profiles/sass_lines.py <report.ncu-rep> <ncu-kernel-regex> <mangled-name-regex> [top_n] [launch_skip]
Needs ggcat_b200/libggcat_b200.so built from the same sources as the profiled run."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def disasm_lines(func_regex):
    tmp = Path(tempfile.mkdtemp())
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.environ.get("GGCAT_B200_PROF_LIB", str(ROOT / "ggcat_b200" / "libggcat_b200.so"))], cwd=tmp,
                          stdout=subprocess.DEVNULL)
    cubin = next(tmp.glob("*.cubin"))
    txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
    out, cur, active = [], None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*//-+ \.text\.(\S+)", ln)
        if m:
            active = re.search(func_regex, m.group(1)) is not None and not out
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    rep, kre, mre = sys.argv[1], sys.argv[2], sys.argv[3]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    skip = sys.argv[5] if len(sys.argv) > 5 else "0"
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip,
                             "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csvtxt.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    sass = []
    for r in rows[hi + 1:]:
        if len(r) <= iinst or not r[ia].startswith("0x"):
            break  # next kernel's block
        sass.append((int(r[ia], 16), float(r[isamp] or 0), float(r[iinst] or 0), r[1]))
    base = sass[0][0]
    dl = disasm_lines(mre)
    by_off = {o: l for o, l, _ in dl}
    agg = defaultdict(lambda: [0.0, 0.0])
    for addr, s, n, _ in sass:
        l = by_off.get(addr - base)
        agg[l][0] += s
        agg[l][1] += n
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    src_cache = {}
    print(f"kernel {kre}: {len(sass)} SASS instr, samples {ts:.0f}, warp-instructions {ti:.0f}")
    for l, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        text = ""
        if l:
            f = next((p for p in (Path(os.environ.get("GGCAT_B200_PROF_SRC", str(ROOT / "ggcat_b200" / "csrc")))).glob(l[0])), None)
            if f:
                src_cache.setdefault(f, f.read_text().splitlines())
                if l[1] - 1 < len(src_cache[f]):
                    text = src_cache[f][l[1] - 1].strip()
        print(f"{str(l):32s} samples {100 * s / ts:5.1f}%  inst {100 * n / ti:5.1f}%  | {text[:100]}")


if __name__ == "__main__":
    main()
