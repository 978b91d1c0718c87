"""SASS census of the built library: per kernel, counts of the mnemonics that prove how it talks to memory
(TMA bulk copies `UBLKCP`, mbarrier `SYNCS`, shared / global atomics, 128-bit CAS, block barriers, tensor-core ops),
plus a short excerpt around the first occurrence of each proof mnemonic.
Usage: python profiles/sass_census.py [libggcat_b200.so] > profiles/sass_rNN.txt   (cuobjdump must be on PATH)"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PROOF = ["UBLKCP", "SYNCS", "ATOMS.CAS.128", "ATOMS.CAS.64", "ATOMS.CAS", "ATOMS", "ATOMG", "RED", "BAR.SYNC", "MATCH", "REDUX",
         "SHFL", "LDG.E.128", "STG.E.128", "LDS.128", "HMMA", "UTCMMA", "IMMA"]
EXCERPT = ["UBLKCP", "SYNCS", "ATOMS.CAS.128", "ATOMS.CAS.64"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o.replace("(anonymous namespace)::", "")).replace("ggb::", "").replace("void ", "") for o in out]


def main(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    arch = next((l.strip() for l in txt if l.strip().startswith("arch =")), "arch = ?")
    funcs = OrderedDict()
    cur = None
    for l in txt:
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4}\*/", l):
            funcs[cur].append(l.rstrip())
    names = demangle(list(funcs))
    print(f"# SASS census of {Path(lib).name} ({arch}); {len(funcs)} kernels; made by profiles/sass_census.py")
    print("# columns: instructions, then counts of the proof mnemonics (substring match on the opcode)")
    print()
    hdr = ["kernel", "insts"] + PROOF
    rows = []
    for (mang, lines), name in zip(funcs.items(), names):
        ops = []
        for l in lines:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops.append(m.group(1))
        c = Counter()
        for o in ops:
            for p in PROOF:
                if o.startswith(p):
                    c[p] += 1
        rows.append([name, len(ops)] + [c[p] for p in PROOF])
    w0 = max(len(r[0]) for r in rows)
    print(" ".join([hdr[0].ljust(w0)] + [h.rjust(max(len(h), 5)) for h in hdr[1:]]))
    for r in rows:
        print(" ".join([r[0].ljust(w0)] + [str(v).rjust(max(len(h), 5)) for v, h in zip(r[1:], hdr[1:])]))
    tot = Counter()
    for r in rows:
        for p, v in zip(PROOF, r[2:]):
            tot[p] += v
    print()
    print("# totals: " + ", ".join(f"{p}={tot[p]}" for p in PROOF))
    print("# tensor-core ops (HMMA / UTCMMA / IMMA) are expected to be 0: nothing on this path is a dense contraction")
    print()
    for p in EXCERPT:
        for (mang, lines), name in zip(funcs.items(), names):
            idx = next((i for i, l in enumerate(lines) if re.search(r"\s" + re.escape(p) + r"[ .]", l)), None)
            if idx is None:
                continue
            print(f"## first {p} in {name}")
            for l in lines[max(0, idx - 3): idx + 4]:
                print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).strip())
            print()
            if p in ("UBLKCP", "ATOMS.CAS.128"):
                continue      # one excerpt per kernel for the rare ones
            break


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "ggcat_b200" / "libggcat_b200.so"))
