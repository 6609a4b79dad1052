"""Top stall sites of one kernel in an .ncu-rep (SASS level, from `ncu --set full --import-source on`).

  python tools/ncu_stalls.py gpurun_out/prof.ncu-rep <kernel regex> [launch-skip] [top-N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{pat}",
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    print(lines[0][:160])
    rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rd[0]
    i_src, i_smp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    rows = []
    total = 0
    reason_tot = {}
    for n, r in enumerate(rd[1:]):
        if len(r) <= i_smp:
            continue
        if not r[i_smp].strip().isdigit():
            continue
        s = int(r[i_smp] or 0)
        total += s
        reasons = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:3]
        for i, h in stall_cols:
            reason_tot[h] = reason_tot.get(h, 0) + int(r[i] or 0)
        rows.append((s, n, r[i_src].strip(), r[i_exec], reasons))
    print(f"total samples {total}")
    print("by reason:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(reason_tot.items(), key=lambda x: -x[1]) if v))
    for s, n, src, ex, reasons in sorted(rows, reverse=True)[:top]:
        rs = " ".join(f"{h[6:]}:{v}" for v, h in reasons if v)
        print(f"{s:6d} {100.0 * s / max(total, 1):5.1f}%  #{n:<5d} x{ex:<8s} {src[:70]:70s} {rs}")


if __name__ == "__main__":
    main()
