"""Per-kernel-name mean of gpu__time_duration from tools/ncu_times.sh output (skips the first launch of each name)."""
import collections, csv, io, sys
for path in sys.argv[1:]:
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(list)
    for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        agg[name + " grid=" + r.get("Grid Size", "?")].append(v)
    print(path)
    for k, v in agg.items():
        w = v[1:] if len(v) > 1 else v
        print(f"  {k:60s} n={len(v):3d}  mean {sum(w) / len(w):8.1f} us   min {min(w):8.1f}")
