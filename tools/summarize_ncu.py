"""Turns ncu outputs brought back from the GPU box into the small text summaries kept under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r01.csv   > profiles/r01_launch_list.md
  python tools/summarize_ncu.py report   gpurun_out/gemm_r01.ncu-rep   > profiles/r01_gemm_i8_ncu.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def launches(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list ({path}): one profiled forward, 64 clips x 16 s, gpu__time_duration (cold-cache, serialised: compare shares)\n")
    print(f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches\n")
    print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {k} | {v[0]} | {v[1] / 1e3:.2f} | {v[1] / v[0]:.1f} | {v[1] / tot * 100:.1f}% |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    print(f"# ncu --set full summary ({path})\n")
    for row in rd[2:]:
        name = row[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "")
        print(f"## {name}  (ID {row[0]})\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"- {m} = {row[i]} {units[i]}")
        print()


def traffic(path, pattern="gemm_i8_tc_kernel"):
    """Average DRAM bytes (read + write) per launch of the dominant kernel -> the JSON bench.py reports as roofline.traffic."""
    import json
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, ids = 0.0, set()
    for r in rd:
        if pattern not in r["Kernel Name"]:
            continue
        if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r["Metric Value"].replace(",", "")) * mult[r["Metric Unit"]]
            ids.add(r["ID"])
    print(json.dumps({"kernel": pattern, "launches": len(ids), "dram_bytes_per_launch": tot / max(1, len(ids)),
                      "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the {len(ids)} launches of one forward ({path})"}))


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2])
