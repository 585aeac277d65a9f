"""Compact per-kernel summary of an ncu report (raw page): time, DRAM bytes, throughput %, occupancy."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
g = lambda r, n: r[hdr.index(n)] if n in hdr else "-"
seen = {}
for r in data:
    name = g(r, "Kernel Name")[:58]
    if name in seen:
        continue
    seen[name] = 1
    rd = float(g(r, "dram__bytes_read.sum").replace(",", "")); wr = float(g(r, "dram__bytes_write.sum").replace(",", ""))
    print(f"{name:58s} {g(r,'gpu__time_duration.sum'):>8s}us rd {rd:7.1f} wr {wr:7.1f} (units {rows[1][hdr.index('dram__bytes_read.sum')]}) "
          f"dram% {float(g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):5.1f} l1% {float(g(r,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed')):5.1f} "
          f"lts% {float(g(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed')):5.1f} issue% {float(g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active')):5.1f} "
          f"warps% {float(g(r,'sm__warps_active.avg.pct_of_peak_sustained_active')):5.1f} regs {g(r,'launch__registers_per_thread')} tensor% {float(g(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') or 0):4.1f}")
