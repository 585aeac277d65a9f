"""Per-kernel share table from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
ni, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in data:
    name = re.sub(r"\(.*", "", r[ni])[:100]
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"{len(data)} launches captured, total {tot:.0f} us (cold-cache, serialised)")
print("share   count   avg   kernel")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / tot * 100:6.2f}% {c:6d} {us / c:10.1f} us  {k}")
