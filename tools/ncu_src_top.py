"""Summarise an `ncu --page source --csv` dump: top instructions by stall samples (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, h in enumerate(hi):
    hdr = rows[h]
    end = hi[k + 1] - 1 if k + 1 < len(hi) else len(rows)
    data = [r for r in rows[h + 1:end] if len(r) == len(hdr)]
    ci = hdr.index
    S, SRC = ci("# Samples"), ci("Source")
    stall = [i for i, x in enumerate(hdr) if x.startswith("stall_")]
    num = lambda x: float(x.replace(",", "") or 0) if x else 0.0
    tot = sum(num(r[S]) for r in data) or 1
    print(rows[h - 1][1][:110] if h else "", "| samples", int(tot), "| instrs", len(data))
    for r in sorted(data, key=lambda r: -num(r[S]))[:n]:
        st = sorted(((num(r[i]), hdr[i][6:]) for i in stall if r[i]), reverse=True)[:2]
        print(f"{int(num(r[S])):7d} {100 * num(r[S]) / tot:5.1f}%  {r[SRC][:64]:64s} {[(int(a), b) for a, b in st]}")
    W, WI = ci("L1 Wavefronts Shared"), ci("L1 Wavefronts Shared Ideal")
    print("  shared wavefronts", int(sum(num(r[W]) for r in data)), "ideal", int(sum(num(r[WI]) for r in data)))
