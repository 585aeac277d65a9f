"""Per-kernel count of the Blackwell-specific SASS mnemonics in the shipped library (cuobjdump -sass):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy,
USETMAXREG = setmaxnreg, SYNCS = mbarrier.  Writes a table (stdout) for profiles/.
  python tools/sass_summary.py [path/to/lib.so]"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                           "gnnome_assembly_b200", "libgnnome_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
names = dict(zip(re.findall(r"Function : (\S+)", out), filt))
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "USETMAXREG", "SYNCS", "HMMA", "FFMA", "REDG", "ATOMG", "RED.E"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for p in pats:
            if op.startswith(p):
                counts[cur][p] += 1
print(f"{lib}: {len(counts)} kernels (sm_100a)")
print(f"{'kernel':110s} " + " ".join(f"{p:>8s}" for p in pats[:9]) + "     FFMA")
for k, c in counts.items():
    nm = re.sub(r"\(.*", "", names.get(k, k))
    nm = nm.replace("gg::", "").replace("void ", "")
    if not any(c[p] for p in pats[:9]) and c["FFMA"] < 50:
        continue
    print(f"{nm[:110]:110s} " + " ".join(f"{c[p]:8d}" for p in pats[:9]) + f" {c['FFMA']:8d}")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("TOTAL " + " ".join(f"{p}={tot[p]}" for p in pats if tot[p]))
