"""Extract per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, per launch) and duration
from ncu --set full reports into profiles/<name>.json; bench.py reports it as roofline.traffic."""
import csv, json, subprocess, sys
out = {}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = hdr.index
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in data:
        name = r[ci("Kernel Name")]
        rd = float(r[ci("dram__bytes_read.sum")].replace(",", "")) * scale[units[ci("dram__bytes_read.sum")]]
        wr = float(r[ci("dram__bytes_write.sum")].replace(",", "")) * scale[units[ci("dram__bytes_write.sum")]]
        dur = float(r[ci("gpu__time_duration.sum")].replace(",", ""))
        e = out.setdefault(name, {"launches": 0, "read": 0.0, "write": 0.0, "us": 0.0, "report": rep.split("/")[-1]})
        e["launches"] += 1; e["read"] += rd; e["write"] += wr; e["us"] += dur
res = {k: {"dram_bytes_per_launch": round((v["read"] + v["write"]) / v["launches"]),
           "dram_read": round(v["read"] / v["launches"]), "dram_write": round(v["write"] / v["launches"]),
           "ncu_us": round(v["us"] / v["launches"], 1), "launches": v["launches"], "report": v["report"]}
       for k, v in out.items()}
json.dump(res, open(sys.argv[1], "w"), indent=1)
for k, v in res.items():
    print(k[:70], v["dram_bytes_per_launch"] / 1e6, "MB", v["ncu_us"], "us")
