#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X`):
launches, summed duration, DRAM bytes.  usage: launch_summary.py launches.csv [out.json]"""
import csv, json, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).strip()
    a = agg.setdefault(name, {"launches": 0, "ms": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        a["launches"] += 1
        a["ms"] += v / {"ns": 1e6, "us": 1e3, "usecond": 1e3, "nsecond": 1e6, "ms": 1.0, "msecond": 1.0, "second": 1e-3}.get(u, 1e6)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        a["dram_read" if "read" in r[mi] else "dram_write"] += v * scale
tot = sum(a["ms"] for n, a in agg.items() if n != "k_synth")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    a["ms"] = round(a["ms"], 3)
    print(f"{n:40s} launches {a['launches']:5d}  ms {a['ms']:9.3f}  share {a['ms'] / tot if n != 'k_synth' else 0:6.3f}  dram GB r {a['dram_read'] / 1e9:8.2f} w {a['dram_write'] / 1e9:8.2f}")
print(f"total without the generator: {tot:.2f} ms, DRAM {sum(a['dram_read'] + a['dram_write'] for n, a in agg.items() if n != 'k_synth') / 1e9:.1f} GB")
if len(sys.argv) > 2:
    json.dump(agg, open(sys.argv[2], "w"), indent=1)
