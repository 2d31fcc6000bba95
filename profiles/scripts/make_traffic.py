#!/usr/bin/env python3
"""profiles/traffic.json from the per-kernel sums of one full cfg2 step under ncu (launch_summary.py's JSON).
usage: make_traffic.py kernels.json launches.csv-name reads out.json"""
import json, sys
k = json.load(open(sys.argv[1]))
src, reads, out = sys.argv[2], float(sys.argv[3]), sys.argv[4]
k = {n: a for n, a in k.items() if n != "k_synth"}
per = {n: round((a["dram_read"] + a["dram_write"]) / reads, 1) for n, a in k.items()}
tj = {"source": f"{src} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one full cfg2 step, "
                f"{int(reads)} reads): DRAM bytes of EVERY launch of the step (parser, extractor, both bin levels, aggregation, growth, coverage compaction), "
                "summed, divided by the reads of the step; the input generator k_synth is excluded",
      "step_dram_bytes_per_read": round(sum(per.values()), 1),
      "per_kernel_dram_bytes_per_read": per,
      "per_kernel_ms_under_ncu": {n: a["ms"] for n, a in k.items()},
      "note": f"ncu dram__bytes_read.sum + dram__bytes_write.sum of every launch of one cfg2 step ({src}), per read x reads per step; NOT measured in the bench run itself"}
json.dump(tj, open(out, "w"), indent=1)
print("step DRAM bytes per read:", tj["step_dram_bytes_per_read"])
