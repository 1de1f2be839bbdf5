#!/usr/bin/env python
"""Builds profiles/traffic.json (what bench.py reports as roofline.traffic / roofline.issue) from the per-capture JSON files that
summarize.py wrote:  python profiles/make_traffic.py r01_v6"""
import json
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_v6"
out = {"source": f"profiles/{tag}_frame_wavefront.md + {tag}_picker_16M.md (ncu --set full; one 4K primary+shadow frame of the reference's "
                 "generated world / one 16 Mi-ray picker batch)", "kernels": {}}
frame = json.load(open(os.path.join(here, f"{tag}_frame_wavefront.json")))
picker = json.load(open(os.path.join(here, f"{tag}_picker_16M.json")))
seen = set()
for k in frame + picker:
    name = k["kernel"].replace("void ", "").split("(")[0]
    if name in seen:      # the capture holds the frame's kernels once more from the counting launch: keep the first
        continue
    seen.add(name)
    out["kernels"][name] = {"dram_bytes": float(k["dram_bytes"]), "warp_instructions": float(k["smsp__inst_executed.sum"]),
                            "duration_us_under_ncu": float(k["gpu__time_duration.sum"]) * (1000.0 if float(k["gpu__time_duration.sum"]) < 50 else 1.0)}
fk = [v for n, v in out["kernels"].items() if "picker" not in n]
out["render_kernel_dram_bytes_per_launch"] = sum(v["dram_bytes"] for v in fk)
out["warp_instructions_per_frame"] = sum(v["warp_instructions"] for v in fk)
out["picker_kernel_dram_bytes_per_launch"] = sum(v["dram_bytes"] for n, v in out["kernels"].items() if "picker" in n)
json.dump(out, open(os.path.join(here, "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
