#!/usr/bin/env python
"""Static look at the traversal loops in the built library (no GPU needed): for every trace kernel, the innermost backward-branch
loops of its SASS with their instruction mix. The frame kernels are instruction-issue bound, so the length of the
walk loop is the first-order proxy for their speed when a change cannot be measured right away.

    python tools/sass_loops.py [voxel-rs_b200/libvoxelrt.so] [kernel-name-filter]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "voxel-rs_b200", "libvoxelrt.so")
flt = sys.argv[2] if len(sys.argv) > 2 else "trace_"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
for fn in re.split(r"\n\s*Function : ", sass)[1:]:
    name = fn.split("\n", 1)[0].strip()
    if flt not in name:
        continue
    ins = [(int(a, 16), i) for a, i in re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", fn)]
    loops = set()
    for a, i in ins:
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)*(0x[0-9a-f]+)", i)
        if m and int(m.group(1), 16) < a:
            loops.add((int(m.group(1), 16), a))
    print(f"{demangle(name)}: {len(ins)} SASS instructions")
    walk = [l for l in loops if sum("FFMA" in i for ad, i in ins if l[0] <= ad <= l[1]) >= 6]   # the walk loops: 3 + 3 fused multiply-adds per step
    for t, a in sorted(walk, key=lambda l: l[1] - l[0])[:2]:
        body = [i for ad, i in ins if t <= ad <= a]
        mix = {k: sum(k in i for i in body) for k in ("LDG", "LDS", "STS", "FFMA", "FMNMX", "LOP3", "SHF", "POPC", "BRA", "BSSY", "VOTE")}
        print(f"    loop {t:#06x}..{a:#06x}: {len(body):4d} instructions  " + " ".join(f"{k}={v}" for k, v in mix.items() if v))
