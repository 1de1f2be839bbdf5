#!/usr/bin/env python
"""A/B of libvoxelrt builds on one GPU box: every variant under voxel-rs_b200/variants/<name>/libvoxelrt.so (made by
voxel-rs_b200/build.py build_variant, see tools/make_variants.py) is swapped in for one `bench.py` run in a fresh process; the
kernel times (CUDA events, L2 flushed before every launch) are what is compared. The default build is restored at the end.

    python tools/ab_kernels.py [--out gpurun_out/ab.jsonl] [--variants a,b,c] [-- extra bench.py flags]

Each output line: {"variant", "value", "ms_per_step", "kernel_ms", "split", "args"}. Never a bench line by itself.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "voxel-rs_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab.jsonl"))
    ap.add_argument("--variants", default="")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("bench_flags", nargs="*")
    args = ap.parse_args()
    vdir = os.path.join(PKG, "variants")
    names = [v for v in args.variants.split(",") if v] or sorted(os.listdir(vdir))
    live = os.path.join(PKG, "libvoxelrt.so")
    keep = live + ".default"
    shutil.copy2(live, keep)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    try:
        for name in names + ["default"]:
            src = keep if name == "default" else os.path.join(vdir, name, "libvoxelrt.so")
            if not os.path.exists(src):
                print(f"{name}: no build", flush=True)
                continue
            shutil.copy(src, live)
            os.utime(live, None)   # newer than the sources: graft.build() must not rebuild over the variant
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(args.steps), "--warmup", str(args.warmup), "--skip-cpu",
                   "--skip-e2e"] + args.bench_flags
            t0 = time.time()
            r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, VOXELRT_AB_VARIANT=name))
            line = None
            for l in r.stdout.splitlines():
                if l.startswith("{"):
                    line = json.loads(l)
            if line is None:
                print(f"{name}: FAILED rc={r.returncode}\n{r.stderr[-2000:]}", flush=True)
                continue
            rf = line["roofline"]
            rec = {"variant": name, "value": round(line["value"], 1), "ms_per_step": round(line["ms_per_step"], 4),
                   "kernel_ms": round(rf["kernel_ms"], 4), "split": rf["kernel_ms_split"], "args": args.bench_flags,
                   "clocks": line.get("clocks"), "wall_s": round(time.time() - t0, 1)}
            print(json.dumps(rec), flush=True)
            with open(args.out, "a") as f:
                f.write(json.dumps(rec) + "\n")
    finally:
        shutil.copy2(keep, live)
        os.utime(live, None)
        os.remove(keep)


if __name__ == "__main__":
    main()
