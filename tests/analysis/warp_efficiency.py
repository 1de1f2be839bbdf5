#!/usr/bin/env python
"""Offline look at warp-level lane utilisation of the frame kernels' walk loops (no GPU): the oracle reports the loop iterations of
every pixel's primary and shadow ray; this script lays the rays out the way the kernels do (8x4 warp tiles in 32x16 macro blocks for
trace_primary_kernel, the strip-ordered compacted list in runs of 32 for trace_shadow_kernel) and computes, for "finish all 32 rays,
then refill" (refill threshold 1):

    utilisation = sum of ray iterations / (32 x sum over warps of the longest ray of the warp)

It is the counterpart of ncu's "threads per instruction" inside the walk loop (profiles/r01_v6_frame_wavefront.md: 27 / 20-22 of 32;
ncu also sees the divergence between PUSH / ADVANCE / POP inside an iteration, which this does not model). Then it re-orders the
shadow list in ways that do not touch any ray's arithmetic, to see what a binning pass in front of trace_shadow_kernel could win:
sorted by the true iteration count (the bound), and binned by cheap predictors known when the entry is appended.

    python tests/analysis/warp_efficiency.py [--width 3840 --height 2160] [--radius 20]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft   # noqa: E402
import bench                      # noqa: E402


def utilisation(steps):
    """steps: 1-D array in issue order; warps take 32 consecutive entries."""
    n = len(steps)
    pad = (-n) % 32
    s = np.concatenate([steps, np.zeros(pad, steps.dtype)]).reshape(-1, 32).astype(np.int64)
    return float(s.sum()) / float(32 * s.max(axis=1).sum()), int(s.max(axis=1).sum())


def tile_order(w, h):
    """Pixel index (y * w + x) in the order the kernels enumerate slots: macro block (32x16) row-major, 4 strips (32x4) each, pixel p of a
    strip = lane p % 32 of tile p / 32 (strip_pixel in kernels.cuh). Pixels outside the frame are dropped."""
    mx, my = (w + 31) // 32, (h + 15) // 16
    macro = np.arange(mx * my)
    strip = np.arange(4)
    p = np.arange(128)
    x = (macro[:, None, None] % mx) * 32 + (p[None, None, :] >> 5) * 8 + (p[None, None, :] & 7) + 0 * strip[None, :, None]
    y = (macro[:, None, None] // mx) * 16 + strip[None, :, None] * 4 + ((p[None, None, :] >> 3) & 3)
    x, y = x.reshape(-1), y.reshape(-1)
    keep = (x < w) & (y < h)
    return (y * w + x), keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--radius", type=int, default=20)
    ap.add_argument("--terrain", default="reference")
    ap.add_argument("--no-lod", action="store_true")
    ap.add_argument("--no-shadows", action="store_true")
    ap.add_argument("--format", default="esvo")
    args = ap.parse_args()
    graft.build()
    pkg, ora = graft.load_pkg(), graft.load_oracle()
    world, _ = bench.build_world(pkg, args)
    reg = pkg.content_registry(pkg.load_atlas())
    vxp = bench.frame_params(pkg, world, args)
    tex, mips = reg.textures()
    scene = ora.Scene(world.gpu_buffer(), reg.materials().tobytes(), tex, mips, fmt=world.fmt)
    W, H = args.width, args.height
    prim, shad = scene.render_steps(vxp, W, H)
    order, keep = tile_order(W, H)
    # primary: the slot order including the dead lanes of ragged tiles (they idle)
    ps = np.where(keep, prim.reshape(-1)[np.where(keep, order, 0)], 0)
    u, crit = utilisation(ps)
    print(f"frame {W}x{H}: {prim.size} primary rays, {int((shad > 0).sum())} shadow rays, iterations {int(prim.sum())} + {int(shad.sum())}")
    print(f"trace_primary_kernel  8x4 warp tiles                       lanes busy {32 * u:5.2f} / 32   warp-iterations {crit / 1e6:8.2f} M")
    rows = ps.reshape(-1, 32)
    for name, o in (("row-major 32x1 warps", np.arange(W * H)),):
        u2, c2 = utilisation(prim.reshape(-1)[o])
        print(f"                      {name:36s} lanes busy {32 * u2:5.2f} / 32   warp-iterations {c2 / 1e6:8.2f} M")
    # shadow list: strip order, pending pixels only
    flat = shad.reshape(-1)
    idx = order[keep]
    lst = idx[flat[idx] > 0]
    s = flat[lst]
    base_u, base_c = utilisation(s)
    print(f"trace_shadow_kernel   compacted list, strip order          lanes busy {32 * base_u:5.2f} / 32   warp-iterations {base_c / 1e6:8.2f} M")

    def report(name, perm):
        u3, c3 = utilisation(s[perm])
        print(f"                      {name:36s} lanes busy {32 * u3:5.2f} / 32   warp-iterations {c3 / 1e6:8.2f} M  ({100.0 * (1 - c3 / base_c):+5.1f} % warp-iterations saved)")

    report("sorted by true iterations (bound)", np.argsort(s, kind="stable"))
    # predictors available in shade_kernel when the entry is appended: the primary ray's own iteration count / hit distance, the pixel row
    pp = prim.reshape(-1)[lst]
    for bins in (4, 16):
        q = np.quantile(pp, np.linspace(0, 1, bins + 1)[1:-1])
        report(f"{bins} bins by the primary ray's iterations", np.argsort(np.digitize(pp, q), kind="stable"))
    yy = lst // W
    for bins in (4, 16):
        report(f"{bins} bins by image row", np.argsort((yy * bins) // H, kind="stable"))
    # strip-local: sort inside every run of 128 / 512 entries (what one CTA could do in shared memory before appending)
    for run in (128, 512, 4096):
        perm = np.concatenate([k + np.argsort(s[k:k + run], kind="stable") for k in range(0, len(s), run)])
        report(f"true iterations sorted inside runs of {run}", perm)


if __name__ == "__main__":
    main()
