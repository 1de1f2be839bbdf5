#!/usr/bin/env python
"""Longer random sweeps of the emulated kernels than the seeded one in tests/test_kernels_emulated.py (run by hand before / after a
kernel change; a few minutes on the CPU):

    python tests/analysis/fuzz_emulated.py frames [seed] [cases]     frame kernels: sizes, shards, bands, refill, CTAs, formats, worlds
    python tests/analysis/fuzz_emulated.py picker [seed] [cases]     picker kernel: batch sizes around the run length, max_dst, refill

Every case must reproduce the oracle bit for bit (frames, RGBA8, counters, picker records). Needs tests/emu/libkernels_emu.so
(built by the test module's fixture: run pytest tests/test_kernels_emulated.py once)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import __graft_entry__ as g      # noqa: E402
import bench                     # noqa: E402

pkg, ora = g.load_pkg(), g.load_oracle()
import helpers                   # noqa: E402
import test_kernels_emulated as T   # noqa: E402

emu = C.CDLL(os.path.join(os.path.dirname(HERE), "emu", "libkernels_emu.so"))
reg = pkg.content_registry(pkg.load_atlas())
mode = sys.argv[1] if len(sys.argv) > 1 else "frames"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
cases = int(sys.argv[3]) if len(sys.argv) > 3 else 120
rng = np.random.default_rng(seed)
terrain, mc = {}, {}
for fmt in (0, 1):
    w = pkg.World(radius=5, center=(-1, 2, 5), seed=1, fmt=fmt)
    w.generate(0, 8)
    w.serialize()
    terrain[fmt] = w
    mc[fmt] = helpers.mc_world(pkg, fmt)
bad = 0

def frames_case(case):
    fmt, use_mc = int(rng.integers(0, 2)), bool(rng.integers(0, 3) == 0)
    world = (mc if use_mc else terrain)[fmt]
    w, h = int(rng.integers(1, 100)), int(rng.integers(1, 70))
    size, bands = int(rng.choice([1, 1, 2, 3, 5, 8])), int(rng.choice([1, 2, 3, 4, 16]))
    refill, ctas, srf = int(rng.integers(1, 33)), int(rng.integers(1, 6)), int(rng.choice([0, 1, 7, 32]))
    shadows, rgba8 = bool(rng.integers(0, 2)), int(rng.integers(0, 2))
    tma = int(rng.integers(0, 2)) if not rgba8 else 0
    if use_mc:
        p = helpers.mc_params(pkg, w, h, shadows=shadows)
        q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
        q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
        vxp = pkg.to_vx_render_params(q)
    else:
        vxp = T.world_params(pkg, world, w, h, shadows=shadows)
    want, want8, cnt = T.oracle_render(pkg, ora, world, reg, vxp, w, h)
    union, union8 = np.full((h, w, 4), -1.0, np.float32), np.zeros((h, w, 4), np.uint8)
    total = {k: 0 for k in cnt}
    for rank in range(size):
        got, got8, c = T.emu_render(emu, pkg, world, reg, vxp, w, h, refill=refill, shadow_refill=srf, ctas=ctas, rgba8=rgba8, tma=tma, rank=rank,
                                    size=size, bands=bands)
        mine = (got8.view(np.uint32)[..., 0] != 0xdeadbeef) if rgba8 else (got[..., 3] != -1.0)
        union[mine], union8[mine] = got[mine], got8[mine]
        for k in c:
            total[k] += c[k]
    ok = ((union8.tobytes() == want8.tobytes()) if rgba8 else (union.tobytes() == want.tobytes())) and total == cnt
    if not ok:
        print("MISMATCH", dict(case=case, fmt=fmt, mc=use_mc, w=w, h=h, size=size, bands=bands, refill=refill, ctas=ctas, srf=srf, shadows=shadows,
                               rgba8=rgba8, tma=tma), total, cnt, flush=True)
    return ok


def picker_case(case):
    kind, fmt = ("t", "mc")[int(rng.integers(0, 2))], int(rng.integers(0, 2))
    world = (terrain if kind == "t" else mc)[fmt]
    n = int(rng.choice([1, 31, 32, 33, 127, 128, 129, 500, 1000, 2500]))
    refill, ctas, md = int(rng.integers(1, 33)), int(rng.integers(1, 6)), float(rng.choice([-1.0, 0.0, 5.0, 30.0, 200.0]))
    if kind == "t":
        tasks = bench.picker_tasks(pkg, world, 5, n, seed=case, max_dst=md)
    else:
        tasks = np.zeros(n, dtype=pkg.TASK_DTYPE)
        tasks["max_dst"] = md
        tasks["pos"] = np.array(world.cnv_block_pos((-2090.0, 75.0, 1690.0))) + rng.uniform(-40, 40, (n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        tasks["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    want, cnt = helpers.oracle_scene(ora, world, reg).raycast(tasks)
    got, c = T.emu_raycast(emu, pkg, world, reg, tasks, refill=refill, ctas=ctas)
    ok = got.tobytes() == want.tobytes() and (c["steps"], c["pushes"], c["leaf_tests"]) == (cnt["steps"], cnt["pushes"], cnt["leaf_tests"])
    if not ok:
        print("MISMATCH", kind, fmt, n, refill, ctas, md, flush=True)
    return ok


for case in range(cases):
    bad += not (frames_case if mode == "frames" else picker_case)(case)
print(f"{mode}: {cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
