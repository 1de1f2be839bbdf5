import sys, ctypes as C
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import torch
import __graft_entry__ as g
pkg = g.load_pkg()
import test_gpu_group as T
W, H = T.W, T.H
reg = pkg.content_registry(pkg.load_atlas())
world = T._world(pkg, 0)
size_mb = world.size_bytes // 1_000_000 + 16
ndev = torch.cuda.device_count()
one = pkg.Svo(reg, size_mb=size_mb, max_width=W, max_height=H, max_rays=1 << 16, flags=world.svo_flags)
world.mark_all_dirty(); one.update(world)
views = T._views(pkg, world)
wants = []
for v in views:
    one.set_option(pkg.OPT_OVERLAP, 0)
    one.render_raw(v, W, H); wants.append((one.read_rgba32f(), one.read_rgba8()))
def rows_of(bad):
    return sorted(set((np.nonzero(bad.any(axis=1))[0] // 16).tolist()))[:16]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
# 1. single context, whole frame, overlap on/off
for overlap in (1, 0):
    one.set_option(pkg.OPT_OVERLAP, overlap)
    fails = 0
    for it in range(N):
        k = it % 3
        one.render_raw(views[k], W, H)
        got = one.read_rgba32f()
        if got.tobytes() != wants[k][0].tobytes():
            fails += 1
            if fails <= 3: print(f"  single ctx overlap={overlap} it={it}: bad rows {rows_of((got != wants[k][0]).any(axis=2))} sync_errors {one.frame_sync_errors()}", flush=True)
    print(f"single ctx vx_render overlap={overlap}: {fails}/{N} frames wrong", flush=True)
# 2. single context, row shards through vx_render_read_rgba8
host = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
for overlap in (1, 0):
    one.set_option(pkg.OPT_OVERLAP, overlap)
    fails = 0
    for it in range(N):
        k = it % 3
        host.fill_(0x5a)
        for r in range(2):
            one.render_read_rgba8(views[k], W, H, host.data_ptr(), bands=2, shard=(r, 2 | pkg.VX_SHARD_ROWS))
        if host.numpy().tobytes() != wants[k][1].tobytes():
            fails += 1
            if fails <= 3: print(f"  single ctx rows overlap={overlap} it={it}: bad rows {rows_of((host.numpy() != wants[k][1]).any(axis=2))}", flush=True)
    print(f"single ctx row shards overlap={overlap}: {fails}/{N} frames wrong", flush=True)
# 3. the group
if ndev >= 2:
    grp = pkg.SvoGroup(reg, [0, 1], size_mb=size_mb, max_width=W, max_height=H, max_rays=1 << 16, flags=world.svo_flags)
    world.mark_all_dirty(); grp.update(world)
    gh = grp.host_frame(W, H)
    for overlap in (1, 0):
        grp.set_option(pkg.OPT_OVERLAP, overlap)
        f32 = f8 = 0
        for it in range(N):
            k = it % 3
            grp.render_raw(views[k], W, H)
            got = grp.read_rgba32f()
            if got.tobytes() != wants[k][0].tobytes():
                f32 += 1
                if f32 <= 3: print(f"  group render overlap={overlap} it={it}: bad rows {rows_of((got != wants[k][0]).any(axis=2))}", flush=True)
            gh[:] = 0x5a
            grp.render_read_rgba8(views[k], W, H, gh.ctypes.data, bands=2)
            if gh.tobytes() != wants[k][1].tobytes():
                f8 += 1
                if f8 <= 3: print(f"  group read overlap={overlap} it={it}: bad rows {rows_of((gh != wants[k][1]).any(axis=2))} untouched {int((gh == 0x5a).all(axis=2).sum())}", flush=True)
        print(f"group overlap={overlap}: vx_group_render {f32}/{N} wrong, vx_group_render_read_rgba8 {f8}/{N} wrong", flush=True)
        for i in range(2):
            n = C.c_uint32(); pkg.lib().vx_frame_sync_errors(grp.ctx(i), C.byref(n)); print("  sync errors dev", i, n.value)
