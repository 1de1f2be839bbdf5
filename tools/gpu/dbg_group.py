import sys, ctypes as C
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import torch
print("devices", torch.cuda.device_count(), flush=True)
import __graft_entry__ as g
pkg = g.load_pkg()
import test_gpu_group as T
W, H = T.W, T.H
reg = pkg.content_registry(pkg.load_atlas())
world = T._world(pkg, 0)
size_mb = world.size_bytes // 1_000_000 + 16
one = pkg.Svo(reg, size_mb=size_mb, max_width=W, max_height=H, max_rays=1 << 16, flags=world.svo_flags)
grp = pkg.SvoGroup(reg, [0, 1], size_mb=size_mb, max_width=W, max_height=H, max_rays=1 << 16, flags=world.svo_flags)
world.mark_all_dirty(); one.update(world); world.mark_all_dirty(); grp.update(world)
views = T._views(pkg, world)
host = grp.host_frame(W, H)
def report(tag, got, want):
    bad = (got != want).any(axis=2)
    rows = np.nonzero(bad.any(axis=1))[0]
    mrows = sorted(set((rows // 16).tolist()))
    fill = int((got[bad] == 0x5a).all(axis=1).sum()) if bad.any() else 0
    zero = int((got[bad] == 0).all(axis=1).sum()) if bad.any() else 0
    print(f"{tag}: bad px {int(bad.sum())} macro rows {mrows[:24]} untouched(0x5a) {fill} zero {zero}", flush=True)

for rep in range(2):
  for overlap in (2, 0):
    grp.set_option(pkg.OPT_OVERLAP, overlap)
    for k, v in enumerate(views):
        one.render_raw(v, W, H); want32, want8 = one.read_rgba32f(), one.read_rgba8()
        grp.render_raw(v, W, H)
        got32 = grp.read_rgba32f()
        print(f"rep {rep} overlap={overlap} view {k}: vx_group_render equal {got32.tobytes() == want32.tobytes()}", flush=True)
        for bands in (2, 1):
            host[:] = 0x5a
            grp.render_read_rgba8(v, W, H, host.ctypes.data, bands=bands)
            report(f"  read_rgba8 bands={bands}", host, want8)
            host[:] = 0x5a
            grp.render_read_rgba8(v, W, H, host.ctypes.data, bands=bands)
            report(f"  read_rgba8 bands={bands} again", host, want8)
