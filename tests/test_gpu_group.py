"""vx_group_*: multi-GPU rendering driven from ONE process through the C ABI (SURVEY §8b "Ownership": the context owns the
NCCL communicators; §5: the reference is a single process). Every result of a group must be byte-identical to the single-GPU
context's on the same inputs:

  * vx_group_svo_commit   packed H2D to device 0 -> ncclBroadcast -> scatter on every replica (initial bulk load + an edit)
  * vx_group_render       interleaved macro blocks, peer stores into device 0's RGBA32F frame, CUDA events between the devices
  * vx_group_render_read_rgba8   whole stripes per device (VX_SHARD_ROWS), every device DMAs its own stripes into one host frame
  * vx_group_raycast      contiguous task slices per device

The two-device cases need 2 GPUs (skipped otherwise); the one-device group and the VX_SHARD_ROWS shard layout run on any GPU box.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
W, H = 1000, 562   # ragged: 562 = 35 macro rows + 2 pixel rows


def _world(pkg, fmt=0):
    world = pkg.World(radius=4, center=(-1, 2, 5), seed=1, fmt=fmt, terrain="reference")
    world.generate(0, 8)
    world.serialize()
    return world


def _views(pkg, world):
    out = []
    for fwd in ((1.0, -0.3, 0.0), (0.6, -0.35, 0.5), (0.2, -0.25, -1.0)):
        p = pkg.render_params(cam_pos=(-24.0, 80.0, 174.0), cam_fwd=fwd, fov_y_deg=72.0, aspect=W / H)
        q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
        q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
        out.append(pkg.to_vx_render_params(q))
    return out


def _tasks(pkg, world, n, seed=3):
    rng = np.random.default_rng(seed)
    t = np.zeros(n, dtype=pkg.TASK_DTYPE)
    t["max_dst"] = -1.0
    o = world.cnv_block_pos((0.0, 0.0, 0.0))
    pos = rng.uniform(-100, 100, (n, 3)).astype(np.float32) + np.array([o[0], 0, o[2]], np.float32)
    pos[:, 1] = o[1] + rng.uniform(40.0, 220.0, n).astype(np.float32)
    t["pos"] = pos
    d = rng.normal(size=(n, 3)).astype(np.float32)
    t["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return t


def _edit(world):
    hgt = world.height_at(-10, 174)
    for dy in range(1, 9):
        world.edit_block(-10, hgt + dy, 174, 4)
    world.serialize()


def _run_group(pkg, devices, fmt):
    """Everything a group can do, next to a single-GPU context fed the same inputs; returns nothing, asserts equality."""
    reg = pkg.content_registry(pkg.load_atlas())
    world = _world(pkg, fmt)
    size_mb = world.size_bytes // 1_000_000 + 16
    one = pkg.Svo(reg, size_mb=size_mb, max_width=W, max_height=H, max_rays=1 << 16, flags=world.svo_flags)
    grp = pkg.SvoGroup(reg, devices, size_mb=size_mb, max_width=W, max_height=H, max_rays=1 << 16, flags=world.svo_flags)
    assert len(grp) == len(devices)
    world.mark_all_dirty()
    one.update(world)                       # drains the serializer's dirty list ...
    world.mark_all_dirty()                  # ... so mark everything again for the group
    grp.update(world)                       # bulk load: larger than the staging block or not, every replica gets the world
    views = _views(pkg, world)
    tasks = _tasks(pkg, world, 50_000)
    host = grp.host_frame(W, H)

    def compare(tag):
        for k, v in enumerate(views):       # a different view per frame: a device that is read too early shows
            one.render_raw(v, W, H)
            want32, want8 = one.read_rgba32f(), one.read_rgba8()
            grp.render_raw(v, W, H)
            got32 = grp.read_rgba32f()
            assert got32.tobytes() == want32.tobytes(), (tag, "vx_group_render", k, int((got32 != want32).any(axis=2).sum()))
            host[:] = 0x5a
            grp.render_read_rgba8(v, W, H, host.ctypes.data, bands=2)
            assert host.tobytes() == want8.tobytes(), (tag, "vx_group_render_read_rgba8", k, int((host != want8).any(axis=2).sum()))
        want = one.raycast_tasks(tasks)
        got = grp.raycast_tasks(tasks)
        assert got.tobytes() == want.tobytes(), (tag, "vx_group_raycast")
        assert (want["dst"] > 0).sum() > 1000

    compare("initial")
    # an edit: only the dirty ranges travel (staged path: one H2D + NCCL broadcast + scatter per replica)
    _edit(world)
    ranges = world.dirty_ranges()
    assert 0 < sum(l for _, l in ranges) < world.size_bytes // 4
    mirror_one = one.host_mirror(world.header_bytes + world.size_bytes)
    assert world.write_changes_to(mirror_one, reset=False)
    one.commit(float(np.float32(2.0 ** -world.depth)), ranges, world.size_bytes, world.depth)
    grp.update(world)
    compare("after edit")
    for i in range(len(grp)):
        n = C.c_uint32()
        assert pkg.lib().vx_svo_scatter_errors(grp.ctx(i), C.byref(n)) == 0 and n.value == 0
        assert pkg.lib().vx_frame_sync_errors(grp.ctx(i), C.byref(n)) == 0 and n.value == 0    # no wait of the overlapped wavefront gave up
    grp.close()
    one.close()


@pytest.mark.parametrize("fmt", [0, 1], ids=["esvo", "csvo"])
def test_group_of_one_device(pkg, fmt):
    _run_group(pkg, [0], fmt)


@pytest.mark.parametrize("fmt", [0, 1], ids=["esvo", "csvo"])
def test_group_of_two_devices_one_process(pkg, fmt):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_group(pkg, [0, 1], fmt)


def test_group_errors(pkg):
    L = pkg.lib()
    cfg = pkg.VxConfig(0, 0, 1_000_000, 64, 64, 16)
    g = C.c_void_p()
    devs = (C.c_int * 2)(0, 0)
    assert L.vx_group_create(C.byref(cfg), devs, 2, C.byref(g)) == -1 and b"twice" in L.vx_group_last_error(None)   # VX_E_ARG
    devs = (C.c_int * 1)(99)
    assert L.vx_group_create(C.byref(cfg), devs, 1, C.byref(g)) == -1                                              # device out of range
    assert L.vx_group_create(C.byref(cfg), None, 0, C.byref(g)) == -1


def test_row_shards_tile_the_frame(pkg):
    """VX_SHARD_ROWS on one GPU: the stripes of 3 shards, rendered one after the other into one host frame through the strided DMA of
    vx_render_read_rgba8, are the whole frame — ragged last stripe included."""
    reg = pkg.content_registry(pkg.load_atlas())
    world = _world(pkg)
    svo = pkg.Svo(reg, size_mb=world.size_bytes // 1_000_000 + 16, max_width=W, max_height=H, max_rays=16)
    world.mark_all_dirty()
    svo.update(world)
    v = _views(pkg, world)[0]
    svo.render_raw(v, W, H)
    want = svo.read_rgba8()
    import torch
    host = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    for bands in (1, 3):
        host.fill_(0x5a)
        for r in range(3):
            svo.render_read_rgba8(v, W, H, host.data_ptr(), bands=bands, shard=(r, 3 | pkg.VX_SHARD_ROWS))
            rows = [y for y in range(H) if (y // 16) % 3 == r]
            assert host.numpy()[rows].tobytes() == want[rows].tobytes(), (bands, r)
        assert host.numpy().tobytes() == want.tobytes()
    svo.close()


def test_overlapped_wavefront_with_bands_is_live(pkg):
    """Regression (round 2): with the overlapped wavefront the shade kernel's CTAs wait for the tracing kernel. At the start of a
    band both kernels become runnable on an idle GPU at the same instant; when the shade grid won every SM first, no tracing CTA
    fitted next to it (shared-memory carve-out), every wait timed out and the band came out wrong — about once in 300 frames.
    The shade kernel now pads its shared memory so that a tracing CTA always fits. 150 banded, sharded frames with the overlap
    forced on: every frame right, no wait gave up."""
    reg = pkg.content_registry(pkg.load_atlas())
    world = _world(pkg)
    svo = pkg.Svo(reg, size_mb=world.size_bytes // 1_000_000 + 16, max_width=W, max_height=H, max_rays=16)
    world.mark_all_dirty()
    svo.update(world)
    views = _views(pkg, world)
    svo.set_option(pkg.OPT_OVERLAP, 0)
    want = []
    for v in views:
        svo.render_raw(v, W, H)
        want.append(svo.read_rgba8())
    svo.set_option(pkg.OPT_OVERLAP, 1)
    import torch
    host = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    for it in range(150):
        k = it % 3
        host.fill_(0x5a)
        for r in range(2):
            svo.render_read_rgba8(views[k], W, H, host.data_ptr(), bands=2 + (it % 2), shard=(r, 2 | pkg.VX_SHARD_ROWS))
        assert host.numpy().tobytes() == want[k].tobytes(), it
    assert svo.frame_sync_errors() == 0
    svo.close()
