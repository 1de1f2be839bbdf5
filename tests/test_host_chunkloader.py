"""systems::chunkloader::ChunkLoader on the host mirror (voxel-rs_b200/host/chunkloader.hpp): the reference's two unit tests
(src/systems/chunkloader.rs:155-259), and the streaming step built from it (voxelrs_b200.follow): a world that followed the camera
into the next chunk is the world one would have generated there. CPU only."""
import ctypes as C

import numpy as np

ORDER = {"load": 0, "unload": 1, "lod": 2}   # derive(Ord) on ChunkEvent: variant order, then the fields


def sorted_events(ev):
    return sorted(ev, key=lambda e: (ORDER[e[0]], e[1], e[2]))


def test_load_and_unload(pkg):
    """chunkloader.rs:155-216"""
    cl = pkg.ChunkLoader(1, 0, 1)
    L = lambda x, y, z: ("load", (x, y, z), 5)
    U = lambda x, y, z: ("unload", (x, y, z), 0)
    assert sorted_events(cl.update((0.0, 0.0, 0.0))) == [L(-1, 0, 0), L(0, 0, -1), L(0, 0, 0), L(0, 0, 1), L(1, 0, 0)]
    assert cl.update((16.0, 16.0, 16.0)) == []                      # same chunk
    assert sorted_events(cl.update((32.0, 0.0, 0.0))) == [L(1, 0, -1), L(1, 0, 1), L(2, 0, 0), U(-1, 0, 0), U(0, 0, -1), U(0, 0, 1)]
    assert sorted_events(cl.update((128.0, 0.0, 0.0))) == [L(3, 0, 0), L(4, 0, -1), L(4, 0, 0), L(4, 0, 1), L(5, 0, 0),
                                                           U(0, 0, 0), U(1, 0, -1), U(1, 0, 0), U(1, 0, 1), U(2, 0, 0)]
    assert sorted_events(cl.update((128.0, 64.0, 0.0))) == [U(3, 0, 0), U(4, 0, -1), U(4, 0, 0), U(4, 0, 1), U(5, 0, 0)]   # y out of reach
    assert cl.update((0.0, 64.0, 0.0)) == []
    assert cl.loaded_count == 0


def lod_scale_on_x_axis(events, z):
    cols = {c[0]: lod for kind, c, lod in events if kind in ("load", "lod") and c[2] == z}
    return [cols[x] for x in sorted(cols)]


def test_changing_lod(pkg):
    """chunkloader.rs:220-240"""
    cl = pkg.ChunkLoader(25, 0, 1)
    ev = cl.update((0.0, 0.0, 0.0))
    z0 = [2] * 6 + [3] * 7 + [4] * 6 + [5] * 13 + [4] * 6 + [3] * 7 + [2] * 6
    z1 = [2] * 5 + [3] * 7 + [4] * 6 + [5] * 13 + [4] * 6 + [3] * 7 + [2] * 5
    assert lod_scale_on_x_axis(ev, -1) == z1 and lod_scale_on_x_axis(ev, 0) == z0 and lod_scale_on_x_axis(ev, 1) == z1
    ev = cl.update((32.0, 0.0, 0.0))
    change = [2, 3, 4, 5, 4, 3, 2]
    assert lod_scale_on_x_axis(ev, -1) == change and lod_scale_on_x_axis(ev, 0) == change and lod_scale_on_x_axis(ev, 1) == change
    # events come nearest first (chunkloader.rs:116-120)
    d = [(c[0] - 1) ** 2 + c[1] ** 2 + c[2] ** 2 for _, c, _ in ev]
    assert d == sorted(d)


def test_loader_rejects_empty_height_range(pkg):
    """assert!(start_y < end_y), chunkloader.rs:35"""
    import pytest
    with pytest.raises(pkg.VxError):
        pkg.ChunkLoader(1, 3, 3)


def _frame(pkg, ora, reg, world, cam):
    w, h = 160, 90
    p = pkg.render_params(cam_pos=cam, cam_fwd=(0.4, -0.6, -1.0), fov_y_deg=72.0, aspect=w / h, render_shadows=True)
    q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
    q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
    tex, mips = reg.textures()
    img, cnt = ora.Scene(world.gpu_buffer(), reg.materials().tobytes(), tex, mips, fmt=world.fmt).render(pkg.to_vx_render_params(q), w, h)
    return img, cnt


def test_following_the_camera_equals_generating_there(pkg, ora):
    """gamelogic::World::update without the job system: loader events + re-centred window. After every step the world holds exactly the
    chunks, at exactly the LODs, of a world generated from scratch around the new camera chunk, and the oracle renders both to the
    same bits (LOD changes included: radius 8 has a LOD-5 and a LOD-4 ring)."""
    reg = pkg.content_registry(pkg.load_atlas())
    r = 8
    cam = [-24.0, 80.0, 174.0]
    world = pkg.World(radius=r, center=(-1, 2, 5), seed=1, terrain="reference")
    loader = pkg.ChunkLoader(r, 0, 8)
    ev = pkg.follow(world, loader, cam)
    assert ev and all(k == "load" for k, _, _ in ev)
    world.serialize()
    for step, move in enumerate(((32.0, 0.0, 0.0), (0.0, 0.0, -32.0), (32.0, 0.0, 32.0))):
        if step:
            cam = [cam[i] + move[i] for i in range(3)]
            ev = pkg.follow(world, loader, cam)
            kinds = {k for k, _, _ in ev}
            assert kinds == {"load", "unload", "lod"}, kinds
            world.serialize()
        center = tuple(int(v) >> 5 for v in cam)
        fresh = pkg.World(radius=r, center=center, seed=1, terrain="reference")
        fresh.generate(0, 7)
        fresh.serialize()
        got = {tuple(c) for c in world.chunks().tolist()}
        want = {tuple(c) for c in fresh.chunks().tolist()}
        assert got == want, (step, len(got), len(want), sorted(got ^ want)[:5])
        # LOD per chunk: the serialized record set of every chunk has the size it has in the fresh world
        assert all(world.chunk_range(c)[1] == fresh.chunk_range(c)[1] for c in sorted(want)), step
        eye = (cam[0], 130.0, cam[2])
        a, ca = _frame(pkg, ora, reg, world, eye)
        b, cb = _frame(pkg, ora, reg, fresh, eye)
        assert ca["leaf_tests"] > 1000 and a.tobytes() == b.tobytes(), step
