"""World-box clipping (traverse.cuh Clip / vx_set_option 12; restated in oracle/oracle.cpp behind vxo_set_clip): a ray stops once it
has left the box that holds every voxel. It is NOT in the reference shader, so it has to be invisible in every output:

  * CPU: the oracle with the extension on and off produces bit-identical frames and picker results on terrain worlds (both SVO
    formats), a one-voxel and a two-voxel world — only the iteration counters shrink;
  * GPU: the same through the C ABI (vx_set_option 12 on / off), and the kernels' counters equal the oracle's in both modes.
"""
import ctypes as C

import numpy as np
import pytest

import helpers


def _terrain(pkg, fmt):
    w = pkg.World(radius=3, center=(-1, 2, 5), seed=1, fmt=fmt, terrain="reference")
    w.generate(0, 8)
    w.serialize()
    return w


def _params(pkg, world, w, h, fwd=(1.0, -0.3, 0.0)):
    p = pkg.render_params(cam_pos=(-24.0, 80.0, 174.0), cam_fwd=fwd, fov_y_deg=72.0, aspect=w / h)
    q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
    q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
    return pkg.to_vx_render_params(q)


def _tasks(pkg, world, n, seed=1):
    rng = np.random.default_rng(seed)
    t = np.zeros(n, dtype=pkg.TASK_DTYPE)
    t["max_dst"] = np.where(rng.random(n) < 0.5, -1.0, rng.uniform(0, 200, n)).astype(np.float32)
    o = world.cnv_block_pos((0.0, 0.0, 0.0))
    t["pos"] = (rng.uniform(-150, 150, (n, 3)) + np.array([o[0], o[1] + 100, o[2]])).astype(np.float32)   # inside, above, below and beside the box
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[: n // 8] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 8)] * rng.choice([-1.0, 1.0], (n // 8, 1)).astype(np.float32)   # axis-parallel
    t["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return t


@pytest.mark.parametrize("fmt", [0, 1], ids=["esvo", "csvo"])
def test_oracle_clip_changes_counters_only(pkg, ora, fmt):
    world = _terrain(pkg, fmt)
    reg = pkg.content_registry(pkg.load_atlas())
    scene = helpers.oracle_scene(ora, world, reg)
    w, h = 160, 90
    tasks = _tasks(pkg, world, 20000)
    out = {}
    for clip in (False, True):
        ora.set_clip(clip)
        try:
            frames = [scene.render(_params(pkg, world, w, h, fwd), w, h) for fwd in ((1.0, -0.3, 0.0), (0.0, 0.6, -1.0), (0.2, -1.0, 0.1))]
            out[clip] = (frames, scene.raycast(tasks))
        finally:
            ora.set_clip(False)
    for (f0, c0), (f1, c1) in zip(out[False][0], out[True][0]):
        assert f0.tobytes() == f1.tobytes()
        assert c1["primary_rays"] == c0["primary_rays"] and c1["shadow_rays"] == c0["shadow_rays"] and c1["leaf_tests"] == c0["leaf_tests"]
        assert c1["steps"] < c0["steps"]
    (r0, c0), (r1, c1) = out[False][1], out[True][1]
    assert r0.tobytes() == r1.tobytes() and c1["steps"] < c0["steps"] and (r0["dst"] > 0).sum() > 300


def test_oracle_clip_tiny_worlds(pkg, ora):
    reg = helpers.shader_test_registry(pkg)
    rng = np.random.default_rng(2)
    for blocks in ([(3, 4, 5, 1)], [(0, 0, 0, 1), (31, 31, 31, 2)], [(x, 0, z, 1) for x in range(32) for z in range(32)]):
        world = helpers.shader_test_world(pkg, blocks)
        scene = helpers.oracle_scene(ora, world, reg)
        t = np.zeros(4000, dtype=pkg.TASK_DTYPE)
        t["max_dst"] = -1.0
        t["pos"] = rng.uniform(-20, 52, (4000, 3)).astype(np.float32)
        d = rng.normal(size=(4000, 3)).astype(np.float32)
        t["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        t["pos"][:50] = np.array([3.5, 4.5, 5.5], np.float32)     # origins inside a voxel (inside_voxel flag, PUSH-below-leaf quirk)
        ora.set_clip(False)
        r0, c0 = scene.raycast(t)
        ora.set_clip(True)
        try:
            r1, c1 = scene.raycast(t)
        finally:
            ora.set_clip(False)
        assert r0.tobytes() == r1.tobytes() and c1["steps"] <= c0["steps"]
        assert (r0["dst"] > 0).any() or len(blocks) < 3


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [0, 1], ids=["esvo", "csvo"])
def test_gpu_clip_changes_counters_only(pkg, ora, fmt):
    world = _terrain(pkg, fmt)
    reg = pkg.content_registry(pkg.load_atlas())
    scene = helpers.oracle_scene(ora, world, reg)
    w, h = 320, 180
    svo = pkg.Svo(reg, size_mb=world.size_bytes // 1_000_000 + 8, max_width=w, max_height=h, max_rays=1 << 16, flags=world.svo_flags)
    world.mark_all_dirty()
    svo.update(world)
    svo.set_option(pkg.OPT_COUNT, 1)
    tasks = _tasks(pkg, world, 30000)
    got = {}
    for clip in (0, 1):
        svo.set_option(pkg.OPT_CLIP, clip)
        ora.set_clip(bool(clip))
        try:
            vxp = _params(pkg, world, w, h)
            svo.render_raw(vxp, w, h)
            frame, st = svo.read_rgba32f(), svo.frame_stats(0)
            want, cnt = scene.render(vxp, w, h)
            for k in ("primary_rays", "shadow_rays", "steps", "pushes", "leaf_tests"):
                assert st[k] == cnt[k], (clip, k, st[k], cnt[k])
            res = svo.raycast_tasks(tasks)
            rst = svo.frame_stats(1)
            rwant, rcnt = scene.raycast(tasks)
            assert res.tobytes() == rwant.tobytes() and rst["steps"] == rcnt["steps"]
            got[clip] = (frame, res, st["steps"], rst["steps"])
        finally:
            ora.set_clip(False)
    assert got[0][0].tobytes() == got[1][0].tobytes() and got[0][1].tobytes() == got[1][1].tobytes()
    assert got[1][2] < got[0][2] and got[1][3] < got[0][3]
    svo.close()
