"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's golden vectors.
Needs a GPU: run with `pytest -m gpu` on the B200 box.

Bars (north_star): hit voxel / material / face bit-exact, hit distance <= 1e-5 relative (here: bit-exact),
shaded RGB8 within +-1 LSB (transcendentals powf/acosf differ by ulps between libm and CUDA).
"""
import ctypes as C
import os

import numpy as np
import pytest
from PIL import Image

import helpers
from golden import reference_vectors as gv
from test_oracle_golden import check_frames, check_result, close

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _oracle_follows_the_products_clip(ora):
    """The kernels stop a ray once it has left the occupied box of the world (vx_set_option 12, on by default): results are the
    shader's bit for bit, the iteration counters are smaller. The oracle restates that extension when asked, so that counters can
    still be compared one to one; it is switched off again for every other test (the oracle's own goldens run the shader as written)."""
    ora.set_clip(True)
    yield
    ora.set_clip(False)



def make_svo(pkg, reg, world, size_mb=8, w=640, h=490, rays=1 << 20, flags=0):
    svo = pkg.Svo(reg, size_mb=size_mb, max_width=w, max_height=h, max_rays=rays, flags=flags | world.svo_flags)
    world.mark_all_dirty()   # a fresh GPU buffer needs the whole RangeBuffer, not just the changes since the last update
    svo.update(world)
    return svo


@pytest.fixture(scope="module")
def reg(pkg):
    return helpers.shader_test_registry(pkg)


def cast_both(pkg, ora, reg, blocks, pos, d, max_dst, translucent, svo_pos=(0, 0, 0), fmt=0):
    w = helpers.shader_test_world(pkg, blocks, svo_pos, fmt=fmt)
    s = helpers.oracle_scene(ora, w, reg)
    svo = make_svo(pkg, reg, w, size_mb=2, w=8, h=8, rays=16)
    o_res, o_frames, o_n = s.debug_cast(pos, d, max_dst, translucent)
    g_res, g_frames, g_n = svo.debug_cast(pos, d, max_dst, translucent)
    svo.close()
    return (o_res, o_frames, o_n), (g_res, g_frames, g_n)


def assert_same_cast(o, g):
    (o_res, o_frames, o_n), (g_res, g_frames, g_n) = o, g
    assert o_n == g_n
    for a, b in zip(o_frames, g_frames):
        assert bytes(a) == bytes(b), (a.as_tuple(), b.as_tuple())
    assert o_res.as_dict() == g_res.as_dict(), (o_res.as_dict(), g_res.as_dict())


def test_golden_traces_on_gpu(pkg, ora, reg):
    """vx_debug_cast reproduces svo_shader_tests.rs:293-334 and :707-753 frame by frame, bit-identical to the oracle."""
    for g in (gv.TRAVERSAL, gv.HIGHER_COORDS):
        o, c = cast_both(pkg, ora, reg, g["blocks"], g["pos"], g["dir"], g["max_dst"], g["cast_translucent"], g.get("svo_pos", (0, 0, 0)))
        assert_same_cast(o, c)
        check_frames(c[1], c[2], g["frames"])
        check_result(c[0], g["result"])


def test_golden_results_on_gpu(pkg, ora, reg):
    """cast_inside_outside_all_axes, uv_coords_on_all_sides, translucent leafs, inside leaf (svo_shader_tests.rs:340-701)."""
    g = gv.ALL_AXES
    for name, pos, d, t, face, hit_pos, uv in g["cases"]:
        o, c = cast_both(pkg, ora, reg, g["blocks"], pos, d, 100.0, False)
        assert_same_cast(o, c)
        check_result(c[0], {"t": t, "value": 1, "face_id": face, "pos": hit_pos, "uv": uv, "color": g["color"], "inside_voxel": False}, name=name)
    g = gv.UV_COORDS
    for pos, d, uv, color in g["cases"]:
        o, c = cast_both(pkg, ora, reg, g["blocks"], pos, d, 32.0, False)
        assert_same_cast(o, c)
        assert close(c[0].as_dict()["uv"], uv) and close(c[0].as_dict()["color"], color)
    g = gv.TRANSLUCENT
    for name, pos, translucent, exp in g["cases"]:
        o, c = cast_both(pkg, ora, reg, g["blocks"], pos, g["dir"], 32.0, translucent)
        assert_same_cast(o, c)
        assert c[0].as_dict()["value"] == exp["value"] and c[0].as_dict()["face_id"] == exp["face_id"], name
    g = gv.INSIDE_LEAF
    for name, pos, d, exp in g["cases"]:
        o, c = cast_both(pkg, ora, reg, g["blocks"], pos, d, 32.0, False)
        assert_same_cast(o, c)
        check_result(c[0], exp, name=name)


def test_picker_golden_host_mirror(pkg):
    """svo_tests::raycast (src/graphics/svo.rs:402-449) through graphics::Svo::raycast's mirror."""
    g = gv.PICKER_RAYCAST
    reg = helpers.svo_render_test_registry(pkg, pkg.load_atlas())
    w = pkg.World()
    w.set_leaf_blocks((0, 0, 0), g["blocks"], compact=False)
    w.serialize()
    svo = make_svo(pkg, reg, w, size_mb=10, w=8, h=8, rays=100)
    rays = [tuple(p) + tuple(d) + (m,) for p, d, m in g["rays"]]
    ro, _ = svo.raycast(rays=rays)
    for r, (dst, inside, pos, normal) in zip(ro, g["expected"]):
        assert close(r[0], dst, 1e-4) and bool(r[1]) == inside and close(r[2:5], pos, 1e-4) and tuple(r[5:8]) == normal
    # AABB probe (player box of game.rs:73) standing on the two blocks: -y distance is the gap to the floor
    _, ao = svo.raycast(aabbs=[((1.0, 1.25, 0.5), (-0.4, 0.0, -0.4), (0.8, 1.8, 0.8))])
    assert close(ao[0][1], 0.25, 1e-4)      # neg.y
    assert ao[0][3] == -1.0 and ao[0][4] == -1.0   # nothing in +x / +y within 10
    svo.close()


def small_scenes(pkg):
    rng = np.random.default_rng(3)
    dense = (rng.random((32, 32, 32)) < 0.08).astype(np.uint32) * rng.integers(1, 5, (32, 32, 32)).astype(np.uint32)
    blocks = [(int(x), int(y), int(z), int(dense[z, y, x])) for z, y, x in zip(*np.nonzero(dense))]
    yield "random8pct", blocks, (0, 0, 0)
    yield "higher", gv.HIGHER_COORDS["blocks"], (15, 15, 15)
    yield "translucent", gv.TRANSLUCENT["blocks"] + [(x, 3, z, 3 + (x & 1)) for x in range(12) for z in range(12)], (0, 0, 0)


def test_picker_random_rays_bit_exact(pkg, ora, reg):
    """200k incoherent rays per scene, inside and outside the tree, limited and unlimited: results byte-identical."""
    for name, blocks, svo_pos in small_scenes(pkg):
        w = helpers.shader_test_world(pkg, blocks, svo_pos)
        s = helpers.oracle_scene(ora, w, reg)
        svo = make_svo(pkg, reg, w, size_mb=4, w=8, h=8, rays=1 << 18)
        size = 32 * (max(svo_pos) + 1)
        for max_dst, seed in ((-1.0, 1), (30.0, 2), (0.75, 3)):
            tasks = helpers.random_tasks(pkg, 200_000, -8.0, size + 8.0, max_dst, seed)
            tasks["dir"][::17, 1] = 0.0          # exercise the epsilon clamp (svo.esvo.glsl:85-89)
            tasks["pos"][::5] = np.floor(tasks["pos"][::5]) + 0.5
            want, ocnt = s.raycast(tasks)
            # bin = ray binning (vx_set_option 15: Z-order of the origin cells, bits per axis, +16 direction octant; 0 = task order). 200 k rays
            # are above its threshold. Results never move: byte-identical in task order.
            for refill, bin in ((24, 7 | 16), (1, 0), (32, 8), (20, 3 | 16)):
                svo.set_option(pkg.OPT_REFILL_PICKER, refill)
                if bin is not None:
                    svo.set_option(15, bin)
                svo.set_option(pkg.OPT_COUNT, 1)
                got = svo.raycast_tasks(tasks)
                assert got.tobytes() == want.tobytes(), (name, max_dst, refill, int((got["dst"] != want["dst"]).sum()))
                st = svo.frame_stats(1)
                assert st["steps"] == ocnt["steps"] and st["pushes"] == ocnt["pushes"] and st["leaf_tests"] == ocnt["leaf_tests"], (st, ocnt)
        svo.close()


def render_both(pkg, ora, reg, world, params, w, h, svo=None, use_world=False):
    own = svo is None
    if own:
        svo = make_svo(pkg, reg, world, size_mb=max(8, world.size_bytes // 1_000_000 + 8), w=w, h=h, rays=16)
    svo.render(params, w, h, world=world if use_world else None)
    got = svo.read_rgba32f()
    got8 = svo.read_rgba8()
    q = pkg.VxhRenderParams.from_buffer_copy(bytes(params))
    if use_world:
        q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(params.cam_pos)))
        if params.has_selected_voxel:
            q.selected_voxel = (C.c_float * 3)(*world.cnv_block_pos(tuple(params.selected_voxel)))
    s = helpers.oracle_scene(ora, world, reg)
    want, cnt = s.render(pkg.to_vx_render_params(q), w, h)
    if own:
        svo.close()
    return got, got8, want, ora.to_rgba8(want), cnt


def assert_frames_match(got, got8, want, want8):
    d8 = np.abs(got8.astype(np.int32) - want8.astype(np.int32))
    assert d8.max() <= 1, f"RGB8 differs by {d8.max()} LSB in {(d8 > 1).sum()} channels"
    # sky / unlit paths share no transcendental with libm differences beyond ulps: float frames agree to 1e-5
    assert np.allclose(got, want, rtol=0, atol=2e-6), float(np.abs(got - want).max())
    return float((got == want).mean())


def test_render_reference_scene(pkg, ora):
    """svo_tests::render scene (src/graphics/svo.rs:342-399): shading, normal maps, shadows, highlight — vs oracle
    (+-1 LSB) and vs the reference's expected PNG (reference metric, threshold 0.001)."""
    reg = helpers.svo_render_test_registry(pkg, pkg.load_atlas())
    w = pkg.World()
    w.set_leaf_blocks((0, 0, 0), helpers.svo_render_test_blocks(), compact=False)
    w.serialize()
    p = helpers.svo_render_test_params(pkg)
    got, got8, want, want8, cnt = render_both(pkg, ora, reg, w, p, 640, 490)
    exact = assert_frames_match(got, got8, want, want8)
    exp = np.asarray(Image.open(os.path.join(os.path.dirname(__file__), "golden", "graphics_svo_render_expected.png")).convert("RGBA"))
    diff = helpers.diff_images(got8[::-1], exp)
    print(f"reference scene: float-exact fraction {exact:.4f}, diff vs expected PNG {diff:.6f}")
    assert diff < 0.001


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [0, 1])
def test_world_end_to_end_png(pkg, ora, fmt):
    """tests::end_to_end, src/gamelogic/world.rs:461-498: the reference's generated world (Perlin terrain, radius 15, LOD rule),
    1024x768 with shadows, through the C ABI — vs the oracle (+-1 LSB, counters equal) and vs the reference's committed image with the
    reference's metric and default threshold 0.001 (both world formats: the image does not depend on SVO_TYPE)."""
    w = helpers.e2e_world(pkg, fmt=fmt)
    reg = pkg.content_registry(pkg.load_atlas())
    p = helpers.e2e_params(pkg)
    width, height = helpers.E2E_SIZE
    svo = make_svo(pkg, reg, w, size_mb=w.size_bytes // 1_000_000 + 8, w=width, h=height, rays=16)
    svo.set_option(pkg.OPT_COUNT, 1)
    got, got8, want, want8, cnt = render_both(pkg, ora, reg, w, p, width, height, svo=svo, use_world=True)
    assert_frames_match(got, got8, want, want8)
    st = svo.frame_stats(0)
    for k in ("primary_rays", "shadow_rays", "steps", "pushes", "leaf_tests", "tex_fetches"):
        assert st[k] == cnt[k], (k, st, cnt)
    svo.close()
    exp = helpers.e2e_expected()
    diff = helpers.diff_images(got8[::-1], exp)
    m = np.abs(got8[::-1][..., :3].astype(int) - exp[..., :3].astype(int)).max(-1)
    print(f"end-to-end world (fmt {fmt}): diff vs the reference's PNG {diff:.6f}, identical {(m == 0).mean():.4f}, within 1 LSB {(m <= 1).mean():.4f}")
    assert diff < 0.001, diff
    assert (m <= 1).mean() > 0.98


@pytest.fixture(scope="module")
def terrain(pkg):
    """Generated-terrain world of the named shape at a small radius (configs 2/3 shrunk for the oracle)."""
    world = pkg.World(radius=5, center=(-1, 2, 5), seed=1)
    world.generate(0, 8)
    world.serialize()
    reg = pkg.content_registry(pkg.load_atlas())
    return world, reg


def terrain_params(pkg, w, h, shadows=True, selected=None):
    return pkg.render_params(cam_pos=(-24.0, 80.0, 174.0), cam_fwd=(1.0, -0.3, 0.0), fov_y_deg=72.0, aspect=w / h, render_shadows=shadows,
                             selected_voxel=selected)


def test_render_terrain_variants(pkg, ora, terrain):
    """Terrain frame with LOD chunks, trilinear texture path (dst > 15), shadows: every kernel variant matches the
    oracle and each other; ray/step counters equal the oracle's."""
    world, reg = terrain
    w, h = 480, 270
    p = terrain_params(pkg, w, h)
    svo = make_svo(pkg, reg, world, size_mb=world.size_bytes // 1_000_000 + 8, w=w, h=h, rays=16)
    frames = {}
    # (refill threshold, CTAs/SM = register-budget build, TMA bulk-copy write-back of whole framebuffer strips; 270 rows leave a ragged strip)
    # and the LIFO hand-over of the wavefront buffers, vx_set_option 14: reverse consumer order + discarded lines vs streaming stores)
    for refill, ctas, tma, lifo in ((8, 0, 0, 7), (1, 5, 1, 3), (32, 6, 0, 0), (20, 8, 1, 1), (1, 0, 0, 0), (1, 0, 0, 7)):
        svo.set_option(14, lifo)
        svo.set_option(pkg.OPT_TMA, tma)
        svo.set_option(pkg.OPT_REFILL, refill)
        svo.set_option(pkg.OPT_CTAS_PER_SM, ctas)
        svo.set_option(pkg.OPT_COUNT, 1)
        got, got8, want, want8, cnt = render_both(pkg, ora, reg, world, p, w, h, svo=svo, use_world=True)
        assert_frames_match(got, got8, want, want8)
        st = svo.frame_stats(0)
        for k in ("primary_rays", "shadow_rays", "steps", "pushes", "leaf_tests", "tex_fetches"):
            assert st[k] == cnt[k], (refill, ctas, k, st, cnt)
        frames[(refill, ctas, tma, lifo)] = got
    base = frames[(8, 0, 0, 7)]
    for k, f in frames.items():
        assert f.tobytes() == base.tobytes(), k
    assert cnt["shadow_rays"] > 0 and cnt["tex_fetches"] > cnt["leaf_tests"]   # trilinear path exercised
    svo.close()


def test_primary_hits_bit_exact(pkg, ora, terrain):
    """north_star's geometry bar, tested directly: for every pixel of a terrain frame the primary ray's hit / miss decision, hit
    distance, block id (material), face id and hit position (hence the hit voxel's coordinates) from the CUDA path
    (vx_read_hit_records: the records trace_primary_kernel hands to shade_kernel) equal the oracle's intersect_octree results
    BIT FOR BIT — so the budget for rays within epsilon of a voxel boundary (< 1e-4 of all rays) is not drawn on: 0 rays differ.
    Both SVO formats would go through the same records; this is the ESVO world."""
    world, reg = terrain
    w, h = 320, 180
    p = terrain_params(pkg, w, h, shadows=False)
    svo = make_svo(pkg, reg, world, size_mb=max(8, world.size_bytes // 1_000_000 + 8), w=w, h=h, rays=16)
    svo.set_option(14, 7)                                                 # LIFO hand-over: shade_kernel consumes (discards) the records
    svo.render(p, w, h, world=world)
    with pytest.raises(pkg.VxError, match="LIFO"):
        svo.read_hit_records()
    svo.set_option(14, 0)                                                 # the default keeps them
    svo.render(p, w, h, world=world)
    got = svo.read_hit_records()
    svo.close()
    s = helpers.oracle_scene(ora, world, reg)
    q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
    q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
    want = s.primary_hits(pkg.to_vx_render_params(q), w, h)
    hit = want["t"] >= 0
    assert hit.sum() > 1000 and (~hit).sum() > 1000                      # terrain and sky both in view
    assert np.array_equal(got["t"] >= 0, hit)                             # the same pixels hit
    assert (got["t"][~hit] == -1.0).all()
    for f in ("t", "value", "face_id", "pos"):                            # bit patterns, not tolerances
        a, b = np.ascontiguousarray(got[f][hit]), np.ascontiguousarray(want[f][hit])
        assert a.tobytes() == b.tobytes(), (f, int((a != b).sum()))
    # hit voxel coordinates = floor(pos) (svo.esvo.glsl:252-258 clamps pos into the voxel): identical as a consequence
    assert np.array_equal(np.floor(got["pos"][hit]).astype(np.int64), np.floor(want["pos"][hit]).astype(np.int64))
    boundary = int(want["near_boundary"][hit].sum())
    assert boundary < hit.sum()                                           # the flag exists for the report; none of those rays differ either


def test_dirty_update_and_errors(pkg, ora, terrain):
    """graphics::Svo::update with partial ranges (esvo.rs:310-339): edit -> serialize -> update -> frame equals the
    oracle on the new buffer; capacity overflow reports VX_E_CAPACITY instead of panicking (esvo.rs:328-331)."""
    world, reg = terrain
    w, h = 320, 180
    p = terrain_params(pkg, w, h, selected=(-20.0, 50.0, 174.0))
    svo = make_svo(pkg, reg, world, size_mb=world.size_bytes // 1_000_000 + 8, w=w, h=h, rays=16)
    for step in range(3):
        hgt = world.height_at(-10 + step, 174)
        for dy in range(1, 6):
            world.edit_block(-10 + step, hgt + dy, 174, 4)
        world.serialize()
        dirty = world.dirty_ranges()
        assert 1 <= len(dirty) <= 4 and sum(l for _, l in dirty) < world.size_bytes
        svo.update(world)
        assert world.dirty_ranges() == []
        got, got8, want, want8, _ = render_both(pkg, ora, reg, world, p, w, h, svo=svo, use_world=True)
        assert_frames_match(got, got8, want, want8)
    assert svo.get_stats()["used_bytes"] == world.size_bytes and svo.get_stats()["depth"] == world.depth
    svo.close()
    # too-small buffer: update must fail loudly, not corrupt memory
    tiny = pkg.Svo(reg, size_mb=1, max_width=8, max_height=8, max_rays=8)
    w2 = pkg.World(radius=5, center=(-1, 2, 5), seed=1)
    w2.generate(0, 8)
    w2.serialize()
    with pytest.raises(pkg.VxError, match="not large enough"):
        tiny.update(w2)
    tiny.close()


def test_render_before_commit_is_an_error(pkg):
    reg = helpers.shader_test_registry(pkg)
    svo = pkg.Svo(reg, size_mb=1, max_width=8, max_height=8, max_rays=8)
    with pytest.raises(pkg.VxError):
        svo.render(pkg.render_params((0, 0, 0), (0, 0, -1)), 8, 8)
    with pytest.raises(pkg.VxError):
        svo.render(pkg.render_params((0, 0, 0), (0, 0, -1)), 64, 64)   # larger than the reserved framebuffer
    svo.close()


def test_sharded_frames_tile_the_image(pkg, terrain):
    """Image-space shards (multi-GPU partition) are disjoint and their union is the unsharded frame."""
    world, reg = terrain
    w, h = 333, 190   # ragged: not a multiple of the 8x4 warp tile nor the 32x16 macro block
    p = terrain_params(pkg, w, h)
    svo = make_svo(pkg, reg, world, size_mb=world.size_bytes // 1_000_000 + 8, w=w, h=h, rays=16)
    svo.render(p, w, h, world=world)
    full = svo.read_rgba32f()
    for n in (2, 3, 8):
        acc = np.full_like(full, np.nan)
        covered = np.zeros((h, w), dtype=np.int32)
        for r in range(n):
            # poison the device frame through a full render of a different view, then render the shard
            svo.render(terrain_params(pkg, w, h, shadows=False), w, h, world=world)
            before = svo.read_rgba32f()
            svo.render(p, w, h, shard=(r, n), world=world)
            part = svo.read_rgba32f()
            changed = (part != before).any(axis=2)
            acc[changed] = part[changed]
            covered += changed
        assert covered.max() <= 1
        same = np.isnan(acc).all(axis=2)   # pixels identical in both views are indistinguishable; accept them from `full`
        acc[same] = full[same]
        assert acc.tobytes() == full.tobytes(), n
    svo.close()


def test_pipelined_render_readback(pkg, terrain):
    """vx_render_read_rgba8 (banded render overlapped with the RGBA8 copy to the host) returns exactly the bytes of
    vx_render + vx_read_frame_rgba8, for any band count, ragged frame sizes and shards."""
    import torch
    world, reg = terrain
    for (w, h) in ((333, 190), (640, 368)):
        p = terrain_params(pkg, w, h)
        svo = make_svo(pkg, reg, world, size_mb=world.size_bytes // 1_000_000 + 8, w=w, h=h, rays=16)
        svo.render(p, w, h, world=world)
        want = svo.read_rgba8()
        q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
        q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
        vxp = pkg.to_vx_render_params(q)
        out = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()
        for bands in (1, 2, 4, 7, 16, 99):
            out.fill_(7)
            svo.render_read_rgba8(vxp, w, h, out.data_ptr(), bands=bands)
            assert out.numpy().tobytes() == want.tobytes(), (w, h, bands)
        # RGBA8 output mode (vx_set_option 8): the kernels store the rounded pixels themselves — same bytes, no conversion pass
        svo.set_option(pkg.OPT_RGBA8_OUT, 1)
        svo.render(p, w, h, world=world)
        assert svo.read_rgba8().tobytes() == want.tobytes()
        out.fill_(9)
        svo.render_read_rgba8(vxp, w, h, out.data_ptr(), bands=3)
        assert out.numpy().tobytes() == want.tobytes()
        with pytest.raises(pkg.VxError):
            svo.read_rgba32f()
        svo.set_option(pkg.OPT_RGBA8_OUT, 0)
        # shard 1 of 3: owned pixels equal the full frame, the others keep what the previous full render left in the device frame
        svo.render_read_rgba8(vxp, w, h, out.data_ptr(), bands=3, shard=(1, 3))
        assert out.numpy().tobytes() == want.tobytes()
        # two frames in flight (vx_render_read_rgba8_begin x 2 before the first _end): a second camera renders into the second device
        # frame while the first frame's copies may still run; _end hands the frames back oldest first; a third _begin is refused
        p2 = terrain_params(pkg, w, h, shadows=False)
        q2 = pkg.VxhRenderParams.from_buffer_copy(bytes(p2))
        q2.cam_pos = (C.c_float * 3)(*world.cnv_block_pos((-20.0, 90.0, 170.0)))
        vxp2 = pkg.to_vx_render_params(q2)
        svo.render_raw(vxp2, w, h)
        want2 = svo.read_rgba8()
        assert want2.tobytes() != want.tobytes()
        out2 = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()
        for rounds in range(3):
            out.fill_(1); out2.fill_(2)
            svo.render_read_rgba8_begin(vxp, w, h, out.data_ptr(), bands=1 + rounds)
            svo.render_read_rgba8_begin(vxp2, w, h, out2.data_ptr(), bands=2)
            with pytest.raises(pkg.VxError, match="in flight"):
                svo.render_read_rgba8_begin(vxp, w, h, out.data_ptr(), bands=1)
            svo.render_read_rgba8_end()
            assert out.numpy().tobytes() == want.tobytes(), (w, h, rounds)
            svo.render_read_rgba8_end()
            assert out2.numpy().tobytes() == want2.tobytes(), (w, h, rounds)
            svo.render_read_rgba8_end()                                   # nothing in flight: a no-op
        assert svo.read_rgba8().tobytes() == want2.tobytes()              # vx_read_frame_rgba8 follows the frame rendered last
        # a stream of frames, one in flight across iterations (the bench's e2e loop)
        outs = (out, out2)
        for k in range(6):
            svo.render_read_rgba8_begin(vxp if k % 2 == 0 else vxp2, w, h, outs[k & 1].data_ptr(), bands=1)
            if k:
                svo.render_read_rgba8_end()
                assert outs[(k - 1) & 1].numpy().tobytes() == (want if (k - 1) % 2 == 0 else want2).tobytes(), k
        svo.render_read_rgba8_end()
        assert out2.numpy().tobytes() == want2.tobytes()
        svo.close()


# ---------------------------------------------------------------------------------------------- full BASELINE sizes --

def test_full_size_4k_frame_vs_oracle(pkg, ora):
    """BASELINE configs[2] at its real size: r=20 LOD world, 3840x2160, primary + shadow rays. The oracle renders the whole
    frame on the host cores in about a second, so this is a direct comparison, not a property test: RGB8 within 1 LSB on
    all 8.3 M pixels, ray / step / push / leaf / texel counters identical, hit mask identical."""
    import bench
    args = type("A", (), dict(radius=20, no_lod=False, width=3840, height=2160, no_shadows=False, terrain="reference"))()
    world, _ = bench.build_world(pkg, args)
    reg = pkg.content_registry(pkg.load_atlas())
    vxp = bench.frame_params(pkg, world, args)
    W, H = args.width, args.height
    svo = pkg.Svo(reg, size_mb=int(world.size_bytes // 1_000_000 + 64), max_width=W, max_height=H, max_rays=16)
    world.mark_all_dirty()
    svo.update(world)
    svo.set_option(pkg.OPT_COUNT, 1)
    svo.render_raw(vxp, W, H)
    st = svo.frame_stats(0)
    got, got8 = svo.read_rgba32f(), svo.read_rgba8()
    want, cnt = helpers.oracle_scene(ora, world, reg).render(vxp, W, H)
    want8 = ora.to_rgba8(want)
    d8 = np.abs(got8.astype(np.int16) - want8.astype(np.int16))
    assert d8.max() <= 1, (int(d8.max()), int((d8 > 1).sum()))
    assert float(np.abs(got - want).max()) <= 2e-6
    for k in ("primary_rays", "shadow_rays", "steps", "pushes", "leaf_tests", "tex_fetches"):
        assert st[k] == cnt[k], (k, st[k], cnt[k])
    assert st["primary_rays"] == W * H and st["shadow_rays"] > 5_000_000
    # determinism + partition at full size: 8 shards rendered one after the other rebuild the identical frame
    svo.set_option(pkg.OPT_COUNT, 0)
    svo.render_raw(vxp, W, H)
    assert svo.read_rgba32f().tobytes() == got.tobytes()
    svo.render_raw(bench.frame_params(pkg, world, type("A", (), dict(radius=20, no_lod=False, width=W, height=H, no_shadows=True, terrain="reference"))()), W, H)
    for r in range(8):
        svo.render_raw(vxp, W, H, shard=(r, 8))
    assert svo.read_rgba32f().tobytes() == got.tobytes()
    svo.close()


def test_full_size_16m_picker_rays_vs_oracle(pkg, ora):
    """BASELINE configs[3] at its real size: 16 Mi random picker rays against the r=40 no-LOD world of the reference's generator (0.8 GB SVO, depth 12).
    Every one of the 16 Mi 48-byte results is byte-identical to the oracle's; counters identical."""
    import bench
    world = pkg.World(radius=40, center=(-1, 2, 5), seed=1, no_lod=True, terrain="reference")
    world.generate(0, 8)
    world.serialize()
    assert world.depth == 12 and world.size_bytes > 700_000_000
    reg = pkg.content_registry(pkg.load_atlas())
    n = 1 << 24
    svo = pkg.Svo(reg, size_mb=int(world.size_bytes // 1_000_000 + 64), max_width=32, max_height=16, max_rays=n)
    world.mark_all_dirty()
    svo.update(world)
    scene = helpers.oracle_scene(ora, world, reg)
    for max_dst, bin in ((-1.0, 7 | 16), (30.0, 0)):   # binned, and in task order (the default)
        tasks = bench.picker_tasks(pkg, world, 40, n, seed=0, max_dst=max_dst)
        svo.set_option(15, bin)
        svo.set_option(pkg.OPT_COUNT, 1)
        got = svo.raycast_tasks(tasks)
        st = svo.frame_stats(1)
        want, cnt = scene.raycast(tasks)
        assert got.tobytes() == want.tobytes(), (max_dst, int((got["dst"] != want["dst"]).sum()))
        assert st["steps"] == cnt["steps"] and st["pushes"] == cnt["pushes"] and st["leaf_tests"] == cnt["leaf_tests"]
        hits = int((got["dst"] > 0).sum())
        assert 0 < hits < n
    svo.close()


# -------------------------------------------------------------------------------------------------- edge cases --

def test_edge_cases(pkg, ora, reg):
    """Empty / degenerate inputs: zero rays, one ray, a 1x1 and a 1-pixel-wide frame, rays that start outside the octree and
    never enter it, axis-parallel directions (epsilon clamp), a ray budget that runs out (MAX_STEPS, svo.esvo.glsl:152)."""
    w = helpers.shader_test_world(pkg, [(x, 0, z, 1) for x in range(32) for z in range(32)] + [(5, 5, 5, 2), (31, 31, 31, 1)])
    s = helpers.oracle_scene(ora, w, reg)
    svo = make_svo(pkg, reg, w, size_mb=2, w=64, h=64, rays=4096)
    # zero rays: nothing to do, no error
    assert len(svo.raycast_tasks(np.zeros(0, dtype=pkg.TASK_DTYPE))) == 0
    # single rays, incl. pathological ones
    cases = [((16.0, 40.0, 16.0), (0.0, -1.0, 0.0), -1.0),      # straight down, two zero components -> epsilon clamp
             ((-50.0, 10.0, 16.0), (-1.0, 0.0, 0.0), -1.0),     # outside, pointing away
             ((16.0, 1.5, 16.0), (1.0, 0.0, 0.0), -1.0),        # skimming 0.5 above the floor along +x
             ((5.5, 5.5, 5.5), (0.0, 1.0, 0.0), -1.0),          # origin inside a voxel
             ((16.0, 40.0, 16.0), (0.0, -1.0, 0.0), 0.0),       # max_dst = 0
             ((1e6, 1e6, 1e6), (-1.0, -1.0, -1.0), -1.0),       # far outside, pointing at the world
             ((16.0, float("nan"), 16.0), (0.0, -1.0, 0.0), -1.0)]
    for pos, d, md in cases:
        t = np.zeros(1, dtype=pkg.TASK_DTYPE)
        t["pos"], t["dir"], t["max_dst"] = pos, np.array(d, np.float32) / np.linalg.norm(d), md
        want, _ = s.raycast(t)
        got = svo.raycast_tasks(t)
        if np.isnan(pos).any():
            # NaN in -> the shader's result is undefined; require the same non-NaN fields and NaN in the same places (the bit
            # pattern of a generated NaN differs between x86 and the GPU)
            for f in ("dst", "inside_voxel", "pos", "normal"):
                assert np.array_equal(got[f], want[f], equal_nan=True), (f, got, want)
            continue
        assert got.tobytes() == want.tobytes(), (pos, d, md, got, want)
    # ragged tiny frames
    for (fw, fh) in ((1, 1), (1, 64), (64, 1), (33, 17)):
        p = pkg.render_params(cam_pos=(16.0, 20.0, 50.0), cam_fwd=(0.0, -0.4, -1.0), fov_y_deg=72.0, aspect=fw / fh)
        got, got8, want, want8, _ = render_both(pkg, ora, reg, w, p, fw, fh, svo=svo)
        assert np.abs(got8.astype(int) - want8.astype(int)).max() <= 1 and got.shape == (fh, fw, 4)
    svo.close()
    # MAX_STEPS (svo.esvo.glsl:18,152,392): 32 chunks in a row, each with isolated voxels on every other cell of its floor. A ray
    # skimming the empty row between them must descend to voxel level every two cells: ~2 iterations per voxel. From x = 0.5 the
    # wall at x = 1023 is more than 1000 iterations away and is NOT found (budget exhausted -> miss); from x = 900.5 it is hit.
    w2 = pkg.World()
    dust = [(x, 0, z, 1) for x in range(0, 32, 2) for z in range(0, 32, 2)]
    for cx in range(32):
        blocks = list(dust) + ([(31, y, z, 2) for y in range(4) for z in range(4)] if cx == 31 else [])
        w2.set_leaf_blocks((cx, 0, 0), blocks, uid=100 + cx, lod=5, compact=True)
    w2.serialize()
    s2 = helpers.oracle_scene(ora, w2, reg)
    svo2 = make_svo(pkg, reg, w2, size_mb=4, w=64, h=64, rays=1 << 12)
    for x0, frames, hit in ((0.5, 1000, False), (512.5, 1000, False), (900.5, None, True)):
        o = s2.debug_cast((x0, 0.5, 1.5), (1.0, 0.0, 0.0), -1.0, False, frames_cap=4)
        c = svo2.debug_cast((x0, 0.5, 1.5), (1.0, 0.0, 0.0), -1.0, False, frames_cap=4)
        assert_same_cast(o, c)
        assert (c[0].t > 0) == hit and (frames is None or c[2] == frames), (x0, c[2], c[0].t)
    tasks = np.zeros(1 << 12, dtype=pkg.TASK_DTYPE)
    tasks["max_dst"] = -1.0
    tasks["pos"] = np.stack([np.linspace(0.5, 1000.5, 1 << 12), np.full(1 << 12, 0.5), 1.5 + 2.0 * (np.arange(1 << 12) % 16)], axis=1)
    tasks["dir"] = (1.0, 0.0, 0.0)
    svo2.set_option(pkg.OPT_COUNT, 1)
    want, cnt = s2.raycast(tasks)
    got = svo2.raycast_tasks(tasks)
    st = svo2.frame_stats(1)
    assert got.tobytes() == want.tobytes()
    assert st["steps"] == cnt["steps"] and cnt["steps"] > 500 * len(tasks)
    assert 0 < (got["dst"] > 0).sum() < len(tasks)          # some rays run out of budget, some reach the wall
    # a frame looking down the row: pixels whose rays exhaust the budget are sky in both implementations
    p = pkg.render_params(cam_pos=(0.5, 0.6, 1.5), cam_fwd=(1.0, -0.0005, 0.0), fov_y_deg=20.0, aspect=1.0)
    got, got8, want, want8, _ = render_both(pkg, ora, reg, w2, p, 64, 64, svo=svo2)
    assert np.abs(got8.astype(int) - want8.astype(int)).max() <= 1
    svo2.close()


# ------------------------------------------------------------------------------------------------------------- CSVO --

def test_csvo_golden_traces_and_results_on_gpu(pkg, ora, reg):
    """The CSVO kernels (VX_FLAG_SVO_CSVO) against the oracle and the reference's csvo_tests goldens
    (svo_shader_tests.rs:756-1224): step traces with byte pointers / depth / crossed_boundary / next_ptr, and all result cases."""
    from test_oracle_golden import check_frames_csvo
    for g in (gv.CSVO_TRAVERSAL, gv.CSVO_HIGHER_COORDS):
        o, c = cast_both(pkg, ora, reg, g["blocks"], g["pos"], g["dir"], g["max_dst"], g["cast_translucent"], g.get("svo_pos", (0, 0, 0)), fmt=1)
        assert_same_cast(o, c)
        check_frames_csvo(c[1], c[2], g["frames"])
        check_result(c[0], g["result"])
    g = gv.ALL_AXES
    for name, pos, d, t, face, hit_pos, uv in g["cases"]:
        o, c = cast_both(pkg, ora, reg, g["blocks"], pos, d, 100.0, False, fmt=1)
        assert_same_cast(o, c)
        check_result(c[0], {"t": t, "value": 1, "face_id": face, "pos": hit_pos, "uv": uv, "color": g["color"], "inside_voxel": False}, name=name)
    g = gv.UV_COORDS
    for pos, d, uv, color in g["cases"]:
        o, c = cast_both(pkg, ora, reg, g["blocks"], pos, d, 32.0, False, fmt=1)
        assert_same_cast(o, c)
        assert close(c[0].as_dict()["uv"], uv) and close(c[0].as_dict()["color"], color)
    g = gv.TRANSLUCENT
    for name, pos, translucent, exp in g["cases"]:
        o, c = cast_both(pkg, ora, reg, g["blocks"], pos, g["dir"], 32.0, translucent, fmt=1)
        assert_same_cast(o, c)
        assert c[0].as_dict()["value"] == exp["value"] and c[0].as_dict()["face_id"] == exp["face_id"], name
    g = gv.INSIDE_LEAF
    for name, pos, d, exp in g["cases"]:
        o, c = cast_both(pkg, ora, reg, g["blocks"], pos, d, 32.0, False, fmt=1)
        assert_same_cast(o, c)
        check_result(c[0], exp, name=name)


def test_csvo_picker_random_rays_bit_exact(pkg, ora, reg):
    """200k incoherent rays per scene on CSVO buffers, origins inside voxels included (the out-of-spec descent follows the
    oracle's stated policy): results byte-identical, counters identical."""
    for name, blocks, svo_pos in small_scenes(pkg):
        w = helpers.shader_test_world(pkg, blocks, svo_pos, fmt=1)
        s = helpers.oracle_scene(ora, w, reg)
        svo = make_svo(pkg, reg, w, size_mb=4, w=8, h=8, rays=1 << 18)
        size = 32 * (max(svo_pos) + 1)
        for max_dst, seed in ((-1.0, 1), (30.0, 2)):
            tasks = helpers.random_tasks(pkg, 200_000, -8.0, size + 8.0, max_dst, seed)
            tasks["dir"][::17, 1] = 0.0
            tasks["pos"][::5] = np.floor(tasks["pos"][::5]) + 0.5
            want, ocnt = s.raycast(tasks)
            svo.set_option(pkg.OPT_COUNT, 1)
            got = svo.raycast_tasks(tasks)
            assert got.tobytes() == want.tobytes(), (name, max_dst, int((got["dst"] != want["dst"]).sum()))
            st = svo.frame_stats(1)
            assert st["steps"] == ocnt["steps"] and st["pushes"] == ocnt["pushes"] and st["leaf_tests"] == ocnt["leaf_tests"], (st, ocnt)
        svo.close()


def test_csvo_render_terrain(pkg, ora):
    """Generated terrain with LOD chunks serialized as CSVO: frame (shading, trilinear textures, shadows, highlight) within
    1 LSB of the oracle's CSVO shader, counters equal; dirty-range update through the 8-byte CSVO header; and the same
    frame as the ESVO path wherever no ray starts inside a voxel."""
    reg = pkg.content_registry(pkg.load_atlas())
    frames = {}
    for fmt in (1, 0):
        world = pkg.World(radius=5, center=(-1, 2, 5), seed=1, fmt=fmt)
        world.generate(0, 8)
        world.serialize()
        w, h = 480, 270
        p = terrain_params(pkg, w, h, selected=(-20.0, 50.0, 174.0))
        svo = make_svo(pkg, reg, world, size_mb=world.size_bytes // 1_000_000 + 8, w=w, h=h, rays=16)
        svo.set_option(pkg.OPT_COUNT, 1)
        got, got8, want, want8, cnt = render_both(pkg, ora, reg, world, p, w, h, svo=svo, use_world=True)
        assert_frames_match(got, got8, want, want8)
        st = svo.frame_stats(0)
        for k in ("primary_rays", "shadow_rays", "steps", "pushes", "leaf_tests", "tex_fetches"):
            assert st[k] == cnt[k], (fmt, k, st, cnt)
        frames[fmt] = got
        if fmt == 1:
            for step in range(2):   # edit -> re-serialize -> partial update -> same as the oracle on the new buffer
                hgt = world.height_at(-10 + step, 174)
                for dy in range(1, 6):
                    world.edit_block(-10 + step, hgt + dy, 174, 4)
                world.serialize()
                assert sum(l for _, l in world.dirty_ranges()) < world.size_bytes
                svo.update(world)
                got, got8, want, want8, _ = render_both(pkg, ora, reg, world, p, w, h, svo=svo, use_world=True)
                assert_frames_match(got, got8, want, want8)
        svo.close()
    differing = (np.abs(frames[0] - frames[1]).max(axis=2) > 0).mean()
    assert differing < 1e-3, differing


# ------------------------------------------------------------------------------------- chunk serialization on the GPU --

def test_gpu_chunk_serialization_matches_host(pkg):
    """vx_serialize_chunks_esvo against the host serializer (itself pinned on esvo.rs:561-1228): random chunks
    of every density at every LOD, then every chunk of a generated-terrain world against the bytes the host put into its
    RangeBuffer; the chunk's SerializationResult (child mask, leaf mask, depth) included."""
    reg = helpers.shader_test_registry(pkg)
    svo = pkg.Svo(reg, size_mb=1, max_width=8, max_height=8, max_rays=8)
    rng = np.random.default_rng(21)
    blocks, lods = [], []
    for density in (0.0, 0.0005, 0.02, 0.3, 0.9, 1.0):
        for lod in (0, 1, 2, 3, 4, 5):
            blocks.append(((rng.random(32768) < density) * rng.integers(1, 1 << 20, 32768)).astype(np.uint32))
            lods.append(lod)
    corner = np.zeros(32768, np.uint32); corner[31] = 1; corner[31 * 32] = 2; corner[31 * 1024] = 3      # the reference's KAT chunk
    blocks.append(corner); lods.append(5)
    blocks = np.stack(blocks)
    infos, rec, ms = svo.serialize_chunks(blocks, lods)
    assert ms > 0
    # ranges are disjoint and tile the output
    order = np.argsort(infos["offset_bytes"], kind="stable")
    nz = [i for i in order if infos["length_bytes"][i] > 0]
    assert sum(int(infos["length_bytes"][i]) for i in nz) == len(rec)
    end = 0
    for i in nz:
        assert int(infos["offset_bytes"][i]) == end
        end += int(infos["length_bytes"][i])
    out = np.zeros(4681 * 12, dtype=np.uint32)
    res = (C.c_uint8 * 3)()
    for i in range(len(blocks)):
        n = pkg.host().vxh_serialize_dense(blocks[i].ctypes.data, lods[i], out.ctypes.data, len(out), res)
        o, l = int(infos["offset_bytes"][i]), int(infos["length_bytes"][i])
        assert l == n * 4, (i, lods[i], l, n * 4)
        assert rec[o:o + l].tobytes() == out[:n].tobytes(), (i, lods[i])
        assert (infos["child_mask"][i], infos["leaf_mask"][i], infos["depth"][i]) == tuple(res), (i, lods[i])
    # a whole generated world: the GPU redoes what the host's job system did for every chunk
    world = pkg.World(radius=8, center=(-1, 2, 5), seed=1)      # LOD 5 within 6 chunks of the centre, LOD 4 beyond (chunkloader.rs:127-134)
    world.generate(0, 8)
    world.serialize()
    image = world.range_bytes()
    chunks = world.chunks()
    cb = [world.chunk_blocks(c) for c in chunks]
    infos, rec, _ = svo.serialize_chunks(np.stack([b for b, _ in cb]), [l for _, l in cb])
    assert len(chunks) > 100 and len({l for _, l in cb}) > 1          # several LODs present
    for c, info in zip(chunks, infos):
        off, length = world.chunk_range(c)
        assert int(info["length_bytes"]) == length
        assert rec[int(info["offset_bytes"]):int(info["offset_bytes"]) + length].tobytes() == image[off:off + length].tobytes(), tuple(c)
    svo.close()


def test_world_rebuilt_from_gpu_serialized_chunks(pkg, terrain):
    """Integration of the device-side serializer: the chunk records of a committed world are wiped on the GPU and put back from
    vx_serialize_chunks_esvo output with vx_svo_write_device (device to device, no host round trip) at the ranges the host's
    RangeBuffer assigned — the frame is byte-identical to the one rendered from the host-serialized buffer."""
    import torch
    world, reg = terrain
    w, h = 480, 270
    p = terrain_params(pkg, w, h)
    svo = make_svo(pkg, reg, world, size_mb=world.size_bytes // 1_000_000 + 8, w=w, h=h, rays=16)
    svo.render(p, w, h, world=world)
    want = svo.read_rgba32f()
    chunks = world.chunks()
    zeros = torch.zeros(4681 * 48, dtype=torch.uint8, device="cuda")
    for c in chunks:                                   # wipe every chunk's records (the world root stays): nothing to hit
        off, length = world.chunk_range(c)
        svo.svo_write_device(off, zeros.data_ptr(), length)
    svo.render(p, w, h, world=world)
    wiped = svo.read_rgba32f()
    assert (wiped != want).any(axis=2).mean() > 0.3
    cb = [world.chunk_blocks(c) for c in chunks]
    infos, _, _ = svo.serialize_chunks(np.stack([b for b, _ in cb]), [l for _, l in cb], want_records=False)
    base = svo.serialize_records_ptr()
    for c, info in zip(chunks, infos):
        off, length = world.chunk_range(c)
        assert int(info["length_bytes"]) == length
        svo.svo_write_device(off, base + int(info["offset_bytes"]), length)
    svo.render(p, w, h, world=world)
    assert svo.read_rgba32f().tobytes() == want.tobytes()
    svo.close()


# ------------------------------------------------------------------ BASELINE configs[0]: the bundled Minecraft world --

def test_mc_world_1280x720_primary_frame(pkg, ora):
    """configs[0]: fixed test world from assets/worlds (tests/golden/mc_world.npz), SVO serialized on the CPU, one 1280x720
    primary-ray frame from a fixed camera, no shadows, diffed against the reference shader's restatement: RGB8 within 1 LSB,
    ray / step / push / leaf / texel counters identical. Then the same frame with shadow rays, the CSVO build of the same
    world, and 200 k random picker rays through forest and water (translucent pass-through) byte-identical."""
    reg = pkg.content_registry(pkg.load_atlas())
    W, H = 1280, 720
    for fmt in (0, 1):
        world = helpers.mc_world(pkg, fmt)
        svo = make_svo(pkg, reg, world, size_mb=world.size_bytes // 1_000_000 + 8, w=W, h=H, rays=1 << 18)
        svo.set_option(pkg.OPT_COUNT, 1)
        for shadows in (False, True):
            p = helpers.mc_params(pkg, W, H, shadows=shadows)
            got, got8, want, want8, cnt = render_both(pkg, ora, reg, world, p, W, H, svo=svo, use_world=True)
            assert_frames_match(got, got8, want, want8)
            st = svo.frame_stats(0)
            for k in ("primary_rays", "shadow_rays", "steps", "pushes", "leaf_tests", "tex_fetches"):
                assert st[k] == cnt[k], (fmt, shadows, k, st, cnt)
            assert cnt["primary_rays"] == W * H and (cnt["shadow_rays"] > 0) == shadows
        # picker rays from inside the fixture's volume (SVO space: chunks 1..9 x 4..6 x 1..8 of the 13^3 window)
        s = helpers.oracle_scene(ora, world, reg)
        rng = np.random.default_rng(17)
        tasks = helpers.random_tasks(pkg, 200_000, 0.0, 1.0, -1.0, seed=17)
        tasks["pos"] = np.stack([rng.uniform(64, 288, 200_000), rng.uniform(5 * 32 + 20, 6 * 32 + 40, 200_000), rng.uniform(64, 256, 200_000)],
                                axis=1).astype(np.float32)
        tasks["max_dst"][::2] = 48.0
        want_r, ocnt = s.raycast(tasks)
        got_r = svo.raycast_tasks(tasks)
        assert got_r.tobytes() == want_r.tobytes(), (fmt, int((got_r["dst"] != want_r["dst"]).sum()))
        assert 0.2 < (got_r["dst"] > 0).mean() < 0.95
        svo.close()
