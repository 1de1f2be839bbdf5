"""The product's CUDA kernels run on the CPU. tests/emu compiles voxel-rs_b200/csrc/{traverse,kernels}.cuh with g++ (a stand-in
cuda_runtime.h: every CUDA thread a fiber, warp collectives and __syncthreads as rendezvous, one CTA at a time) and issues the same
kernel sequence as vx_render / vx_raycast. What comes out is compared with the oracle like the -m gpu tests do on the B200, so the
logic of the kernels (warp-level work fetch, refill, hit records, shadow-list compaction, traversal, shading, both SVO formats) is
checked on every CPU run. It says nothing about timing, the memory model or the PTX paths (those stay with the GPU tests); the GPU
build is unaffected: the only source difference is behind `#ifdef VX_HOST_EMULATION`, which nvcc never defines."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
ROOT = os.path.dirname(HERE)


@pytest.fixture(autouse=True)
def _oracle_follows_the_products_clip(ora):
    """The kernels stop a ray once it has left the occupied box of the world (vx_set_option 12, on by default): results are the
    shader's bit for bit, the iteration counters are smaller. The oracle restates that extension when asked, so that counters can
    still be compared one to one; it is switched off again for every other test (the oracle's own goldens run the shader as written)."""
    ora.set_clip(True)
    yield
    ora.set_clip(False)


@pytest.fixture(scope="module")
def emu(pkg):
    src = [os.path.join(EMU, "emu_harness.cpp"), os.path.join(EMU, "cuda_runtime.h"), os.path.join(ROOT, "include", "voxelrt.h"),
           os.path.join(ROOT, "voxel-rs_b200", "csrc", "kernels.cuh"), os.path.join(ROOT, "voxel-rs_b200", "csrc", "traverse.cuh"),
           os.path.join(ROOT, "voxel-rs_b200", "csrc", "chunks.cuh")]
    out = os.path.join(EMU, "libkernels_emu.so")
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in src):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        fma = ["-mfma"] if "fma" in open("/proc/cpuinfo").read().split() else []
        # the numeric contract of the CUDA build (--fmad=false, IEEE) and of the oracle: no contraction, no fast-math
        subprocess.check_call([cxx, "-O1", "-std=c++17", "-fPIC", "-shared", *fma, "-ffp-contract=off", "-fno-fast-math", "-fno-extern-tls-init", "-fvisibility=hidden", "-Wl,-Bsymbolic",
                               "-Wno-unknown-pragmas",
                               "-I", EMU, "-o", out, src[0]])
    lib = C.CDLL(out)
    lib.emu_render.restype = C.c_int
    lib.emu_raycast.restype = C.c_int
    return lib


def scene_args(world, reg):
    buf = world.gpu_buffer()
    mats = reg.materials()
    tex, mips = reg.textures()
    tex = np.ascontiguousarray(tex)
    layers, th, tw = tex.shape[0], tex.shape[1], tex.shape[2]
    keep = (buf, mats, tex)
    return keep, [C.c_void_p(buf.ctypes.data), C.c_uint64(len(buf)), C.c_int(world.fmt), C.c_uint32(world.depth), C.c_void_p(mats.ctypes.data),
                  C.c_uint32(len(mats)), C.c_void_p(tex.ctypes.data), C.c_uint32(tw), C.c_uint32(th), C.c_uint32(layers), C.c_uint32(mips)]


def opts(refill=1, shadow_refill=0, ctas=3, count=1, rgba8=0, tma=0, rank=0, size=1, bands=1, overlap=0, lifo=1):
    return (C.c_uint32 * 10)(refill, shadow_refill, ctas, count, rgba8, tma | (lifo << 1), rank, size, bands, overlap)


def emu_render(emu, pkg, world, reg, vxp, w, h, **kw):
    keep, args = scene_args(world, reg)
    frame = np.zeros((h, w, 4), np.float32)
    frame8 = np.zeros((h, w, 4), np.uint8)
    cnt = (C.c_uint64 * 6)()
    rc = emu.emu_render(*args, C.byref(vxp), C.c_uint32(w), C.c_uint32(h), opts(**kw), C.c_void_p(frame.ctypes.data), C.c_void_p(frame8.ctypes.data), cnt)
    assert rc == 0
    names = ("primary_rays", "shadow_rays", "steps", "pushes", "leaf_tests", "tex_fetches")
    return frame, frame8, dict(zip(names, [int(v) for v in cnt]))


def oracle_render(pkg, ora, world, reg, vxp, w, h):
    img, cnt = helpers.oracle_scene(ora, world, reg).render(vxp, w, h)
    return img, ora.to_rgba8(img), cnt


def world_params(pkg, world, w, h, shadows=True, selected=None):
    p = pkg.render_params(cam_pos=(-24.0, 80.0, 174.0), cam_fwd=(1.0, -0.3, 0.0), fov_y_deg=72.0, aspect=w / h, render_shadows=shadows,
                          selected_voxel=selected)
    q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
    q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
    if selected is not None:
        q.selected_voxel = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.selected_voxel)))
    return pkg.to_vx_render_params(q)


@pytest.fixture(scope="module")
def terrains(pkg):
    out = {}
    for fmt in (0, 1):
        w = pkg.World(radius=5, center=(-1, 2, 5), seed=1, fmt=fmt)
        w.generate(0, 8)
        w.serialize()
        out[fmt] = w
    return out, pkg.content_registry(pkg.load_atlas())


@pytest.mark.parametrize("fmt", [0, 1])
def test_emulated_frame_equals_oracle(emu, pkg, ora, terrains, fmt):
    """Terrain frame with LOD chunks, the trilinear texture path, shadows and the highlighted voxel through the three kernels: float
    frame, RGBA8 frame and all six counters equal the oracle's — for refill thresholds 1 / 8 / 32, 1 and 5 persistent CTAs."""
    worlds, reg = terrains
    world = worlds[fmt]
    w, h = 72, 44                                  # ragged: 3 macro columns (one partial), 3 macro rows (one partial)
    vxp = world_params(pkg, world, w, h, selected=(-20.0, 50.0, 174.0))
    want, want8, cnt = oracle_render(pkg, ora, world, reg, vxp, w, h)
    assert cnt["shadow_rays"] > 0 and cnt["tex_fetches"] > cnt["leaf_tests"]
    # lifo: the LIFO hand-over of the wavefront buffers (vx_set_option 14) — consumers walk the buffers backwards and discard what they
    # have read; the emulator poisons every discarded line, so a line dropped too early (or somebody else's) changes the frame
    for refill, ctas, lifo in ((1, 3, 1), (8, 1, 1), (32, 5, 1), (1, 2, 0)):
        got, _, c = emu_render(emu, pkg, world, reg, vxp, w, h, refill=refill, ctas=ctas, lifo=lifo)
        assert got.tobytes() == want.tobytes(), (fmt, refill, float(np.abs(got - want).max()))
        assert c == cnt, (fmt, refill, c, cnt)
    # RGBA8 output mode (vx_set_option 8 / vx_render_read_rgba8) and the staged tile write-back (vx_set_option 9)
    _, got8, _ = emu_render(emu, pkg, world, reg, vxp, w, h, rgba8=1)
    assert got8.tobytes() == want8.tobytes()
    got, _, _ = emu_render(emu, pkg, world, reg, vxp, w, h, tma=1)
    assert got.tobytes() == want.tobytes()
    # the overlapped wavefront's strip-completion flags: every strip's producers count exactly the pixels its consumer waits for
    # (emu_render fails if a wait gives up), also with mid-tile refills and on the ragged frame edges
    for refill in (1, 8):
        got, _, c = emu_render(emu, pkg, world, reg, vxp, w, h, refill=refill, overlap=1)
        assert got.tobytes() == want.tobytes() and c == cnt
    # the banded wavefront of vx_render_read_rgba8 (per-band macro-block ranges and work counters), also sharded
    _, got8, c = emu_render(emu, pkg, world, reg, vxp, w, h, rgba8=1, bands=3)
    assert got8.tobytes() == want8.tobytes() and c == cnt
    union = np.zeros((h, w, 4), np.uint8)
    for rank in range(2):
        _, part, _ = emu_render(emu, pkg, world, reg, vxp, w, h, rgba8=1, bands=2, rank=rank, size=2)
        mine = part.view(np.uint32)[..., 0] != 0xdeadbeef
        union[mine] = part[mine]
    assert union.tobytes() == want8.tobytes()


def test_emulated_shards_tile_the_frame(emu, pkg, ora, terrains):
    """Image-space shards (macro block m belongs to rank m % size): every pixel is written by exactly one of the 3 ranks, with the
    value of the unsharded frame; rays are counted once."""
    worlds, reg = terrains
    world = worlds[0]
    w, h = 100, 40
    vxp = world_params(pkg, world, w, h)
    want, _, cnt = oracle_render(pkg, ora, world, reg, vxp, w, h)
    for rows in (0, 0x80000000):                   # interleaved macro blocks / VX_SHARD_ROWS: whole 16-pixel stripes
        union = np.full((h, w, 4), -1.0, np.float32)
        total = {k: 0 for k in cnt}
        for rank in range(3):
            got, _, c = emu_render(emu, pkg, world, reg, vxp, w, h, rank=rank, size=3 | rows, bands=2 if rows else 1)
            mine = got[..., 3] != -1.0             # untouched pixels keep the harness's -1 fill (alpha is never negative)
            assert not (mine & (union[..., 3] != -1.0)).any()
            if rows:                               # a shard's pixels are exactly the stripes y // 16 % 3 == rank
                assert np.array_equal(mine, np.repeat(((np.arange(h) // 16) % 3 == rank)[:, None], w, axis=1))
            union[mine] = got[mine]
            for k in c:
                total[k] += c[k]
        assert union.tobytes() == want.tobytes() and total == cnt


def test_emulated_reference_scene(emu, pkg, ora):
    """svo_tests::render's scene (src/graphics/svo.rs:342-399) at a quarter of its size: normal maps, specular, shadows, highlight."""
    reg = helpers.svo_render_test_registry(pkg, pkg.load_atlas())
    world = pkg.World()
    world.set_leaf_blocks((0, 0, 0), helpers.svo_render_test_blocks(), compact=False)
    world.serialize()
    w, h = 160, 122
    vxp = pkg.to_vx_render_params(helpers.svo_render_test_params(pkg, w, h))
    want, _, cnt = oracle_render(pkg, ora, world, reg, vxp, w, h)
    got, _, c = emu_render(emu, pkg, world, reg, vxp, w, h)
    assert got.tobytes() == want.tobytes() and c == cnt


@pytest.mark.parametrize("fmt", [0, 1])
def test_emulated_picker_equals_oracle(emu, pkg, ora, terrains, fmt):
    """trace_picker_kernel (picker.glsl): 3000 random rays (ragged last run), unlimited and max_dst = 30, refill 24 and 1."""
    import bench
    worlds, reg = terrains
    world = worlds[fmt]
    scene = helpers.oracle_scene(ora, world, reg)
    keep, args = scene_args(world, reg)
    # bin = ray binning (vx_set_option 15): the batch is traced in Z-order of the origin cells (bits per axis, +16: direction octant below);
    # results stay in task order, so nothing changes in the output
    for max_dst, refill, bin in ((-1.0, 24, 0), (30.0, 1, 0), (-1.0, 20, 7 | 16), (30.0, 24, 8), (-1.0, 1, 2)):
        tasks = bench.picker_tasks(pkg, world, 5, 3000, seed=3, max_dst=max_dst)
        want, cnt = scene.raycast(tasks)
        got = np.zeros(len(tasks), dtype=pkg.RESULT_DTYPE)
        c = (C.c_uint64 * 6)()
        assert emu.emu_raycast(*args, C.c_void_p(tasks.ctypes.data), C.c_uint64(len(tasks)), C.c_void_p(got.ctypes.data), opts(refill=refill, rgba8=bin), c) == 0
        assert got.tobytes() == want.tobytes(), (fmt, max_dst)
        assert (int(c[2]), int(c[3]), int(c[4])) == (cnt["steps"], cnt["pushes"], cnt["leaf_tests"])
        assert 0 < int((got["dst"] > 0).sum()) < len(tasks)


def test_emulated_ray_binning_is_a_sorted_permutation(emu, pkg):
    """bin_count / bin_scan_* / bin_scatter kernels: order[] is a permutation of the task indices, keys are non-decreasing along it, and the
    key is the Z-order code of the origin cell (+ octant) — also for origins outside the octree, NaN origins and a ragged batch."""
    rng = np.random.default_rng(5)
    n = 10_000
    tasks = np.zeros(n, dtype=pkg.TASK_DTYPE)
    tasks["pos"] = rng.uniform(-40.0, 300.0, (n, 3)).astype(np.float32)      # octree of 256^3 voxels (scale 2^-8): some origins lie outside
    tasks["pos"][:7] = np.nan
    tasks["dir"] = rng.normal(size=(n, 3)).astype(np.float32)
    emu.emu_bin_order.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]
    emu.emu_bin_order.restype = C.c_int
    for bits, octant in ((7, 1), (8, 0), (3, 1), (1, 0)):
        order, keys = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        assert emu.emu_bin_order(tasks.ctypes.data, n, 2.0 ** -8, bits | (octant << 4), order.ctypes.data, keys.ctypes.data) == 0
        assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
        assert (np.diff(keys[order].astype(np.int64)) >= 0).all()
        cells = 1 << bits
        with np.errstate(invalid="ignore"):
            c = np.clip(np.nan_to_num(tasks["pos"].astype(np.float32) * np.float32(2.0 ** -8 * cells), nan=0.0), 0, cells - 1).astype(np.uint32)
        want = np.zeros(n, np.uint32)
        for b in range(bits):
            for ax in range(3):
                want |= ((c[:, ax] >> b) & 1) << (3 * b + ax)
        if octant:
            d = tasks["dir"]
            want = (want << 3) | (d[:, 0] > 0) | ((d[:, 1] > 0).astype(np.uint32) << 1) | ((d[:, 2] > 0).astype(np.uint32) << 2)
        assert np.array_equal(keys, want), (bits, octant)


def emu_raycast(emu, pkg, world, reg, tasks, **kw):
    keep, args = scene_args(world, reg)
    got = np.zeros(max(len(tasks), 1), dtype=pkg.RESULT_DTYPE)
    c = (C.c_uint64 * 6)()
    assert emu.emu_raycast(*args, C.c_void_p(tasks.ctypes.data if len(tasks) else got.ctypes.data), C.c_uint64(len(tasks)), C.c_void_p(got.ctypes.data),
                           opts(**kw), c) == 0
    return got[:len(tasks)], {"steps": int(c[2]), "pushes": int(c[3]), "leaf_tests": int(c[4])}


@pytest.mark.parametrize("fmt", [0, 1])
def test_emulated_translucent_world(emu, pkg, ora, fmt):
    """The bundled Minecraft world fixture (water, leaves: texels with alpha 0): primary and shadow rays pass through translucent
    leaves by the rule of svo.esvo.glsl:241-242 / 264-265 — render_leaf, translucent_leaf_accepts and walk_skip_leaf in both kernels."""
    reg = pkg.content_registry(pkg.load_atlas())
    world = helpers.mc_world(pkg, fmt)
    w, h = 96, 54
    p = helpers.mc_params(pkg, w, h, shadows=True)
    q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
    q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
    vxp = pkg.to_vx_render_params(q)
    want, want8, cnt = oracle_render(pkg, ora, world, reg, vxp, w, h)
    assert cnt["leaf_tests"] > 1.3 * (cnt["primary_rays"] + cnt["shadow_rays"]) * 0.5      # leaves are tested and passed through
    for refill in (1, 12):
        got, _, c = emu_render(emu, pkg, world, reg, vxp, w, h, refill=refill)
        assert got.tobytes() == want.tobytes() and c == cnt, (fmt, refill, c, cnt)


def test_emulated_edge_cases(emu, pkg, ora):
    """The degenerate inputs of the GPU edge-case test, on the emulated kernels: no rays, single pathological rays, 1-pixel frames, and
    the MAX_STEPS budget (svo.esvo.glsl:152,392) running out in the middle of a batch."""
    reg = helpers.shader_test_registry(pkg)
    world = helpers.shader_test_world(pkg, [(x, 0, z, 1) for x in range(32) for z in range(32)] + [(5, 5, 5, 2), (31, 31, 31, 1)])
    scene = helpers.oracle_scene(ora, world, reg)
    got, _ = emu_raycast(emu, pkg, world, reg, np.zeros(0, dtype=pkg.TASK_DTYPE))
    assert len(got) == 0
    cases = [((16.0, 40.0, 16.0), (0.0, -1.0, 0.0), -1.0), ((-50.0, 10.0, 16.0), (-1.0, 0.0, 0.0), -1.0), ((16.0, 1.5, 16.0), (1.0, 0.0, 0.0), -1.0),
             ((5.5, 5.5, 5.5), (0.0, 1.0, 0.0), -1.0), ((16.0, 40.0, 16.0), (0.0, -1.0, 0.0), 0.0), ((1e6, 1e6, 1e6), (-1.0, -1.0, -1.0), -1.0),
             ((16.0, float("nan"), 16.0), (0.0, -1.0, 0.0), -1.0)]
    t = np.zeros(len(cases), dtype=pkg.TASK_DTYPE)
    for i, (pos, d, md) in enumerate(cases):
        t["pos"][i], t["dir"][i], t["max_dst"][i] = pos, np.array(d, np.float32) / np.linalg.norm(d), md
    want, _ = scene.raycast(t)
    got, _ = emu_raycast(emu, pkg, world, reg, t)
    assert got.tobytes() == want.tobytes()          # same libm, same NaN: byte-identical even for the NaN origin
    for (fw, fh) in ((1, 1), (1, 40), (40, 1), (33, 17)):
        vxp = pkg.to_vx_render_params(pkg.render_params(cam_pos=(16.0, 20.0, 50.0), cam_fwd=(0.0, -0.4, -1.0), fov_y_deg=72.0, aspect=fw / fh))
        want, _, cnt = oracle_render(pkg, ora, world, reg, vxp, fw, fh)
        got, _, c = emu_render(emu, pkg, world, reg, vxp, fw, fh)
        assert got.tobytes() == want.tobytes() and c == cnt, (fw, fh)
    # budget: a row of "dust" chunks makes a skimming ray descend to voxel level every other cell
    w2 = pkg.World()
    dust = [(x, 0, z, 1) for x in range(0, 32, 2) for z in range(0, 32, 2)]
    for cx in range(32):
        w2.set_leaf_blocks((cx, 0, 0), list(dust) + ([(31, y, z, 2) for y in range(4) for z in range(4)] if cx == 31 else []), uid=100 + cx, lod=5, compact=True)
    w2.serialize()
    n = 256
    tasks = np.zeros(n, dtype=pkg.TASK_DTYPE)
    tasks["max_dst"] = -1.0
    tasks["pos"] = np.stack([np.linspace(0.5, 1000.5, n), np.full(n, 0.5), 1.5 + 2.0 * (np.arange(n) % 16)], axis=1)
    tasks["dir"] = (1.0, 0.0, 0.0)
    want, cnt = helpers.oracle_scene(ora, w2, reg).raycast(tasks)
    got, c = emu_raycast(emu, pkg, w2, reg, tasks)
    assert got.tobytes() == want.tobytes() and c["steps"] == cnt["steps"] and cnt["steps"] > 500 * n
    assert 0 < (got["dst"] > 0).sum() < n


def test_emulated_chunk_serialization(emu, pkg):
    """serialize_chunks_kernel on the emulator against the host serializer: random chunks of every density at every LOD,
    the reference's KAT chunk, and the chunks of a small generated world against the bytes in the host's RangeBuffer."""
    rng = np.random.default_rng(21)
    blocks, lods = [], []
    for density in (0.0, 0.0005, 0.02, 0.3, 1.0):
        for lod in (0, 1, 2, 3, 4, 5):
            blocks.append(((rng.random(32768) < density) * rng.integers(1, 1 << 20, 32768)).astype(np.uint32))
            lods.append(lod)
    corner = np.zeros(32768, np.uint32); corner[31] = 1; corner[31 * 32] = 2; corner[31 * 1024] = 3
    blocks.append(corner); lods.append(5)

    def serialize(blocks, lods):
        blocks = np.ascontiguousarray(np.stack(blocks), dtype=np.uint32)
        lods_a = np.ascontiguousarray(lods, dtype=np.uint8)
        infos = np.zeros(len(blocks), dtype=pkg.CHUNK_INFO_DTYPE)
        rec = np.zeros(len(blocks) * 4681 * 48, dtype=np.uint8)
        total = C.c_uint64()
        assert emu.emu_serialize_chunks(C.c_void_p(blocks.ctypes.data), C.c_uint32(len(blocks)), C.c_void_p(lods_a.ctypes.data), C.c_void_p(infos.ctypes.data),
                                        C.c_void_p(rec.ctypes.data), C.c_uint64(len(rec)), C.byref(total)) == 0
        return infos, rec[:total.value]

    infos, rec = serialize(blocks, lods)
    assert sum(int(l) for l in infos["length_bytes"]) == len(rec)
    out = np.zeros(4681 * 12, dtype=np.uint32)
    res = (C.c_uint8 * 3)()
    for i in range(len(blocks)):
        n = pkg.host().vxh_serialize_dense(blocks[i].ctypes.data, lods[i], out.ctypes.data, len(out), res)
        o, l = int(infos["offset_bytes"][i]), int(infos["length_bytes"][i])
        assert l == n * 4 and rec[o:o + l].tobytes() == out[:n].tobytes(), (i, lods[i])
        assert (infos["child_mask"][i], infos["leaf_mask"][i], infos["depth"][i]) == tuple(res), (i, lods[i])
    world = pkg.World(radius=7, center=(-1, 2, 5), seed=1, terrain="reference")     # LOD 5 and LOD 4 chunks
    world.generate(0, 8)
    world.serialize()
    image, chunks = world.range_bytes(), world.chunks()[::3]
    cb = [world.chunk_blocks(c) for c in chunks]
    assert len({l for _, l in cb}) > 1
    infos, rec = serialize([b for b, _ in cb], [l for _, l in cb])
    for c, info in zip(chunks, infos):
        off, length = world.chunk_range(c)
        assert int(info["length_bytes"]) == length
        assert rec[int(info["offset_bytes"]):int(info["offset_bytes"]) + length].tobytes() == image[off:off + length].tobytes(), tuple(c)


def test_emulated_dirty_scatter_and_shard_copy(emu, pkg):
    """scatter_ranges_kernel applies a packed dirty set exactly like the host reference of voxelrs_b200.sharded; shard_copy_kernel packs the
    macro blocks of a rank and puts them back; rgba8_kernel is Framebuffer::read_pixels' rounding."""
    from importlib import import_module
    sharded = import_module(pkg.__name__ + ".sharded")
    rng = np.random.default_rng(5)
    # ESVO: 24-byte head, word-aligned ranges; CSVO: 8-byte head, ranges at arbitrary byte offsets and of arbitrary length (tails of
    # 1-3 bytes, a 1-byte range, two ranges sharing a word)
    for head, ranges in ((24, [(0, 48), (96, 480), (1024, 4), (4000, 96)]), (8, [(0, 13), (13, 1), (14, 5), (101, 479), (1023, 6), (3001, 1003)])):
        mirror = rng.integers(0, 256, 4096 + head, dtype=np.uint8)
        packed = np.ascontiguousarray(sharded.pack_dirty_host(mirror, ranges, head=head))
        replica = np.zeros_like(mirror)
        want = replica.copy()
        assert sharded.apply_packed_host(want, packed, len(ranges), head=head) == (len(packed), 0)
        payload = len(packed) - 16 * len(ranges)
        errors = C.c_uint(0)
        emu.emu_scatter_ranges(C.c_void_p(replica.ctypes.data), C.c_void_p(packed.ctypes.data), C.c_uint32(len(ranges)), C.c_uint64(payload),
                               C.c_uint32(head), C.c_uint32(3), C.c_uint64(len(replica)), C.byref(errors))
        assert replica.tobytes() == want.tobytes() and errors.value == 0
        for off, ln in ranges:
            assert replica[head + off:head + off + ln].tobytes() == mirror[head + off:head + off + ln].tobytes()
    # a header that points outside the buffer: that range is skipped and counted, the others are applied, nothing is written out of bounds
    head, ranges = 24, [(0, 48), (4090, 64), (200, 8)]
    mirror = rng.integers(0, 256, 8192, dtype=np.uint8)
    packed = np.ascontiguousarray(sharded.pack_dirty_host(mirror, ranges, head=head))
    guard = np.zeros(4096 + head + 256, np.uint8)
    replica = guard[:4096 + head]
    want = replica.copy()
    assert sharded.apply_packed_host(want, packed, len(ranges), head=head)[1] == 1
    errors = C.c_uint(0)
    emu.emu_scatter_ranges(C.c_void_p(replica.ctypes.data), C.c_void_p(packed.ctypes.data), C.c_uint32(len(ranges)), C.c_uint64(len(packed) - 48),
                           C.c_uint32(head), C.c_uint32(2), C.c_uint64(len(replica)), C.byref(errors))
    assert replica.tobytes() == want.tobytes() and errors.value == 1 and not guard[4096 + head:].any()
    w, h, size = 70, 40, 3
    frame = rng.random((h, w, 4), dtype=np.float32)
    rebuilt = np.zeros_like(frame)
    for rank in range(size):
        n_macros = ((w + 31) // 32) * ((h + 15) // 16)
        owned = (n_macros - rank + size - 1) // size
        buf = np.zeros((owned, 512, 4), np.float32)
        assert emu.emu_shard_copy(C.c_void_p(frame.ctypes.data), C.c_void_p(buf.ctypes.data), C.c_uint32(w), C.c_uint32(h), C.c_uint32(rank), C.c_uint32(size), 1) == owned
        assert emu.emu_shard_copy(C.c_void_p(rebuilt.ctypes.data), C.c_void_p(buf.ctypes.data), C.c_uint32(w), C.c_uint32(h), C.c_uint32(rank), C.c_uint32(size), 0) == owned
    assert rebuilt.tobytes() == frame.tobytes()
    f = np.array([[0.0, 1.0, 0.5, 2.0], [-1.0, 0.25, 1e-9, np.nan], [0.499 / 255, 0.5 / 255, 254.5 / 255, 1.0]], np.float32)
    out = np.zeros(3, np.uint32)
    emu.emu_rgba8(C.c_void_p(f.ctypes.data), C.c_void_p(out.ctypes.data), C.c_uint64(3))
    want = np.floor(np.clip(np.nan_to_num(f, nan=0.0), 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.uint32)
    assert [int(v) for v in out] == [int(r[0] | r[1] << 8 | r[2] << 16 | r[3] << 24) for r in want]


def test_emulated_random_configurations(emu, pkg, ora, terrains):
    """Seeded sweep over what the host can ask of the frame kernels: ragged frame sizes, shard counts, bands, refill thresholds, numbers of
    persistent CTAs, both formats, shadows on/off, float and RGBA8 output. The union of the shards always equals the oracle's frame."""
    worlds, reg = terrains
    rng = np.random.default_rng(2024)
    for case in range(14):
        fmt = int(rng.integers(0, 2))
        world = worlds[fmt]
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 60))
        size = int(rng.choice([1, 1, 2, 3, 5]))
        rows = int(rng.integers(0, 2)) << 31       # VX_SHARD_ROWS
        bands = int(rng.choice([1, 2, 4, 16]))
        refill, ctas = int(rng.integers(1, 33)), int(rng.integers(1, 5))
        shadow_refill = int(rng.choice([0, 1, 7, 32]))
        shadows, rgba8 = bool(rng.integers(0, 2)), int(rng.integers(0, 2))
        overlap = int(rng.integers(0, 2))
        lifo = case % 3 != 0
        vxp = world_params(pkg, world, w, h, shadows=shadows, selected=(-20.0, 50.0, 174.0) if case % 2 else None)
        want, want8, cnt = oracle_render(pkg, ora, world, reg, vxp, w, h)
        union = np.full((h, w, 4), -1.0, np.float32)
        union8 = np.zeros((h, w, 4), np.uint8)
        total = {k: 0 for k in cnt}
        for rank in range(size):
            got, got8, c = emu_render(emu, pkg, world, reg, vxp, w, h, refill=refill, shadow_refill=shadow_refill, ctas=ctas, rgba8=rgba8, rank=rank,
                                      size=size | rows, bands=bands, overlap=overlap, lifo=int(lifo))
            mine = (got8.view(np.uint32)[..., 0] != 0xdeadbeef) if rgba8 else (got[..., 3] != -1.0)
            union[mine] = got[mine]
            union8[mine] = got8[mine]
            for k in c:
                total[k] += c[k]
        cfg = dict(case=case, fmt=fmt, w=w, h=h, size=size, rows=bool(rows), bands=bands, refill=refill, ctas=ctas, shadows=shadows, rgba8=rgba8)
        assert (union8.tobytes() == want8.tobytes()) if rgba8 else (union.tobytes() == want.tobytes()), cfg
        assert total == cnt, (cfg, total, cnt)
