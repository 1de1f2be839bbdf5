"""BASELINE configs[0]: the "fixed test world from assets/worlds" — the part of the reference's bundled Minecraft world that
is present in the checkout, decoded by voxel-rs_b200/anvil.py (the reference's `--mc-world` path, src/systems/storage.rs) and
committed as tests/golden/mc_world.npz by tests/golden/make_mc_fixture.py. Water, leaves and glass-like translucent texels make
it the stress case for the translucency rule (svo.esvo.glsl:241-242) that the generated terrain never exercises.
CPU part here; the 1280x720 frame and the ray batches on the GPU are in tests/test_gpu_parity.py."""
import ctypes as C
import importlib.util
import os

import numpy as np

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _anvil():
    spec = importlib.util.spec_from_file_location("vx_anvil", os.path.join(ROOT, "voxel-rs_b200", "anvil.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_block_mapping():
    """storage.rs:126-151"""
    a = _anvil()
    assert a.block_id("minecraft:iron_ore") == a.AIR and a.block_id("minecraft:deepslate_gold_ore") == a.AIR
    assert a.block_id("minecraft:birch_leaves") == a.OAK_LEAVES and a.block_id("minecraft:spruce_log") == a.OAK_LOG
    assert a.block_id("minecraft:jungle_planks") == a.OAK_PLANKS
    for n, v in (("minecraft:air", a.AIR), ("minecraft:kelp", a.AIR), ("minecraft:dirt", a.DIRT), ("minecraft:grass_block", a.GRASS),
                 ("minecraft:clay", a.GRAVEL), ("minecraft:sandstone", a.SAND), ("minecraft:water", a.WATER), ("minecraft:tuff", a.STONE),
                 ("minecraft:cobblestone", a.COBBLESTONE), ("minecraft:obsidian", a.AIR)):
        assert a.block_id(n) == v, n


def test_nbt_and_section_unpacking():
    """A hand-made region file with one chunk: NBT parsing, zlib payload, 4-bit and 5-bit palette indices that do not span
    64-bit words, the [y][z][x] section order and the 2 x 2 column assembly of an engine chunk."""
    import struct
    import zlib
    a = _anvil()

    def s(x): b = x.encode(); return struct.pack(">H", len(b)) + b
    def tag(t, name, payload): return bytes([t]) + s(name) + payload
    def compound(items): return b"".join(items) + b"\x00"
    def palette(names): return bytes([10]) + struct.pack(">i", len(names)) + b"".join(compound([tag(8, "Name", s(n))]) for n in names)
    def longs(idx, bits):
        per = 64 // bits
        words = []
        for i in range(0, 4096, per):
            w = 0
            for k, v in enumerate(idx[i:i + per]):
                w |= int(v) << (bits * k)
            words.append(w)
        return struct.pack(">i", len(words)) + b"".join(struct.pack(">Q", w) for w in words)

    rng = np.random.default_rng(4)
    names4 = ["minecraft:air", "minecraft:stone", "minecraft:water"]
    names5 = ["minecraft:air"] + ["minecraft:stone", "minecraft:dirt", "minecraft:sand", "minecraft:oak_log", "minecraft:oak_leaves"] * 4   # 21 -> 5 bits
    idx4, idx5 = rng.integers(0, len(names4), 4096), rng.integers(0, len(names5), 4096)
    sec = lambda y, names, idx, bits: compound([tag(1, "Y", struct.pack(">b", y)),
                                                tag(10, "block_states", compound([tag(9, "palette", palette(names)), tag(12, "data", longs(idx, bits))]))])
    chunk = bytes([10]) + s("") + compound([tag(9, "sections", bytes([10]) + struct.pack(">i", 2) + sec(2, names4, idx4, 4) + sec(3, names5, idx5, 5))])
    comp = zlib.compress(chunk)
    region = bytearray(8192 + 4096 * ((len(comp) + 5 + 4095) // 4096))
    slot = 3 + 32 * 5                                     # chunk (3, 5) of the region
    region[4 * slot:4 * slot + 4] = (2).to_bytes(3, "big") + bytes([1])
    region[8192:8192 + 5 + len(comp)] = struct.pack(">IB", len(comp) + 1, 2) + comp
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "r.-1.2.mca"), "wb").write(region)
        w = a.MinecraftWorld(d)
    assert list(w.chunks) == [(-32 + 3, 64 + 5)]
    jc = w.chunks[(-29, 69)]
    want4 = np.array([a.block_id(n) for n in names4], np.uint8)[idx4].reshape(16, 16, 16)
    want5 = np.array([a.block_id(n) for n in names5], np.uint8)[idx5].reshape(16, 16, 16)
    assert (jc.sections[2] == want4).all() and (jc.sections[3] == want5).all()
    # engine chunk (cx, 1, cz) covers world heights 32..63 = sections 2 and 3; Minecraft column (-29, 69) is its (dx, dz) = (1, 1) quarter
    b = w.engine_chunk(-15, 1, 34).reshape(32, 32, 32)    # [z][y][x]
    assert (b[16:, :16, 16:] == want4.transpose(1, 0, 2)).all() and (b[16:, 16:, 16:] == want5.transpose(1, 0, 2)).all()
    assert not b[:16].any() and not b[:, :, :16].any()


def test_fixture_world(pkg, ora):
    """The committed fixture builds the same voxels into both SVO formats; the oracle sees water, sand, grass, logs and leaves
    from the fixed camera, rejects translucent texels (more leaf tests than hits) and renders the same frame through the ESVO
    and the CSVO shader (up to shadow rays that start inside a voxel)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "mc_world.npz"))
    assert z["blocks"].shape == (166, 32768) and set(np.unique(z["blocks"]).tolist()) >= {0, 1, 2, 3, 7, 8, 9, 10}
    reg = pkg.content_registry(pkg.load_atlas())
    frames = {}
    for fmt in (0, 1):
        w = helpers.mc_world(pkg, fmt)
        assert w.depth == 4 + 5
        s = helpers.oracle_scene(ora, w, reg)
        p = helpers.mc_params(pkg, 160, 90, shadows=True)
        q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
        q.cam_pos = (C.c_float * 3)(*w.cnv_block_pos(tuple(p.cam_pos)))
        vxp = pkg.to_vx_render_params(q)
        img, cnt = s.render(vxp, 160, 90)
        hits = s.primary_hits(vxp, 160, 90)
        seen = set(np.unique(hits["value"][hits["t"] >= 0]).tolist())
        assert seen >= {1, 7, 8, 9, 10}, seen
        assert cnt["leaf_tests"] > 1.5 * (hits["t"] >= 0).sum()          # translucent leaves / water texels are passed through
        frames[fmt] = img
    assert (np.abs(frames[0] - frames[1]).max(axis=2) > 0).mean() < 2e-3
