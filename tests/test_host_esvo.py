"""Known-answer tests of the host ESVO builder against the reference's serializer tests
(src/world/hds/esvo.rs:561-1228, src/world/hds/internal.rs:279-455, src/systems/worldsvo.rs:514-557,
src/systems/chunkloader.rs:127-134). CPU only."""
import ctypes as C

import numpy as np

REL = 1 << 31
BLOCKS_3 = [(31, 0, 0, 1), (0, 31, 0, 2), (0, 0, 31, 3)]   # esvo.rs:564-567, 863-866


def kat(pkg, blocks, lod, expand_to=5, compact=True):
    arr = np.array(blocks, dtype=np.uint32).reshape(-1, 4)
    out = np.zeros(1 << 16, dtype=np.uint32)
    res = (C.c_uint8 * 3)()
    n = pkg.host().vxh_kat_block_octree(arr.ctypes.data, len(arr), expand_to, int(compact), lod, out.ctypes.data, len(out), res)
    return out[:n].tolist(), tuple(res)


def subtree(idx, value, levels):
    """`levels` inner records descending through child `idx`, then the leaf record (esvo.rs:607-707 pattern)."""
    words = []
    hdr_word, shift = idx // 2, 16 * (idx % 2)
    for l in range(levels):
        hdr = [0, 0, 0, 0]
        mask = (1 << idx) << 8
        if l == levels - 1:
            mask |= 1 << idx     # the next record holds voxels: leaf bit set
        hdr[hdr_word] = mask << shift
        body = [0] * 8
        body[idx] = REL | (12 - 4 - idx)
        words += hdr + body
    leaf = [0] * 12
    leaf[4 + idx] = value
    return words + leaf


def expected_lod(lod):
    """Expected buffers of serialize_with_lod (esvo.rs:861-1228) built from its pattern."""
    depth = lod if lod else 5
    if depth == 1:
        return [0, 0, 0, 0, 0, 1, 2, 0, 3, 0, 0, 0], (2 | 4 | 16, 2 | 4 | 16, 1)
    inner = depth - 2                     # inner records per subtree below the core octant
    sub_len = (inner + 1) * 12
    subs = [subtree(1, 1, inner), subtree(2, 2, inner), subtree(4, 3, inner)]
    leafbit = lambda i: (1 << i) if inner == 0 else 0
    hdr = [(((2 << 8) | leafbit(1)) << 16), (4 << 8) | leafbit(2), (16 << 8) | leafbit(4), 0]
    body = [0, REL | 7, REL | (6 + sub_len), 0, REL | (4 + 2 * sub_len), 0, 0, 0]
    return hdr + body + subs[0] + subs[1] + subs[2], (2 | 4 | 16, 0, depth)


def test_serialize_with_lod(pkg):
    """esvo.rs:861-1228: LOD 5..1 of a chunk with three corner voxels."""
    # spot-check the generator against literal rows of the reference test first (esvo.rs:873-881, 1164-1172, 1213-1222)
    full, res = expected_lod(5)
    assert full[:12] == [(2 << 8) << 16, 4 << 8, 16 << 8, 0, 0, REL | 7, REL | (6 + 4 * 12), 0, REL | (4 + 8 * 12), 0, 0, 0]
    assert expected_lod(2)[0][:12] == [((2 << 8) | 2) << 16, 4 << 8 | 4, 16 << 8 | 16, 0, 0, REL | 7, REL | (6 + 12), 0, REL | (4 + 2 * 12), 0, 0, 0]
    for lod in (5, 4, 3, 2, 1):
        got, gres = kat(pkg, BLOCKS_3, lod)
        want, wres = expected_lod(lod)
        assert got == want, lod
        assert gres == wres, lod
    got0, _ = kat(pkg, BLOCKS_3, 0)
    assert got0 == expected_lod(5)[0]


def test_esvo_serialize_world(pkg):
    """esvo.rs:562-742: one SerializedChunk at world position (1,0,0): root_info, RangeBuffer ranges, write_to."""
    w = pkg.World()
    w.set_leaf_blocks((1, 0, 0), BLOCKS_3, uid=100, lod=0, compact=True)
    w.serialize()
    assert w.root_info() == (156, 2, 0, 6)            # esvo.rs:587-594
    chunk, _ = expected_lod(5)
    assert len(chunk) == 156
    root = [(2 | 4 | 16) << 8 << 16, 0, 0, 0, 0, 5, 0, 0, 0, 0, 0, 0]   # esvo.rs:709-716
    assert w.dirty_ranges() == [(0, 672)]             # esvo.rs:721
    assert w.root_range() == (624, 48)                # esvo.rs:724
    buf = w.gpu_buffer()
    assert np.frombuffer(buf[:4].tobytes(), np.float32)[0] == 2.0 ** -6
    words = np.frombuffer(buf[4:].tobytes(), dtype=np.uint32).tolist()
    assert words == [2 << 8, 0, 0, 0, 156 + 5] + chunk + root            # esvo.rs:731-741
    # every octant record starts at GPU-buffer byte 24 + 48k -> 16-byte vector loads are legal
    assert w.size_bytes % 48 == 0


def esvo32(pkg):
    H = pkg.host()

    class E:
        def __init__(s):
            s.h = H.vxh_esvo32_new()

        def set_leaf(s, pos, v, ser=True):
            o = (C.c_uint32 * 2)()
            H.vxh_esvo32_set_leaf(s.h, *pos, v, int(ser), o)
            return tuple(o)

        def move_leaf(s, leaf, pos):
            o, old = (C.c_uint32 * 2)(), C.c_uint32()
            had = H.vxh_esvo32_move_leaf(s.h, leaf[0], leaf[1], *pos, o, C.byref(old))
            return tuple(o), (old.value if had else None)

        def remove_leaf(s, leaf):
            old = C.c_uint32()
            return old.value if H.vxh_esvo32_remove_leaf(s.h, leaf[0], leaf[1], C.byref(old)) else None

        def serialize(s):
            H.vxh_esvo32_serialize(s.h)

        def root_info(s):
            off, m = C.c_uint64(), (C.c_uint8 * 3)()
            H.vxh_esvo32_root_info(s.h, C.byref(off), m)
            return off.value, m[0], m[1], m[2]

        def words(s):
            n = H.vxh_esvo32_bytes(s.h, None, 0)
            b = np.zeros(n, np.uint8)
            H.vxh_esvo32_bytes(s.h, b.ctypes.data, n)
            return np.frombuffer(b.tobytes(), np.uint32).tolist()

        def ranges(s, kind):
            arr = (pkg.VxRange * 16)()
            n = H.vxh_esvo32_ranges(s.h, kind, arr, 16)
            return [(arr[i].offset, arr[i].length) for i in range(n)]

        def range_of(s, uid):
            r = pkg.VxRange()
            return (r.offset, r.length) if H.vxh_esvo32_range_of(s.h, uid, C.byref(r)) else None

    return E()


def test_serialize_with_remove_and_move(pkg):
    """esvo.rs:745-858 with the reference's u32 leaf fake (worldsvo.rs:236-245)."""
    H = pkg.host()
    e = esvo32(pkg)
    e.set_leaf((0, 0, 0), 10); e.serialize()
    e.set_leaf((1, 0, 0), 20); e.serialize()
    assert e.root_info() == (1, 2 | 1, 0, 2)
    expected = [10, (((1 << 8) | 1) << 16) | ((1 << 8) | 1), 0, 0, 0, 5, 18, 0, 0, 0, 0, 0, 0, 20]
    assert e.words() == expected
    assert e.ranges(0) == [] and e.ranges(1) == [(0, 56)]
    assert e.range_of(10) == (0, 4) and e.range_of(20) == (52, 4) and e.range_of(2 ** 64 - 1) == (4, 48)
    H.vxh_esvo32_clear_updated(e.h)
    buf = np.zeros(800, np.uint8)
    size = H.vxh_esvo32_write_to(e.h, buf.ctypes.data)
    assert np.frombuffer(buf[:size].tobytes(), np.uint32).tolist() == [(2 | 1) << 8, 0, 0, 0, 1 + 5] + expected
    new_leaf, old = e.move_leaf((0, 1), (1, 1, 1))
    assert new_leaf == (0, 7) and old is None
    assert e.remove_leaf((0, 0)) == 10
    e.serialize()
    assert e.root_info() == (0, 1 << 7, 0, 2)
    expected2 = [0, 0, 0, ((1 << 8) | 1) << 16, 0, 0, 0, 0, 0, 0, 0, 18, 0, 20]
    assert e.words() == expected2
    assert e.ranges(0) == [(48, 4)] and e.ranges(1) == [(0, 48)]
    assert H.vxh_esvo32_write_changes_to(e.h, buf.ctypes.data, 800, 1) == 0
    # esvo.rs:849-855; its word 3, `(1 << 8) << 8 << 16`, shifts the bit out of a u32 and is 0
    got = np.frombuffer(buf[:size].tobytes(), np.uint32).tolist()
    assert got[:5] == [(1 << 7) << 8, 0, 0, 0, 5]
    assert got[5:] == expected2
    H.vxh_esvo32_free(e.h)


def test_range_buffer(pkg):
    """internal.rs:204-277 behaviour: first-fit reuse, adjacent free ranges merge, updated ranges merge."""
    H = pkg.host()
    r = H.vxh_rangebuf_new()
    data = lambda v, n: np.full(n, v, np.uint8)
    ins = lambda i, v, n: H.vxh_rangebuf_insert(r, i, data(v, n).ctypes.data, n)
    ranges = lambda k: [(a.offset, a.length) for a in list((lambda arr, n: arr[:n])(*(lambda arr: (arr, H.vxh_rangebuf_ranges(r, k, arr, 16)))((pkg.VxRange * 16)())))]
    assert ins(1, 1, 8) == 0 and ins(2, 2, 8) == 8 and ins(3, 3, 8) == 16
    assert ranges(1) == [(0, 24)] and ranges(0) == []
    H.vxh_rangebuf_remove(r, 1); H.vxh_rangebuf_remove(r, 2)
    assert ranges(0) == [(0, 16)]                    # merged
    assert ins(4, 4, 4) == 0 and ranges(0) == [(4, 12)]
    assert ins(5, 5, 12) == 4 and ranges(0) == []
    assert ins(6, 6, 4) == 24                        # appended
    assert ins(3, 7, 8) == 16                        # re-insert same id: freed then reused first-fit
    b = np.zeros(28, np.uint8)
    assert H.vxh_rangebuf_bytes(r, b.ctypes.data, 28) == 28
    assert b.tolist() == [4] * 4 + [5] * 12 + [7] * 8 + [6] * 4
    H.vxh_rangebuf_free(r)


def test_dense_fast_path_equals_generic(pkg):
    """serialize_dense_chunk == Chunk::fill_with (construct_octants_with) + serialize_octant for every LOD."""
    H = pkg.host()
    rng = np.random.default_rng(0)
    for fill in (0.0005, 0.02, 0.3, 1.0):
        blocks = ((rng.random(32 ** 3) < fill) * rng.integers(1, 13, 32 ** 3)).astype(np.uint32)
        for lod in (0, 5, 4, 3, 2, 1):
            a, b = np.zeros(40000 * 12, np.uint32), np.zeros(40000 * 12, np.uint32)
            ra, rb = (C.c_uint8 * 3)(), (C.c_uint8 * 3)()
            na = H.vxh_serialize_dense(blocks.ctypes.data, lod, a.ctypes.data, len(a), ra)
            nb = H.vxh_serialize_filled(blocks.ctypes.data, lod, b.ctypes.data, len(b), rb)
            assert na == nb and tuple(ra) == tuple(rb), (fill, lod)
            assert np.array_equal(a[:na], b[:nb]), (fill, lod)
    empty = np.zeros(32 ** 3, np.uint32)
    res = (C.c_uint8 * 3)()
    assert H.vxh_serialize_dense(empty.ctypes.data, 0, None, 0, res) == 0 and tuple(res) == (0, 0, 0)


def test_expand_chain_is_serialized(pkg):
    """a chunk placed at (15,15,15) in an empty world octree leaves 3 empty octants under child 0; they are
    serialized, which is what makes the golden pointer 11057 of svo_shader_tests.rs:735 come out."""
    w = pkg.World()
    w.set_leaf_blocks((15, 15, 15), [(0, 0, 0, 1)], compact=True)
    w.serialize()
    off, child_mask, leaf_mask, depth = w.root_info()
    assert child_mask == (1 << 0) | (1 << 7) and depth == 4 + 5
    start, length = w.root_range()
    assert length == 48 * (1 + 3 + 3)     # root + 3-octant empty chain + 3 octants down to the chunk


def test_svo_coord_space(pkg):
    """SvoCoordSpace tests, src/systems/worldsvo.rs:514-557."""
    # coord_space_positive (:514-524)
    w = pkg.World(radius=2, center=(4, 5, 12))
    world_pos = (32.0 * 5 + 16.25, 32.0 * 3 + 4.25, 32.0 * 10 + 20.5)
    svo_pos = w.cnv_block_pos(world_pos)
    assert tuple(svo_pos) == (32.0 * 3 + 16.25, 32.0 * 0 + 4.25, 32.0 * 0 + 20.5)
    assert tuple(w.cnv_svo_pos(svo_pos)) == world_pos
    # coord_space_negative (:527-537)
    w = pkg.World(radius=2, center=(-1, -1, -1))
    world_pos = (-16.25, -4.25, -20.5)
    svo_pos = w.cnv_block_pos(world_pos)
    assert tuple(svo_pos) == (32.0 * 2 + 15.75, 32.0 * 2 + 27.75, 32.0 * 2 + 11.5)
    assert tuple(w.cnv_svo_pos(svo_pos)) == world_pos
    # cnv_chunk_pos (:540-556)
    w = pkg.World(radius=1, center=(0, 0, 0))
    assert w.cnv_chunk_pos((-1, 0, 0)) == (0, 1, 1)
    assert w.cnv_chunk_pos((0, 0, 0)) == (1, 1, 1)
    assert w.cnv_chunk_pos((1, 0, 0)) == (2, 1, 1)
    assert w.cnv_chunk_pos((-2, 0, 0)) is None
    assert w.cnv_chunk_pos((2, 0, 0)) is None
    assert w.cnv_chunk_pos((1, 0, 1)) is None


def test_lod_rule(pkg):
    """ChunkLoader::calculate_lod, src/systems/chunkloader.rs:127-134."""
    lod = lambda dx, dz: pkg.host().vxh_calculate_lod(0, 0, 0, dx, 9, dz)
    assert [lod(d, 0) for d in (0, 6, 7, 12, 13, 19, 20, 40)] == [5, 5, 4, 4, 3, 3, 2, 2]
    assert lod(5, 5) == 4 and lod(4, 4) == 5


def test_generated_world_shape(pkg):
    """Named-shape world: disc of chunks, depth = world depth + 5, records 48-byte aligned, LOD shrinks far chunks."""
    w = pkg.World(radius=6, center=(-1, 2, 5), seed=1)
    n = w.generate(0, 8)
    w.serialize()
    assert n == w.chunk_count and n > 100
    assert w.depth == 4 + 5                      # 13^3 window -> world depth 4
    assert w.size_bytes % 48 == 0
    assert w.height_at(-24, 174) < 80            # the default camera is above ground
    w2 = pkg.World(radius=6, center=(-1, 2, 5), seed=1, no_lod=True)
    w2.generate(0, 8)
    w2.serialize()
    assert w2.size_bytes >= w.size_bytes
    # deterministic
    w3 = pkg.World(radius=6, center=(-1, 2, 5), seed=1)
    w3.generate(0, 8, threads=1)
    w3.serialize()
    assert w3.gpu_buffer().tobytes() == w.gpu_buffer().tobytes()
