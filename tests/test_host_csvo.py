"""Known-answer tests of the host CSVO serializer against the reference's own serializer tests
(src/world/hds/csvo.rs:330-391 Csvo::serialize, :592-712 SerializedChunk::serialize_octant_*). CPU only."""
import ctypes as C

import numpy as np


def octant(pkg, blocks, expand_to, depth_minus=0, compact=True):
    arr = np.array(blocks, dtype=np.uint32).reshape(-1, 4)
    out = np.zeros(1 << 16, dtype=np.uint8)
    mats = np.zeros(1 << 12, dtype=np.uint32)
    n_m = C.c_uint32()
    n = pkg.host().vxh_kat_csvo_octant(arr.ctypes.data, len(arr), expand_to, int(compact), depth_minus, out.ctypes.data, len(out),
                                       mats.ctypes.data, len(mats), C.byref(n_m))
    return out[:n].tolist(), mats[:n_m.value].tolist()


BLOCKS_3 = [(31, 0, 0, 1), (0, 31, 0, 2), (0, 0, 31, 3)]
CHUNK_NODES = [
    0b00010100, 0b00000001, 0, 9, 18,
    0b00000100, 0, 0,
    2, 0,
    2, 0, 0, 2,
    0b00010000, 0, 0,
    4, 0,
    4, 1, 0, 4,
    0, 0b00000001, 0,
    16, 0,
    16, 2, 0, 16,
]


def test_serialize_octant_single_leaf(pkg):
    """csvo.rs:592-607"""
    nodes, mats = octant(pkg, [(0, 0, 0, 1)], expand_to=4)
    assert nodes == [1, 0, 0,  1, 0,  1, 0, 0, 1]
    assert mats == [1]


def test_serialize_octant_multiple_leaves(pkg):
    """csvo.rs:609-629"""
    nodes, mats = octant(pkg, [(0, 0, 0, 1), (3, 3, 3, 2), (5, 4, 4, 1), (6, 7, 7, 2)], expand_to=4)
    assert nodes == [1, 0, 0,
                     1 | (1 << 7), 0, 5,
                     1 | (1 << 7), 0, 0, 1, 1 << 7,
                     1 | (1 << 7), 2, 0, 2, 1 << 6]
    assert mats == [1, 2, 1, 2]


def test_serialize_octant_chunk(pkg):
    """csvo.rs:631-656"""
    nodes, mats = octant(pkg, BLOCKS_3, expand_to=5)
    assert nodes == CHUNK_NODES
    assert mats == [1, 2, 3]


def test_serialize_octant_chunk_with_lod(pkg):
    """csvo.rs:658-711: depth - 1 .. depth - 4"""
    want = {
        1: [0b00010100, 0b00000001, 0, 6, 12,  2, 0,  2, 0, 0, 2,  4, 0,  4, 1, 0, 4,  16, 0,  16, 2, 0, 16],
        2: [0b00010110, 0, 4, 8,  2, 0, 0, 2,  4, 1, 0, 4,  16, 2, 0, 16],
        3: [0b00010110, 0, 0, 2, 4, 16],
        4: [22],
    }
    for minus, nodes_want in want.items():
        nodes, mats = octant(pkg, BLOCKS_3, expand_to=5, depth_minus=minus)
        assert nodes == nodes_want, minus
        assert mats == [1, 2, 3], minus


def test_csvo_serialize_world(pkg):
    """csvo.rs:330-391: one chunk at Position(1, 0, 0): chunk record [lod][material bytes][materials][nodes], then the root
    octant with an absolute 32-bit pointer; write_to = [root offset][bytes]."""
    w = pkg.World(fmt=pkg.FORMAT_CSVO)
    w.set_leaf_blocks((1, 0, 0), BLOCKS_3, uid=2435999049025295583, lod=5, compact=True)
    w.serialize()
    expected = [5,  12, 0, 0, 0,  1, 0, 0, 0,  2, 0, 0, 0,  3, 0, 0, 0] + CHUNK_NODES + [0b00001100, 0,  0, 0, 0, 1 << 7]
    assert pkg.host().vxh_world_csvo_root_offset(w.h) == 49
    out = np.zeros(256, np.uint8)
    n = pkg.host().vxh_world_range_bytes(w.h, out.ctypes.data, len(out))
    assert out[:n].tolist() == expected
    assert w.dirty_ranges() == [(0, 55)]
    assert w.depth == 1 + 5 and w.size_bytes == 55 and w.header_bytes == 8
    buf = w.gpu_buffer()
    assert buf[4:8].tolist() == [49, 0, 0, 0] and buf[8:8 + 55].tolist() == expected
    assert np.frombuffer(buf[:4].tobytes(), np.float32)[0] == np.float32(2.0 ** -6)


def test_csvo_dense_fast_path_equals_generic(pkg):
    rng = np.random.default_rng(11)
    for density, lod in ((0.02, 5), (0.3, 5), (0.3, 4), (0.1, 3), (0.5, 2), (0.2, 1), (0.05, 0)):
        blocks = ((rng.random(32 ** 3) < density) * rng.integers(1, 13, 32 ** 3)).astype(np.uint32)
        assert pkg.host().vxh_csvo_dense_equals_generic(blocks.ctypes.data, lod) == 1, (density, lod)


def test_csvo_world_is_smaller_than_esvo(pkg):
    """The point of the format: byte-packed nodes instead of 48-byte records. (Measured here: 1.7x smaller,
    not more, because every voxel still costs a 4-byte BlockId in the chunk's material list.)"""
    sizes = {}
    for fmt in (pkg.FORMAT_ESVO, pkg.FORMAT_CSVO):
        w = pkg.World(radius=3, center=(-1, 2, 5), seed=1, fmt=fmt)
        w.generate(0, 8)
        w.serialize()
        sizes[fmt] = w.size_bytes
        assert w.depth == 3 + 5
    assert sizes[pkg.FORMAT_CSVO] * 3 < sizes[pkg.FORMAT_ESVO] * 2
