"""The host octree and range buffer (voxel-rs_b200/host/octree.hpp, esvo.hpp) replayed against the reference's own unit tests:
src/world/hds/octree.rs:508-866 (exact octant tables after set / move / remove / construct / compact) and
src/world/hds/internal.rs:288-455 (RangeBuffer insert / remove states, merge_ranges cases). CPU only.

Octant notation of the expected states, like the reference's: (parent or None, [children], children_count) with a child written as
None, ("O", octant id) or ("L", leaf value)."""
import ctypes as C

import numpy as np

N = None


def O(i):
    return ("O", i)


def L(v):
    return ("L", v)


class Tree:
    def __init__(self, pkg):
        self.H = pkg.host()
        self.h = self.H.vxh_octree_new()

    def close(self):
        self.H.vxh_octree_free(self.h)

    def _out(self):
        return (C.c_int64 * 4)()

    def set_leaf(self, pos, v):
        o = self._out()
        self.H.vxh_octree_set_leaf(self.h, *pos, v, o)
        return (o[0], o[1]), (o[3] if o[2] else None)

    def move_leaf(self, leaf_id, pos):
        o = self._out()
        self.H.vxh_octree_move_leaf(self.h, leaf_id[0], leaf_id[1], *pos, o)
        return (o[0], o[1]), (o[3] if o[2] else None)

    def remove_leaf(self, pos):
        o = self._out()
        self.H.vxh_octree_remove_leaf(self.h, *pos, o)
        return (o[3] if o[2] else None), ((o[0], o[1]) if o[0] >= 0 else None)

    def remove_leaf_by_id(self, leaf_id):
        v = self.H.vxh_octree_remove_leaf_by_id(self.h, *leaf_id)
        return None if v < 0 else v

    def get_leaf(self, pos):
        v = self.H.vxh_octree_get_leaf(self.h, *pos)
        return None if v < 0 else v

    def compact(self):
        self.H.vxh_octree_compact(self.h)

    def construct(self, depth, cells):
        a = np.array([list(p) + [v] for p, v in cells], dtype=np.uint32).reshape(-1, 4)
        self.H.vxh_octree_construct(self.h, depth, a.ctypes.data if len(a) else None, len(a))

    def state(self):
        head = (C.c_int64 * 4)()
        octs = np.zeros((64, 18), dtype=np.int64)
        free = np.zeros(64, dtype=np.int64)
        self.H.vxh_octree_dump(self.h, head, octs.ctypes.data, 64, free.ctypes.data, 64)
        root, depth, n, nf = head[0], head[1], head[2], head[3]
        octants = []
        for r in octs[:n]:
            ch = [None if k == 0 else (("O", int(v)) if k == 1 else ("L", int(v))) for k, v in zip(r[2:10], r[10:18])]
            octants.append((None if r[0] < 0 else int(r[0]), ch, int(r[1])))
        return {"octants": octants, "free_list": [int(x) for x in free[:nf]], "root": None if root < 0 else int(root), "depth": int(depth)}


def test_octree_add_leaf_single(pkg):
    """octree.rs:515-545"""
    t = Tree(pkg)
    assert t.set_leaf((1, 1, 3), 20) == ((2, 7), None)
    assert t.state() == {
        "octants": [(1, [N] * 8, 0), (None, [O(0), N, N, N, O(2), N, N, N], 2), (1, [N, N, N, N, N, N, N, L(20)], 1)],
        "free_list": [], "root": 1, "depth": 2}
    assert t.get_leaf((1, 1, 3)) == 20 and t.get_leaf((1, 1, 1)) is None
    t.close()


def test_octree_add_leaf_multiple(pkg):
    """octree.rs:549-611"""
    t = Tree(pkg)
    assert t.set_leaf((6, 7, 5), 10) == ((4, 6), None)
    assert t.set_leaf((0, 0, 0), 20) == ((0, 0), None)
    assert t.set_leaf((1, 0, 6), 30) == ((6, 1), None)
    assert t.state() == {
        "octants": [
            (1, [L(20), N, N, N, N, N, N, N], 1),
            (2, [O(0), N, N, N, N, N, N, N], 1),
            (None, [O(1), N, N, N, O(5), N, N, O(3)], 3),   # root
            (2, [N, N, N, O(4), N, N, N, N], 1),
            (3, [N, N, N, N, N, N, L(10), N], 1),
            (2, [N, N, N, N, O(6), N, N, N], 1),
            (5, [N, L(30), N, N, N, N, N, N], 1)],
        "free_list": [], "root": 2, "depth": 3}
    assert t.get_leaf((6, 7, 5)) == 10 and t.get_leaf((0, 0, 0)) == 20 and t.get_leaf((1, 0, 6)) == 30 and t.get_leaf((1, 1, 1)) is None
    assert t.set_leaf((0, 0, 0), 40) == ((0, 0), 20)   # replace by adding
    assert t.get_leaf((0, 0, 0)) == 40
    t.close()


def test_octree_add_leaf_replacing(pkg):
    """octree.rs:615-644"""
    t = Tree(pkg)
    t.set_leaf((0, 0, 0), 10)
    assert t.state() == {"octants": [(None, [L(10), N, N, N, N, N, N, N], 1)], "free_list": [], "root": 0, "depth": 1}
    t.set_leaf((0, 0, 0), 20)
    assert t.state() == {"octants": [(None, [L(20), N, N, N, N, N, N, N], 1)], "free_list": [], "root": 0, "depth": 1}
    t.close()


def test_octree_remove_and_add_leaf(pkg):
    """octree.rs:648-696"""
    t = Tree(pkg)
    t.set_leaf((0, 0, 0), 10)
    t.set_leaf((1, 0, 0), 20)
    assert t.state() == {"octants": [(None, [L(10), L(20), N, N, N, N, N, N], 2)], "free_list": [], "root": 0, "depth": 1}
    assert t.remove_leaf((0, 0, 0)) == (10, (0, 0))
    assert t.remove_leaf_by_id((0, 1)) == 20
    assert t.state() == {"octants": [(None, [N] * 8, 0)], "free_list": [], "root": 0, "depth": 1}
    t.set_leaf((0, 0, 0), 30)
    assert t.state() == {"octants": [(None, [L(30), N, N, N, N, N, N, N], 1)], "free_list": [], "root": 0, "depth": 1}
    t.close()


def test_octree_move_leaf(pkg):
    """octree.rs:701-790: empty slot, itself, an occupied slot, a new parent (the tree grows)."""
    t = Tree(pkg)
    t.set_leaf((0, 0, 0), 10)
    t.set_leaf((1, 1, 1), 20)
    one = lambda ch, cnt: {"octants": [(None, ch, cnt)], "free_list": [], "root": 0, "depth": 1}
    assert t.state() == one([L(10), N, N, N, N, N, N, L(20)], 2)
    assert t.move_leaf((0, 0), (1, 0, 0)) == ((0, 1), None)
    assert t.state() == one([N, L(10), N, N, N, N, N, L(20)], 2)
    assert t.move_leaf((0, 1), (1, 0, 0)) == ((0, 1), None)
    assert t.state() == one([N, L(10), N, N, N, N, N, L(20)], 2)
    assert t.move_leaf((0, 1), (1, 1, 1)) == ((0, 7), 20)
    assert t.state() == one([N, N, N, N, N, N, N, L(10)], 1)
    assert t.move_leaf((0, 7), (2, 0, 0)) == ((2, 0), None)
    assert t.state() == {
        "octants": [(1, [N] * 8, 0), (None, [O(0), O(2), N, N, N, N, N, N], 2), (1, [L(10), N, N, N, N, N, N, N], 1)],
        "free_list": [], "root": 1, "depth": 2}
    t.close()


def test_octree_construct_octants(pkg):
    """octree.rs:794-836"""
    t = Tree(pkg)
    t.set_leaf((1, 1, 3), 2)           # previous state that has to go away
    t.construct(2, [])
    assert t.state() == {"octants": [], "free_list": [], "root": None, "depth": 0}
    t.construct(2, [((2, 2, 2), 1)])
    assert t.state() == {
        "octants": [(1, [L(1), N, N, N, N, N, N, N], 1), (None, [N, N, N, N, N, N, N, O(0)], 1)],
        "free_list": [], "root": 1, "depth": 2}
    assert t.get_leaf((2, 2, 2)) == 1 and t.get_leaf((1, 1, 1)) is None
    t.close()


def test_octree_compact(pkg):
    """octree.rs:840-866"""
    t = Tree(pkg)
    t.set_leaf((0, 1, 3), 10)
    t.set_leaf((1, 1, 3), 20)
    assert t.state() == {
        "octants": [(1, [N] * 8, 0), (None, [O(0), N, N, N, O(2), N, N, N], 2), (1, [N, N, N, N, N, N, L(10), L(20)], 2)],
        "free_list": [], "root": 1, "depth": 2}
    t.compact()
    assert t.state() == {
        "octants": [(None, [N] * 8, 0), (None, [N, N, N, N, O(2), N, N, N], 1), (1, [N, N, N, N, N, N, L(10), L(20)], 2)],
        "free_list": [0], "root": 1, "depth": 2}
    t.remove_leaf((0, 1, 3))
    t.remove_leaf((1, 1, 3))
    t.compact()
    assert t.state() == {"octants": [], "free_list": [], "root": None, "depth": 0}
    t.close()


# ---------------------------------------------------------------------------------------------------- RangeBuffer --

class Buf:
    def __init__(self, pkg, capacity):
        self.pkg, self.H = pkg, pkg.host()
        self.h = self.H.vxh_rangebuf_with_capacity(capacity)

    def insert(self, i, values):
        a = np.array(values, dtype=np.uint8)
        return self.H.vxh_rangebuf_insert(self.h, i, a.ctypes.data, len(a))

    def remove(self, i):
        self.H.vxh_rangebuf_remove(self.h, i)

    def state(self):
        n = self.H.vxh_rangebuf_bytes(self.h, None, 0)
        b = np.zeros(max(n, 1), np.uint8)
        self.H.vxh_rangebuf_bytes(self.h, b.ctypes.data, n)
        def ranges(kind):
            arr = (self.pkg.VxRange * 32)()
            k = self.H.vxh_rangebuf_ranges(self.h, kind, arr, 32)
            return [(arr[i].offset, arr[i].length) for i in range(k)]
        ids = np.zeros(96, np.uint64)
        k = self.H.vxh_rangebuf_ids(self.h, ids.ctypes.data, 32)
        return (b[:n].tolist(), ranges(0), ranges(1), {int(ids[3 * i]): (int(ids[3 * i + 1]), int(ids[3 * i + 2])) for i in range(k)})


def test_buffer_insert_remove(pkg):
    """internal.rs:288-384 (the element type does not matter to the bookkeeping: bytes here, u32 there)."""
    b = Buf(pkg, 10)
    assert b.state() == ([0] * 10, [(0, 10)], [], {})
    b.insert(1, [0, 1, 2, 3, 4]); b.insert(2, [5, 6]); b.insert(3, [7, 8, 9])
    assert b.state() == (list(range(10)), [], [(0, 10)], {1: (0, 5), 2: (5, 2), 3: (7, 3)})
    b.insert(4, [10])                                                  # exceed the initial capacity
    assert b.state() == (list(range(11)), [], [(0, 11)], {1: (0, 5), 2: (5, 2), 3: (7, 3), 4: (10, 1)})
    b.insert(3, [11])                                                  # replace existing data
    assert b.state() == ([0, 1, 2, 3, 4, 5, 6, 11, 8, 9, 10], [(8, 2)], [(0, 11)], {1: (0, 5), 2: (5, 2), 3: (7, 1), 4: (10, 1)})
    b.remove(2); b.remove(3)
    assert b.state() == ([0, 1, 2, 3, 4, 5, 6, 11, 8, 9, 10], [(5, 5)], [(0, 11)], {1: (0, 5), 4: (10, 1)})
    b.insert(5, [12, 13, 14])                                          # into the free space
    assert b.state() == ([0, 1, 2, 3, 4, 12, 13, 14, 8, 9, 10], [(8, 2)], [(0, 11)], {1: (0, 5), 4: (10, 1), 5: (5, 3)})
    b.remove(5); b.remove(4); b.remove(1)
    assert b.state() == ([0, 1, 2, 3, 4, 12, 13, 14, 8, 9, 10], [(0, 11)], [(0, 11)], {})
    pkg.host().vxh_rangebuf_free(b.h)


def test_merge_ranges(pkg):
    """internal.rs:388-455"""
    cases = [
        ("join adjacent ranges", [(0, 1), (1, 1), (2, 1)], [(0, 3)]),
        ("ignore non-adjacent ranges", [(0, 1), (2, 1)], [(0, 1), (2, 1)]),
        ("remove fully contained ranges", [(0, 5), (3, 1)], [(0, 5)]),
        ("remove and extend contained ranges", [(0, 5), (3, 5)], [(0, 8)]),
        ("works in inverse order", [(3, 5), (0, 5)], [(0, 8)]),
    ]
    for name, inp, want in cases:
        arr = (pkg.VxRange * len(inp))(*[pkg.VxRange(a, b) for a, b in inp])
        n = pkg.host().vxh_merge_ranges(arr, len(inp))
        assert [(arr[i].offset, arr[i].length) for i in range(n)] == want, name
