"""systems::worldsvo::Svo's re-centring of the SVO window (src/systems/worldsvo.rs:133-196) on the host mirror: the reference's
shift_chunks unit tests (:249-375) and, through the oracle, the two properties that matter to the ray caster — a shifted world
renders exactly like one built at the new centre, and uploading only the dirty ranges of the shift is enough. CPU only."""
import ctypes as C

import numpy as np

import helpers


class Esvo32:
    """Esvo<u32> + the leaf_ids map of worldsvo::Svo, like the reference's test set-up."""

    def __init__(self, pkg):
        self.H = pkg.host()
        self.h = self.H.vxh_esvo32_new()
        self.leaf_ids = {}

    def set_leaf(self, chunk, svo_pos, value):
        out = (C.c_uint32 * 2)()
        self.H.vxh_esvo32_set_leaf(self.h, *svo_pos, value, 1, out)
        self.leaf_ids[chunk] = (out[0], out[1])
        return (out[0], out[1])

    def get_leaf(self, svo_pos):
        v = self.H.vxh_esvo32_get_leaf(self.h, *svo_pos)
        return None if v < 0 else v

    def shift(self, center, dst):
        keys = sorted(self.leaf_ids)
        n = len(keys)
        chunks = np.array(keys, dtype=np.int32).reshape(-1, 3).copy() if n else np.zeros((1, 3), np.int32)
        ids = np.array([self.leaf_ids[k] for k in keys], dtype=np.uint32).reshape(-1, 2).copy() if n else np.zeros((1, 2), np.uint32)
        k = self.H.vxh_kat_shift_chunks(self.h, *center, dst, chunks.ctypes.data, ids.ctypes.data, n)
        self.leaf_ids = {tuple(int(v) for v in chunks[i]): (int(ids[i][0]), int(ids[i][1])) for i in range(k)}

    def close(self):
        self.H.vxh_esvo32_free(self.h)


def three_chunks(pkg):
    e = Esvo32(pkg)
    c0 = e.set_leaf((-1, 0, 0), (0, 1, 1), 1)
    c1 = e.set_leaf((0, 0, 0), (1, 1, 1), 2)
    c2 = e.set_leaf((1, 0, 0), (2, 1, 1), 3)
    assert e.leaf_ids == {(-1, 0, 0): c0, (0, 0, 0): c1, (1, 0, 0): c2}
    assert [e.get_leaf(p) for p in ((0, 1, 1), (1, 1, 1), (2, 1, 1))] == [1, 2, 3]
    return e, c0, c1, c2


def test_shift_chunks_x_positive(pkg):
    """worldsvo.rs:249-300"""
    e, c0, c1, c2 = three_chunks(pkg)
    row = lambda: [e.get_leaf(p) for p in ((0, 1, 1), (1, 1, 1), (2, 1, 1))]
    e.shift((1, 0, 0), 1)
    assert e.leaf_ids == {(0, 0, 0): c0, (1, 0, 0): c1} and row() == [2, 3, None]
    e.shift((2, 0, 0), 1)
    assert e.leaf_ids == {(1, 0, 0): c0} and row() == [3, None, None]
    e.shift((3, 0, 0), 1)
    assert e.leaf_ids == {} and row() == [None, None, None]
    e.close()


def test_shift_chunks_x_negative(pkg):
    """worldsvo.rs:304-352"""
    e, c0, c1, c2 = three_chunks(pkg)
    row = lambda: [e.get_leaf(p) for p in ((0, 1, 1), (1, 1, 1), (2, 1, 1))]
    e.shift((-1, 0, 0), 1)
    assert e.leaf_ids == {(-1, 0, 0): c1, (0, 0, 0): c2} and row() == [None, 1, 2]
    e.shift((-2, 0, 0), 1)
    assert e.leaf_ids == {(-1, 0, 0): c2} and row() == [None, None, 1]
    e.shift((-3, 0, 0), 1)
    assert e.leaf_ids == {} and row() == [None, None, None]
    e.close()


def test_shift_chunks_x_out_of_range(pkg):
    """worldsvo.rs:357-375: a leap of the centre past the whole window."""
    e, c0, c1, c2 = three_chunks(pkg)
    e.shift((3, 0, 0), 1)
    assert e.leaf_ids == {} and [e.get_leaf(p) for p in ((0, 1, 1), (1, 1, 1), (2, 1, 1))] == [None, None, None]
    e.close()


def _render(pkg, ora, reg, world, buf, cam_world, w=192, h=108):
    p = pkg.render_params(cam_pos=cam_world, cam_fwd=(0.3, -0.5, -1.0), fov_y_deg=72.0, aspect=w / h, render_shadows=True)
    q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
    q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
    tex, mips = reg.textures()
    scene = ora.Scene(buf, reg.materials().tobytes(), tex, mips, fmt=world.fmt)
    img, cnt = scene.render(pkg.to_vx_render_params(q), w, h)
    return img, cnt


def test_recentred_world_renders_like_a_fresh_one(pkg, ora):
    """The window follows the player by one chunk in +x and then in -z. Per format: (1) the shifted world and a world built at the new centre
    with the same chunks give bit-identical frames (same SVO coordinates, the same leaves; the shifted octree keeps the octants that a
    move emptied, like the reference, so its rays may take a few more PUSH/POP steps through empty space); (2) a GPU-buffer
    image that only ever received write_changes_to()'s dirty ranges renders the same again; (3) a shift re-serializes no chunk: the
    dirty bytes are the world-root octants, a small fraction of the buffer."""
    reg = pkg.content_registry(pkg.load_atlas())
    for fmt in (0, 1):
        c0 = (-1, 2, 5)
        a = pkg.World(radius=3, center=c0, seed=1, no_lod=True, fmt=fmt, terrain="reference")
        a.generate(0, 8)
        a.serialize()
        before = {tuple(c) for c in a.chunks().tolist()}
        image = np.zeros(a.header_bytes + a.size_bytes + 4096, dtype=np.uint8)   # the "GPU buffer" that only sees dirty ranges
        assert a.write_changes_to(image)
        for c1 in ((0, 2, 5), (0, 2, 4)):
            assert a.set_center(c1)
            a.serialize()
            dirty = sum(l for _, l in a.dirty_ranges())
            assert 0 < dirty < a.size_bytes // 20, (dirty, a.size_bytes)
            assert a.write_changes_to(image)
            kept = {tuple(c) for c in a.chunks().tolist()}
            assert kept < before and all(a.cnv_chunk_pos(c) is not None for c in kept)
            b = pkg.World(radius=3, center=c1, seed=1, no_lod=True, fmt=fmt, terrain="reference")
            b.generate(0, 8)
            for c in b.chunks().tolist():
                if tuple(c) not in kept:
                    b.remove_chunk(c)
            b.serialize()
            assert {tuple(c) for c in b.chunks().tolist()} == kept
            cam = (32.0 * c1[0] + 8.0, 110.0, 32.0 * c1[2] + 20.0)
            img_a, cnt_a = _render(pkg, ora, reg, a, a.gpu_buffer(), cam)
            img_b, cnt_b = _render(pkg, ora, reg, b, b.gpu_buffer(), cam)
            img_i, cnt_i = _render(pkg, ora, reg, a, image, cam)
            assert cnt_a["primary_rays"] == 192 * 108 and cnt_a["leaf_tests"] > 1000
            assert img_a.tobytes() == img_b.tobytes(), (fmt, c1)
            same = ("primary_rays", "shadow_rays", "leaf_tests", "tex_fetches")
            assert all(cnt_a[k] == cnt_b[k] for k in same) and cnt_a["steps"] >= cnt_b["steps"], (cnt_a, cnt_b)
            assert img_a.tobytes() == img_i.tobytes() and cnt_a == cnt_i, (fmt, c1)
            before = kept
