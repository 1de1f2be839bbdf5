"""Host logic above the C ABI (PickerBatch, registry, camera) against the reference's own tests, and the ABI
itself: the library loads without a GPU, exports every symbol include/voxelrt.h declares, has the reference's
record sizes, and refuses to run without CUDA (no CPU fallback). CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def tasks_of(pkg, rays, aabbs):
    r = np.ascontiguousarray(np.array(rays, np.float32).reshape(-1, 7))
    a = np.ascontiguousarray(np.array(aabbs, np.float32).reshape(-1, 9))
    out = np.zeros(256, dtype=pkg.TASK_DTYPE)
    n = pkg.host().vxh_picker_serialize(r.ctypes.data, len(r), a.ctypes.data, len(a), out.ctypes.data, len(out))
    return out[:n]


RAYS = [((1, 0, 1), (0, 1, 0), 20.0), ((2, 0, 2), (1, 0, 0), 40.0)]
AABBS = [((0.5, 0.0, 0.5), (-0.5, 0.0, -0.5), (1.0, 1.0, 1.0)), ((0, 0, 0), (0, 0, 0), (1.5, 1.5, 1.5))]
flat = lambda rows: [tuple(p) + tuple(d) + ((m,) if not isinstance(m, tuple) else tuple(m)) for p, d, m in rows]


def expected_aabb_tasks(pos, offset, extents):
    """The enumeration the reference's expected vector spells out (svo_picker.rs:337-416): lattice points x,y,z
    ascending, per point the axes x,y,z on which the point lies on the box surface."""
    n = [int(np.ceil(e)) for e in extents]
    step = [np.float32(e) / np.float32(k) for e, k in zip(extents, n)]
    out = []
    for x in range(n[0] + 1):
        for y in range(n[1] + 1):
            for z in range(n[2] + 1):
                p = (x, y, z)
                for i in range(3):
                    if p[i] != 0 and p[i] != n[i]:
                        continue
                    d = [0.0, 0.0, 0.0]
                    d[i] = -1.0 if p[i] == 0 else 1.0
                    out.append((tuple(np.float32(pos[k]) + np.float32(offset[k]) + np.float32(p[k]) * step[k] for k in range(3)), tuple(d)))
    return out


def test_picker_batch_serialization(pkg):
    """picker_batch_serialization, src/graphics/svo_picker.rs:310-418: 2 rays + unit AABB (24) + 1.5-AABB (54) = 80 tasks."""
    t = tasks_of(pkg, flat(RAYS), flat(AABBS))
    assert len(t) == 80
    assert t[0]["max_dst"] == 20.0 and tuple(t[0]["pos"]) == (1, 0, 1) and tuple(t[0]["dir"]) == (0, 1, 0)
    assert t[1]["max_dst"] == 40.0 and tuple(t[1]["pos"]) == (2, 0, 2) and tuple(t[1]["dir"]) == (1, 0, 0)
    exp = expected_aabb_tasks(*AABBS[0]) + expected_aabb_tasks(*AABBS[1])
    assert len(exp) == 78
    for got, (pos, d) in zip(t[2:], exp):
        assert got["max_dst"] == 10.0 and tuple(got["pos"]) == tuple(float(v) for v in pos) and tuple(got["dir"]) == d
    # literal rows of the reference vector (svo_picker.rs:338-343, 363-368, 389-390, 414-416)
    lit = {2: ((0, 0, 0), (-1, 0, 0)), 3: ((0, 0, 0), (0, -1, 0)), 4: ((0, 0, 0), (0, 0, -1)), 5: ((0, 0, 1), (-1, 0, 0)),
           7: ((0, 0, 1), (0, 0, 1)), 26: ((0, 0, 0), (-1, 0, 0)), 29: ((0, 0, 0.75), (-1, 0, 0)), 31: ((0, 0, 1.5), (-1, 0, 0)),
           52: ((0.75, 0.75, 0), (0, 0, -1)), 53: ((0.75, 0.75, 1.5), (0, 0, 1)), 77: ((1.5, 1.5, 1.5), (1, 0, 0)),
           79: ((1.5, 1.5, 1.5), (0, 0, 1))}
    for i, (pos, d) in lit.items():
        assert tuple(t[i]["pos"]) == pos and tuple(t[i]["dir"]) == d, i


def test_picker_batch_deserialization(pkg):
    """picker_batch_deserialization, src/graphics/svo_picker.rs:421-536: per-axis minimum over the AABB's probes."""
    rays = [((0, 0, 0), (-1, 0, 0), 20.0), ((0, 0, 0), (1, 0, 0), 20.0)]
    res = np.zeros(80, dtype=pkg.RESULT_DTYPE)
    res["dst"] = -1.0
    res[1] = (10.0, 1, (-1, 0, 0), (10, 0, 0))
    # aabb 1 (results 2..25), dst column of svo_picker.rs:442-465
    a1 = [8, 8, 8, -1, -1, -1, -1, 4, -1, -1, -1, 4, 4, -1, -1, -1, 7, -1, -1, -1, -1, 2, -1, 1]
    # aabb 2 (results 26..79), dst column of svo_picker.rs:467-520
    a2 = [-1] * 54
    for i, v in {3: 9, 4: 8, 12: 5, 26: 7, 33: 5, 38: 5, 40: 3, 49: 1, 50: 4}.items():
        a2[i] = v
    assert len(a1) == 24
    res["dst"][2:26] = a1
    res["dst"][26:26 + len(a2)] = a2
    r = np.ascontiguousarray(np.array(flat(rays), np.float32))
    a = np.ascontiguousarray(np.array(flat(AABBS), np.float32))
    ro, ao = np.zeros((2, 8), np.float32), np.zeros((2, 6), np.float32)
    pkg.host().vxh_picker_deserialize(r.ctypes.data, 2, a.ctypes.data, 2, res.ctypes.data, 80, ro.ctypes.data, ao.ctypes.data)
    assert ro[0].tolist() == [-1, 0, 0, 0, 0, 0, 0, 0]
    assert ro[1].tolist() == [10, 1, -1, 0, 0, 10, 0, 0]
    assert ao[0].tolist() == [8, 7, 8, 2, 4, 1]          # AabbResult { neg: (8,7,8), pos: (2,4,1) }  svo_picker.rs:532
    assert ao[1].tolist() == [9, 8, 7, 1, 4, 3]          # AabbResult { neg: (9,8,7), pos: (1,4,3) }  svo_picker.rs:533


def test_player_aabb_is_32_rays(pkg):
    """A 0.8 x 1.8 x 0.8 player box (src/gamelogic/game.rs:73) expands to 32 probe rays."""
    t = tasks_of(pkg, [], [((0, 0, 0, 0, 0, 0, 0.8, 1.8, 0.8))])
    assert len(t) == 32


def test_registry_material_table(pkg):
    """blocks::new_registry (src/gamelogic/content.rs:20-60) -> MaterialInstance table (svo_registry.rs:135-165)."""
    reg = pkg.content_registry(pkg.load_atlas())
    m = reg.materials()
    assert len(m) == 13 and m.dtype.itemsize == 32
    assert m[0]["tex"].tolist() == [-1] * 6 and m[0]["specular_pow"] == 0
    # grass: top=grass_top(4) side=grass_side(2) bottom=dirt(0), normals = name + "_normal"
    assert m[1]["tex"].tolist() == [4, 2, 0, 5, 3, 1] and m[1]["specular_pow"] == 14.0 and np.isclose(m[1]["specular_strength"], 0.4)
    assert m[3]["tex"].tolist() == [6, 6, 6, 7, 7, 7] and m[3]["specular_pow"] == 70.0
    assert m[5]["tex"].tolist() == [10, 10, 10, -1, -1, -1]          # glass: no normal maps
    assert m[9]["tex"].tolist() == [18, 16, 18, 19, 17, 19]          # oak log
    tex, mips = reg.textures()
    assert tex.shape == (25, 64, 64, 4) and mips == 6
    # images are flipped vertically on load (texture_array.rs:92,126)
    assert np.array_equal(tex[0], pkg.load_atlas()["dirt"][::-1])


def test_view_matrix(pkg):
    """look_to_rh(pos, fwd, up).invert() (svo.rs:197): camera->world matrix maps the eye to the origin column and -z to fwd."""
    eye, fwd, up = np.array([2.5, 2.5, 7.5], np.float32), np.array([0, 0, -1], np.float32), np.array([0, 1, 0], np.float32)
    out = np.zeros(16, np.float32)
    pkg.host().vxh_look_to_rh_inverted(eye.ctypes.data, fwd.ctypes.data, up.ctypes.data, out.ctypes.data)
    m = out.reshape(4, 4).T   # column-major -> rows
    assert np.allclose(m, [[1, 0, 0, 2.5], [0, 1, 0, 2.5], [0, 0, 1, 7.5], [0, 0, 0, 1]])
    fwd = np.array([1, -0.3, 0.2], np.float32)
    pkg.host().vxh_look_to_rh_inverted(eye.ctypes.data, fwd.ctypes.data, up.ctypes.data, out.ctypes.data)
    m = out.reshape(4, 4).T.astype(np.float64)
    f = fwd / np.linalg.norm(fwd)
    assert np.allclose(m[:3, :3] @ np.array([0, 0, -1.0]), f, atol=1e-6)
    assert np.allclose(m[:3, 3], eye) and np.allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3), atol=1e-6)


def test_abi_symbols_and_sizes(pkg):
    L = pkg.lib()
    header = open(os.path.join(pkg.ROOT, "include", "voxelrt.h")).read()
    declared = set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(pkg.VX_SYMBOLS), declared ^ set(pkg.VX_SYMBOLS)
    for sym in declared:
        assert getattr(L, sym) is not None
    assert C.sizeof(pkg.VxMaterial) == 32                    # MaterialInstance, svo_registry.rs:29-40
    assert pkg.TASK_DTYPE.itemsize == 48 and pkg.RESULT_DTYPE.itemsize == 48   # svo_picker.rs:13-32
    assert C.sizeof(pkg.VxDebugFrame) == 36 and C.sizeof(pkg.VxRange) == 16
    assert b"sm_100a" in L.vx_build_info() and b"fmad=false" in L.vx_build_info()


def test_no_cpu_fallback(pkg):
    """Without a CUDA device vx_create must fail loudly; nothing in the product computes on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    cfg = pkg.VxConfig(device=0, flags=0, svo_capacity_bytes=1 << 20, max_width=8, max_height=8, max_rays=8)
    ctx = C.c_void_p()
    rc = pkg.lib().vx_create(C.byref(cfg), C.byref(ctx))
    assert rc == -3 and not ctx.value
    assert b"no CPU fallback" in pkg.lib().vx_last_error(None)
    with pytest.raises(pkg.VxError):
        pkg.Svo(pkg.Registry(1).add_texture("t", np.zeros((4, 4, 4), np.uint8)).add_material(0), size_mb=1)
    # the product package must not reference the oracle
    for root, _, files in os.walk(os.path.join(pkg.ROOT, "voxel-rs_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh")):
                src = open(os.path.join(root, f), errors="ignore").read()
                for needle in ("liboracle", "load_oracle", "vxo_", "oracle.binding", "oracle/binding"):
                    assert needle not in src, (f, needle)
