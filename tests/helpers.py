"""Shared scene builders for the tests. Mirrors the fixtures of the reference's own tests."""
import numpy as np


def shader_test_atlas():
    """The 4x4 synthetic atlas of src/graphics/svo_shader_tests.rs:117-146, as given to add_rgba8 (row 0 = top)."""
    R, G, T = (255, 0, 0, 255), (0, 255, 0, 255), (0, 0, 0, 0)
    full = np.array([[R] * 4] * 4, dtype=np.uint8)
    coords = np.array([[(51 * x, 153 - 51 * y, 0, 255) for x in range(4)] for y in range(4)], dtype=np.uint8)
    t1 = np.array([[T, T, R, R]] * 4, dtype=np.uint8)
    t2 = np.array([[T, T, G, G]] * 4, dtype=np.uint8)
    return {"full": full, "coords": coords, "transparent_1": t1, "transparent_2": t2}


def shader_test_registry(pkg):
    """create_test_materials(), svo_shader_tests.rs:117-202: 1 mip level, materials 0..4 without normals."""
    r = pkg.Registry(mip_levels=1)
    for name, img in shader_test_atlas().items():
        r.add_texture(name, img)
    r.add_material(0)
    r.add_material(1, all_sides="full")
    r.add_material(2, all_sides="coords")
    r.add_material(3, all_sides="transparent_1")
    r.add_material(4, all_sides="transparent_2")
    return r


def shader_test_world(pkg, blocks, svo_pos=(0, 0, 0), fmt=0):
    """create_test_world(), svo_shader_tests.rs:78-115: pooled chunk storage, set_block, storage.compact(), one leaf."""
    w = pkg.World(fmt=fmt)
    w.set_leaf_blocks(svo_pos, blocks, uid=1, lod=5, compact=True)
    w.serialize()
    return w


def svo_render_test_registry(pkg, atlas):
    """create_voxel_registry(), src/graphics/svo.rs:323-338."""
    r = pkg.Registry(mip_levels=6)
    for name, stem in [("stone", "stone"), ("stone_normal", "stone_n"), ("dirt", "dirt"), ("dirt_normal", "dirt_n"),
                       ("grass_side", "grass_side"), ("grass_side_normal", "grass_side_n"), ("grass_top", "grass_top"),
                       ("grass_top_normal", "grass_top_n")]:
        r.add_texture(name, atlas[stem])
    r.add_material(0)
    r.add_material(1, (70.0, 0.4), all_sides="stone", with_normals=True)
    r.add_material(2, (14.0, 0.4), top="grass_top", side="grass_side", bottom="dirt", with_normals=True)
    return r


def svo_render_test_blocks():
    """Scene of svo_tests::render, src/graphics/svo.rs:347-363."""
    b = [(x, 0, z, 1) for x in range(5) for z in range(5)]
    for z in (1, 3):
        b += [(1, 1, z, 2), (3, 1, z, 2), (1, 3, z, 2), (3, 3, z, 2)]
    return b


def svo_render_test_params(pkg, width=640, height=490):
    """RenderParams of svo_tests::render, src/graphics/svo.rs:371-383."""
    return pkg.render_params(cam_pos=(2.5, 2.5, 7.5), cam_fwd=(0, 0, -1), cam_up=(0, 1, 0), fov_y_deg=72.0, aspect=width / height,
                             ambient=0.3, selected_voxel=(1.0, 1.0, 3.0), render_shadows=True, shadow_distance=500.0)


def oracle_scene(ora, world, registry):
    tex, mips = registry.textures()
    return ora.Scene(world.gpu_buffer(), registry.materials().tobytes(), tex, mips, fmt=getattr(world, "fmt", 0))


def diff_images(a8, b8):
    """diff_images(), src/graphics/framebuffer.rs:120-134: mean |dRGB| / 255."""
    d = np.abs(a8[..., :3].astype(np.int64) - b8[..., :3].astype(np.int64)).sum()
    return d / (255.0 * 3.0 * a8.shape[0] * a8.shape[1])


def random_tasks(pkg, n, lo, hi, max_dst, seed=0):
    rng = np.random.default_rng(seed)
    t = np.zeros(n, dtype=pkg.TASK_DTYPE)
    t["max_dst"] = max_dst
    t["pos"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    t["dir"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return t


MC_CENTER = (-69, 1, 52)   # engine chunk at the middle of the fixture area


def mc_world(pkg, fmt=0):
    """The reference's bundled Minecraft world (tests/golden/mc_world.npz, made by make_mc_fixture.py) as a world SVO:
    166 engine chunks around MC_CENTER at full detail, set the way systems::worldsvo::Svo::set_chunk does."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc_world.npz"))
    w = pkg.World(radius=6, center=MC_CENTER, seed=1, fmt=fmt)   # 9 x 8 columns: the corner (4, 4) is inside the radius-6 disc
    for uid, (c, b) in enumerate(zip(z["coords"], z["blocks"])):
        sp = w.cnv_chunk_pos(tuple(int(v) for v in c))
        assert sp is not None
        w.set_leaf_dense(sp, b.astype(np.uint32), uid=1000 + uid, lod=5)
    w.serialize()
    return w


def mc_params(pkg, w, h, shadows=False):
    """Fixed camera of the BASELINE configs[0] frame: over the water, looking at a wooded island (leaves, logs, grass, sand)."""
    return pkg.render_params(cam_pos=(-2090.0, 75.0, 1690.0), cam_fwd=(0.6, -0.4, 1.0), fov_y_deg=72.0, aspect=w / h, render_shadows=shadows)


# ---- the reference's generated world (gamelogic::world tests::end_to_end, src/gamelogic/world.rs:461-498) ----

E2E_SIZE = (1024, 768)


def e2e_world(pkg, fmt=0):
    """World::new(.., 72.0, true, 15, false, None, 800) after loading finished: Generator::new(1, cfg) terrain (noise 0.8.2 Perlin +
    splines), chunk disc of radius 15 around the player's chunk, y-chunks 0..8, LOD rule."""
    w = pkg.World(radius=15, center=(-1, 2, 5), seed=1, fmt=fmt, terrain="reference")
    w.generate(0, 8)
    w.serialize()
    return w


def e2e_params(pkg):
    """player at (-24, 80, 174), euler_rotation (0, -90 deg, 0) -> Entity::get_forward (physics.rs:21-27); World::render
    (gamelogic/world.rs:269-281): ambient 0.3, sun (-1,-1,-1) normalised, shadows on, shadow distance 500."""
    import math
    yaw = np.float32(math.radians(-90.0))
    fwd = (float(np.cos(yaw, dtype=np.float32)), 0.0, float(np.sin(yaw, dtype=np.float32)))
    w, h = E2E_SIZE
    return pkg.render_params(cam_pos=(-24.0, 80.0, 174.0), cam_fwd=fwd, cam_up=(0, 1, 0), fov_y_deg=72.0, aspect=w / h, ambient=0.3,
                             render_shadows=True, shadow_distance=500.0)


def e2e_expected():
    import os
    from PIL import Image
    return np.asarray(Image.open(os.path.join(os.path.dirname(__file__), "golden", "gamelogic_world_end_to_end_expected.png")).convert("RGBA"))
