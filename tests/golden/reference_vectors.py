"""Golden vectors transcribed from the reference's own tests (tim-oster/voxel-rs). Each block cites its source.
Frames are (t_min, ptr, idx, parent_octant_idx, scale, is_child, is_leaf) of StackFrame (svo_shader_tests.rs:51-63).
"""

# src/graphics/svo_shader_tests.rs:293-334  esvo_tests::shader_svo_traversal
TRAVERSAL = {
    "blocks": [(31, 0, 0, 1)],
    "pos": (0.0, 0.5, 0.5), "dir": (1.0, 0.0, 0.0), "max_dst": 32.0, "cast_translucent": False,
    "frames": [
        (0.0, 0, 0, 0, 22, 1, 0), (0.0, 65, 0, 0, 21, 0, 0), (16.0, 65, 1, 0, 21, 1, 0), (16.0, 5, 0, 1, 20, 0, 0),
        (24.0, 5, 1, 1, 20, 1, 0), (24.0, 17, 0, 1, 19, 0, 0), (28.0, 17, 1, 1, 19, 1, 0), (28.0, 29, 0, 1, 18, 0, 0),
        (30.0, 29, 1, 1, 18, 1, 0), (30.0, 41, 0, 1, 17, 0, 0), (31.0, 41, 1, 1, 17, 1, 1),
    ],
    "result": {"t": 31.0, "value": 1, "face_id": 0, "pos": (31.000008, 0.5, 0.5), "uv": (0.5, 0.5), "color": (1.0, 0.0, 0.0, 1.0),
               "inside_voxel": False},
}

# src/graphics/svo_shader_tests.rs:340-489  esvo_tests::cast_inside_outside_all_axes
# each case is cast from `pos` ("inside") and from pos - normalize(dir) with t + 1 ("outside"); max_dst 100, opaque mode
ALL_AXES = {
    "blocks": [(30, 0, 0, 1), (0, 30, 0, 1), (0, 0, 30, 1), (30, 30, 30, 1)],
    "cases": [
        ("x pos", (0.5, 0.5, 0.5), (1.0, 0.0, 0.0), 29.5, 0, (30.000008, 0.5, 0.5), (0.5, 0.5)),
        ("x neg", (31.5, 0.5, 0.5), (-1.0, 0.0, 0.0), 0.5, 1, (30.999992, 0.5, 0.5), (0.5, 0.5)),
        ("y pos", (0.5, 0.5, 0.5), (0.0, 1.0, 0.0), 29.5, 2, (0.5, 30.000008, 0.5), (0.5, 0.5)),
        ("y neg", (0.5, 31.5, 0.5), (0.0, -1.0, 0.0), 0.5, 3, (0.5, 30.999992, 0.5), (0.5, 0.5)),
        ("z pos", (0.5, 0.5, 0.5), (0.0, 0.0, 1.0), 29.5, 4, (0.5, 0.5, 30.000008), (0.5, 0.5)),
        ("z neg", (0.5, 0.5, 31.5), (0.0, 0.0, -1.0), 0.5, 5, (0.5, 0.5, 30.999992), (0.5, 0.5)),
        ("diagonal pos", (0.6, 0.5, 0.6), (1.0, 1.0, 1.0), 51.095497, 2, (30.099998, 30.000008, 30.099998), (0.099998474, 0.9000015)),
        ("diagonal neg", (31.4, 31.5, 31.4), (-1.0, -1.0, -1.0), 0.86602306, 3, (30.900002, 30.999992, 30.900002), (0.9000015, 0.9000015)),
    ],
    "value": 1, "color": (1.0, 0.0, 0.0, 1.0),
}

# src/graphics/svo_shader_tests.rs:495-604  esvo_tests::uv_coords_on_all_sides  (block id 2 = "coords" atlas), max_dst 32
UV_COORDS = {
    "blocks": [(0, 0, 0, 2)],
    "cases": [  # pos, dir, expected uv, expected colour
        ((0.1, 0.1, -0.1), (0.0, 0.0, 1.0), (0.1, 0.1), (0.0, 0.0, 0.0, 1.0)),
        ((0.1, 0.5, -0.1), (0.0, 0.0, 1.0), (0.1, 0.5), (0.0, 0.4, 0.0, 1.0)),
        ((0.5, 0.1, -0.1), (0.0, 0.0, 1.0), (0.5, 0.1), (0.4, 0.0, 0.0, 1.0)),
        ((0.5, 0.5, -0.1), (0.0, 0.0, 1.0), (0.5, 0.5), (0.4, 0.4, 0.0, 1.0)),
        ((0.1, 0.1, 1.1), (0.0, 0.0, -1.0), (0.9, 0.1), (0.6, 0.0, 0.0, 1.0)),
        ((0.1, 0.5, 1.1), (0.0, 0.0, -1.0), (0.9, 0.5), (0.6, 0.4, 0.0, 1.0)),
        ((-0.1, 0.1, 0.1), (1.0, 0.0, 0.0), (0.9, 0.1), (0.6, 0.0, 0.0, 1.0)),
        ((-0.1, 0.5, 0.1), (1.0, 0.0, 0.0), (0.9, 0.5), (0.6, 0.4, 0.0, 1.0)),
        ((1.1, 0.1, 0.1), (-1.0, 0.0, 0.0), (0.1, 0.1), (0.0, 0.0, 0.0, 1.0)),
        ((1.1, 0.5, 0.1), (-1.0, 0.0, 0.0), (0.1, 0.5), (0.0, 0.4, 0.0, 1.0)),
        ((0.1, -0.1, 0.1), (0.0, 1.0, 0.0), (0.1, 0.9), (0.0, 0.6, 0.0, 1.0)),
        ((0.1, -0.1, 0.5), (0.0, 1.0, 0.0), (0.1, 0.5), (0.0, 0.4, 0.0, 1.0)),
        ((0.1, 1.1, 0.1), (0.0, -1.0, 0.0), (0.1, 0.1), (0.0, 0.0, 0.0, 1.0)),
        ((0.1, 1.1, 0.5), (0.0, -1.0, 0.0), (0.1, 0.5), (0.0, 0.4, 0.0, 1.0)),
    ],
}

# src/graphics/svo_shader_tests.rs:609-658  esvo_tests::casting_against_translucent_leafs
TRANSLUCENT = {
    "blocks": [(0, 0, 0, 3), (0, 0, 1, 3), (5, 0, 0, 3), (5, 0, 1, 4)],
    "dir": (0.75 - 0.25, 0.5 - 0.5, 1.0 - -0.1),
    "cases": [  # name, pos, cast_translucent, expected (tolerance 0.01 on t/pos/uv where given)
        ("do not cast translucent", (0.25, 0.5, -0.1), False,
         {"t": 0.1, "value": 3, "face_id": 4, "pos": (0.295, 0.5, 0.0), "uv": (0.295, 0.5), "color": (0.0, 0.0, 0.0, 0.0), "inside_voxel": False}),
        ("cast translucent with adjacent identical", (0.25, 0.5, -0.1), True,
         {"t": -1.0, "value": 0, "face_id": 0, "pos": (0.0, 0.0, 0.0), "uv": (0.0, 0.0), "color": (0.0, 0.0, 0.0, 0.0), "inside_voxel": False}),
        ("cast translucent with adjacent different", (5.25, 0.5, -0.1), True,
         {"t": 1.2, "value": 4, "face_id": 4, "pos": (5.75, 0.5, 1.0), "uv": (0.75, 0.5), "color": (0.0, 1.0, 0.0, 1.0), "inside_voxel": False}),
    ],
}

# src/graphics/svo_shader_tests.rs:663-701  esvo_tests::detect_inside_leaf_voxel
INSIDE_LEAF = {
    "blocks": [(0, 0, 0, 1)],
    "cases": [
        ("inside block", (0.5, 0.5, 0.5), (1.0, 0.0, 0.0),
         {"t": -1.0, "value": 0, "face_id": 0, "pos": (0.0, 0.0, 0.0), "uv": (0.0, 0.0), "color": (0.0, 0.0, 0.0, 0.0), "inside_voxel": True}),
        ("outside block", (-0.5, 0.5, 0.5), (1.0, 0.0, 0.0),
         {"t": 0.5, "value": 1, "face_id": 0, "pos": (8e-6, 0.5, 0.5), "uv": (0.5, 0.5), "color": (1.0, 0.0, 0.0, 1.0), "inside_voxel": False}),
    ],
}

# src/graphics/svo_shader_tests.rs:707-753  esvo_tests::check_at_higher_coordinates (chunk at SVO position 15,15,15)
HIGHER_COORDS = {
    "svo_pos": (15, 15, 15),
    "blocks": [(x, y, z, 1) for x in range(32) for z in range(32) for y in range(5)],
    "pos": (484.9203, 485.95938, 493.8467), "dir": (0.0, -1.0, 0.0), "max_dst": 10.0, "cast_translucent": False,
    "frames": [
        (0.0, 0, 7, 0, 22, 1, 0), (0.0, 11009, 7, 7, 21, 1, 0), (0.0, 11057, 7, 7, 20, 1, 0), (0.0, 11069, 7, 7, 19, 1, 0),
        (0.0, 11081, 0, 7, 18, 1, 0), (0.0, 5, 4, 0, 17, 1, 0), (0.0, 17, 7, 4, 16, 1, 0), (0.0, 1397, 0, 7, 15, 1, 0),
        (0.0, 2021, 6, 0, 14, 0, 0), (0.9593506, 2021, 4, 0, 14, 1, 1),
    ],
    "result": {"t": 0.9593506, "value": 1, "face_id": 3, "pos": (484.9203, 484.99994, 493.84668), "uv": (0.9202881, 0.8466797),
               "color": (1.0, 0.0, 0.0, 1.0), "inside_voxel": False},
}

# src/graphics/svo.rs:402-449  svo_tests::raycast  (blocks (0,0,0)=1, (1,0,0)=1; NO compact; content registry of svo.rs:323-338)
PICKER_RAYCAST = {
    "blocks": [(0, 0, 0, 1), (1, 0, 0, 1)],
    "rays": [((0.5, 1.5, 0.5), (0.0, -1.0, 0.0), 1.0), ((0.5, 0.5, 0.5), (1.0, 0.0, 0.0), 1.0), ((0.5, 0.5, -2.0), (0.0, 0.0, 1.0), 1.0)],
    "expected": [  # dst, inside_voxel, pos, normal   (dst/pos tolerance 1e-4)
        (0.5, False, (0.5, 1.0, 0.5), (0.0, 1.0, 0.0)),
        (0.5, True, (1.0, 0.5, 0.5), (-1.0, 0.0, 0.0)),
        (-1.0, False, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)),
    ],
}


# ---------------------------------------------------------------------------------------------------------- CSVO ----
# src/graphics/svo_shader_tests.rs:756-1224 csvo_tests. cast_inside_outside_all_axes, uv_coords_on_all_sides,
# casting_against_translucent_leafs and detect_inside_leaf_voxel (:806-1176) are line for line the ESVO cases above (same
# scenes, same expected results); the two step traces differ because the node encoding does.
# Frames: (t_min, ptr, idx, depth [in StackFrame.parent_octant_idx], scale, is_child, is_leaf, crossed_boundary, next_ptr).
U32_MAX = 4294967295

# svo_shader_tests.rs:763-804  csvo_tests::shader_svo_traversal
CSVO_TRAVERSAL = dict(TRAVERSAL, frames=[
    (0.0, 21, 0, 6, 22, 1, 0, 1, 0), (0.0, 9, 0, 5, 21, 0, 0, 0, U32_MAX), (16.0, 9, 1, 5, 21, 1, 0, 0, 12),
    (16.0, 12, 0, 4, 20, 0, 0, 0, U32_MAX), (24.0, 12, 1, 4, 20, 1, 0, 0, 15), (24.0, 15, 0, 3, 19, 0, 0, 0, U32_MAX),
    (28.0, 15, 1, 3, 19, 1, 0, 0, 17), (28.0, 17, 0, 2, 18, 0, 0, 0, U32_MAX), (30.0, 17, 1, 2, 18, 1, 0, 0, 20),
    (30.0, 20, 0, 1, 17, 0, 0, 0, U32_MAX), (31.0, 20, 1, 1, 17, 1, 1, 0, 23),
])

# svo_shader_tests.rs:1177-1223  csvo_tests::check_at_higher_coordinates
CSVO_HIGHER_COORDS = dict(HIGHER_COORDS, frames=[
    (0.0, 21814, 7, 9, 22, 1, 0, 0, 21826), (0.0, 21826, 7, 8, 21, 1, 0, 0, 21829), (0.0, 21829, 7, 7, 20, 1, 0, 0, 21832),
    (0.0, 21832, 7, 6, 19, 1, 0, 1, 0), (0.0, 20485, 0, 5, 18, 1, 0, 0, 20494), (0.0, 20494, 4, 4, 17, 1, 0, 0, 20662),
    (0.0, 20662, 7, 3, 16, 1, 0, 0, 20736), (0.0, 20736, 0, 2, 15, 1, 0, 0, 20739), (0.0, 20739, 6, 1, 14, 0, 0, 0, U32_MAX),
    (0.9593506, 20739, 4, 1, 14, 1, 1, 0, 20744),
])
