"""voxel-rs_b200 — B200-native (sm_100a CUDA) replacement for voxel-rs's GLSL ray-cast path.

Python surface = thin ctypes bindings over two in-tree shared libraries:

  libvoxelrt.so        the product: CUDA kernels behind the C ABI of include/voxelrt.h
  libvoxelrs_host.so   C++ mirror of the reference's Rust host interfaces (graphics::Svo,
                       PickerBatch, VoxelRegistry, world::hds::esvo, systems::worldsvo coordinates)

There is NO CPU fallback: every render / raycast goes through vx_* into CUDA kernels and raises
if the extension is missing or no GPU is present. The directory name contains a '-', so import it
through `__graft_entry__.load_pkg()` (importlib) rather than an `import` statement.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)


class NativeMissing(RuntimeError):
    pass


# ------------------------------------------------------------------ C structs --

class VxConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("svo_capacity_bytes", C.c_uint64),
                ("max_width", C.c_uint32), ("max_height", C.c_uint32), ("max_rays", C.c_uint64)]


class VxMaterial(C.Structure):
    _fields_ = [("specular_pow", C.c_float), ("specular_strength", C.c_float), ("tex_top", C.c_int32),
                ("tex_side", C.c_int32), ("tex_bottom", C.c_int32), ("tex_top_normal", C.c_int32),
                ("tex_side_normal", C.c_int32), ("tex_bottom_normal", C.c_int32)]


class VxRenderParams(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("fov_y_rad", C.c_float), ("aspect_ratio", C.c_float),
                ("ambient_intensity", C.c_float), ("light_dir", C.c_float * 3), ("cam_pos", C.c_float * 3),
                ("highlight_pos", C.c_float * 3), ("render_shadows", C.c_uint32), ("shadow_distance", C.c_float)]


class VxShard(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("world_size", C.c_uint32)]


class VxFrameStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("steps", C.c_uint64), ("pushes", C.c_uint64),
                ("leaf_tests", C.c_uint64), ("tex_fetches", C.c_uint64), ("kernel_ms", C.c_float), ("trace_ms", C.c_float),
                ("shade_ms", C.c_float), ("shadow_ms", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class VxStats(C.Structure):
    _fields_ = [("used_bytes", C.c_uint64), ("capacity_bytes", C.c_uint64), ("depth", C.c_uint32)]


class VxRange(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("length", C.c_uint64)]


class VxChunkInfo(C.Structure):
    _fields_ = [("offset_bytes", C.c_uint64), ("length_bytes", C.c_uint64), ("child_mask", C.c_uint8), ("leaf_mask", C.c_uint8),
                ("depth", C.c_uint8), ("_pad", C.c_uint8 * 5)]


CHUNK_INFO_DTYPE = np.dtype([("offset_bytes", "<u8"), ("length_bytes", "<u8"), ("child_mask", "u1"), ("leaf_mask", "u1"), ("depth", "u1"),
                             ("_pad", "u1", 5)])


class VxDebugFrame(C.Structure):
    _fields_ = [("t_min", C.c_float), ("ptr", C.c_uint32), ("idx", C.c_uint32), ("parent_octant_idx", C.c_uint32),
                ("scale", C.c_int32), ("is_child", C.c_int32), ("is_leaf", C.c_int32), ("crossed_boundary", C.c_int32),
                ("next_ptr", C.c_uint32)]

    def as_tuple(self):
        return (self.t_min, self.ptr, self.idx, self.parent_octant_idx, self.scale, self.is_child, self.is_leaf)


class VxOctreeResult(C.Structure):
    _fields_ = [("t", C.c_float), ("value", C.c_uint32), ("face_id", C.c_int32), ("pos", C.c_float * 3), ("uv", C.c_float * 2),
                ("color", C.c_float * 4), ("lod", C.c_float), ("inside_voxel", C.c_uint32)]

    def as_dict(self):
        return {"t": self.t, "value": self.value, "face_id": self.face_id, "pos": tuple(self.pos), "uv": tuple(self.uv),
                "color": tuple(self.color), "lod": self.lod, "inside_voxel": bool(self.inside_voxel)}


class VxhRenderParams(C.Structure):
    """graphics::svo::RenderParams (src/graphics/svo.rs:85-106)."""
    _fields_ = [("ambient_intensity", C.c_float), ("light_dir", C.c_float * 3), ("cam_pos", C.c_float * 3),
                ("cam_fwd", C.c_float * 3), ("cam_up", C.c_float * 3), ("fov_y_rad", C.c_float), ("aspect_ratio", C.c_float),
                ("has_selected_voxel", C.c_int32), ("selected_voxel", C.c_float * 3), ("render_shadows", C.c_int32),
                ("shadow_distance", C.c_float)]


# VxPickerTask / VxPickerResult as numpy record dtypes (48 bytes each, include/voxelrt.h)
TASK_DTYPE = np.dtype({"names": ["max_dst", "pos", "dir"], "formats": ["<f4", ("<f4", 3), ("<f4", 3)], "offsets": [0, 16, 32], "itemsize": 48})
RESULT_DTYPE = np.dtype({"names": ["dst", "inside_voxel", "pos", "normal"], "formats": ["<f4", "<u4", ("<f4", 3), ("<f4", 3)],
                         "offsets": [0, 4, 16, 32], "itemsize": 48})

HIT_DTYPE = np.dtype([("t", "<f4"), ("value", "<u4"), ("face_id", "<i4"), ("pos", "<f4", 3), ("uv", "<f4", 2)])   # VxHitRecord, 32 bytes

VX_FLAG_NO_L2_WINDOW = 1
VX_FLAG_SVO_CSVO = 4
FORMAT_ESVO, FORMAT_CSVO = 0, 1
OPT_COUNT, OPT_CTAS_PER_SM, OPT_L2_WINDOW, OPT_REFILL, OPT_REFILL_PICKER, OPT_RGBA8_OUT, OPT_TMA, OPT_REFILL_SHADOW, OPT_OVERLAP, OPT_CLIP, OPT_MORTON = 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13

# every symbol include/voxelrt.h declares (checked by tests/test_abi.py)
VX_SYMBOLS = [
    "vx_create", "vx_destroy", "vx_last_error", "vx_set_materials", "vx_set_textures", "vx_svo_host_mirror", "vx_svo_commit",
    "vx_svo_set_hot_range", "vx_svo_commit_packed_device", "vx_svo_pack_dirty", "vx_stats", "vx_render", "vx_render_wait",
    "vx_read_frame_rgba8", "vx_read_frame_rgba32f", "vx_frame_device_ptr", "vx_raycast", "vx_raycast_device", "vx_raycast_wait",
    "vx_debug_cast", "vx_frame_stats", "vx_set_option", "vx_launch_count", "vx_build_info",
    "vx_shard_bytes", "vx_pack_shard", "vx_unpack_shard", "vx_set_streams", "vx_stream",
    "vx_frame_ipc_handle", "vx_open_peer_frame", "vx_close_peer_frame", "vx_render_read_rgba8",
    "vx_sync_ipc_handle", "vx_open_peer_sync", "vx_close_peer_sync", "vx_frame_signal", "vx_frame_wait", "vx_frame_gate",
    "vx_frame_sync_errors", "vx_frame8_ipc_handle", "vx_open_peer_frame8",
    "vx_serialize_chunks_esvo", "vx_serialize_chunks_result", "vx_svo_write_device",
    "vx_svo_scatter_errors", "vx_frame_flags_reset", "vx_read_hit_records",
    "vx_group_create", "vx_group_destroy", "vx_group_last_error", "vx_group_size", "vx_group_ctx", "vx_group_set_materials",
    "vx_group_set_textures", "vx_group_set_option", "vx_group_svo_host_mirror", "vx_group_svo_set_hot_range", "vx_group_svo_commit",
    "vx_group_stats", "vx_group_render", "vx_group_wait", "vx_group_read_frame_rgba8", "vx_group_read_frame_rgba32f",
    "vx_group_host_frame", "vx_group_render_read_rgba8", "vx_group_raycast", "vx_probe_read_bandwidth",
    "vx_render_read_rgba8_begin", "vx_render_read_rgba8_end",
]
VX_SHARD_ROWS = 0x80000000

_lib = None
_host = None


def _load(name):
    path = os.path.join(_PKG, name)
    if not os.path.exists(path):
        raise NativeMissing(f"{path} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                            "There is no CPU fallback.")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def lib():
    """libvoxelrt.so with argtypes set. Raises NativeMissing if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    L = _load("libvoxelrt.so")
    P = C.c_void_p
    L.vx_create.argtypes = [C.POINTER(VxConfig), C.POINTER(P)]; L.vx_create.restype = C.c_int
    L.vx_destroy.argtypes = [P]; L.vx_destroy.restype = None
    L.vx_last_error.argtypes = [P]; L.vx_last_error.restype = C.c_char_p
    L.vx_set_materials.argtypes = [P, C.POINTER(VxMaterial), C.c_uint32]; L.vx_set_materials.restype = C.c_int
    L.vx_set_textures.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]; L.vx_set_textures.restype = C.c_int
    L.vx_svo_host_mirror.argtypes = [P]; L.vx_svo_host_mirror.restype = P
    L.vx_svo_commit.argtypes = [P, C.c_float, C.POINTER(VxRange), C.c_uint32, C.c_uint64, C.c_uint32]; L.vx_svo_commit.restype = C.c_int
    L.vx_svo_set_hot_range.argtypes = [P, C.c_uint64, C.c_uint64]; L.vx_svo_set_hot_range.restype = C.c_int
    L.vx_svo_commit_packed_device.argtypes = [P, P, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32]; L.vx_svo_commit_packed_device.restype = C.c_int
    L.vx_svo_pack_dirty.argtypes = [P, C.POINTER(VxRange), C.c_uint32, P, C.c_uint64]; L.vx_svo_pack_dirty.restype = C.c_int64
    L.vx_stats.argtypes = [P, C.POINTER(VxStats)]; L.vx_stats.restype = C.c_int
    L.vx_render.argtypes = [P, C.POINTER(VxRenderParams), C.c_uint32, C.c_uint32, C.POINTER(VxShard), P]; L.vx_render.restype = C.c_int
    L.vx_render_read_rgba8.argtypes = [P, C.POINTER(VxRenderParams), C.c_uint32, C.c_uint32, C.POINTER(VxShard), P, C.c_uint32]
    L.vx_render_read_rgba8.restype = C.c_int
    L.vx_render_wait.argtypes = [P]; L.vx_render_wait.restype = C.c_int
    L.vx_read_frame_rgba8.argtypes = [P, P]; L.vx_read_frame_rgba8.restype = C.c_int
    L.vx_read_frame_rgba32f.argtypes = [P, P]; L.vx_read_frame_rgba32f.restype = C.c_int
    L.vx_frame_device_ptr.argtypes = [P, C.POINTER(P), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]; L.vx_frame_device_ptr.restype = C.c_int
    L.vx_raycast.argtypes = [P, P, C.c_uint64, P]; L.vx_raycast.restype = C.c_int
    L.vx_raycast_device.argtypes = [P, P, C.c_uint64, P]; L.vx_raycast_device.restype = C.c_int
    L.vx_raycast_wait.argtypes = [P]; L.vx_raycast_wait.restype = C.c_int
    L.vx_debug_cast.argtypes = [P, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_float, C.c_uint32, C.POINTER(VxOctreeResult),
                                C.POINTER(VxDebugFrame), C.c_uint32, C.POINTER(C.c_uint32)]
    L.vx_debug_cast.restype = C.c_int
    L.vx_frame_stats.argtypes = [P, C.c_int, C.POINTER(VxFrameStats)]; L.vx_frame_stats.restype = C.c_int
    L.vx_set_option.argtypes = [P, C.c_uint32, C.c_uint64]; L.vx_set_option.restype = C.c_int
    L.vx_launch_count.argtypes = [P]; L.vx_launch_count.restype = C.c_uint64
    L.vx_build_info.argtypes = []; L.vx_build_info.restype = C.c_char_p
    L.vx_shard_bytes.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(VxShard)]; L.vx_shard_bytes.restype = C.c_uint64
    L.vx_pack_shard.argtypes = [P, C.POINTER(VxShard), P]; L.vx_pack_shard.restype = C.c_int
    L.vx_unpack_shard.argtypes = [P, C.POINTER(VxShard), P]; L.vx_unpack_shard.restype = C.c_int
    L.vx_set_streams.argtypes = [P, P, P, P]; L.vx_set_streams.restype = C.c_int
    L.vx_stream.argtypes = [P, C.c_int, C.POINTER(P)]; L.vx_stream.restype = C.c_int
    L.vx_frame_ipc_handle.argtypes = [P, P]; L.vx_frame_ipc_handle.restype = C.c_int
    L.vx_open_peer_frame.argtypes = [P, P]; L.vx_open_peer_frame.restype = C.c_int
    L.vx_close_peer_frame.argtypes = [P]; L.vx_close_peer_frame.restype = C.c_int
    L.vx_serialize_chunks_esvo.argtypes = [P, P, C.c_uint32, P, P, P, C.c_uint64, C.POINTER(C.c_uint64)]; L.vx_serialize_chunks_esvo.restype = C.c_int
    L.vx_serialize_chunks_result.argtypes = [P, C.POINTER(P), C.POINTER(C.c_float)]; L.vx_serialize_chunks_result.restype = C.c_int
    L.vx_svo_write_device.argtypes = [P, C.c_uint64, P, C.c_uint64]; L.vx_svo_write_device.restype = C.c_int
    L.vx_frame8_ipc_handle.argtypes = [P, P]; L.vx_frame8_ipc_handle.restype = C.c_int
    L.vx_open_peer_frame8.argtypes = [P, P]; L.vx_open_peer_frame8.restype = C.c_int
    L.vx_sync_ipc_handle.argtypes = [P, P]; L.vx_sync_ipc_handle.restype = C.c_int
    L.vx_open_peer_sync.argtypes = [P, P]; L.vx_open_peer_sync.restype = C.c_int
    L.vx_close_peer_sync.argtypes = [P]; L.vx_close_peer_sync.restype = C.c_int
    L.vx_frame_signal.argtypes = [P, C.c_uint32, C.c_uint32]; L.vx_frame_signal.restype = C.c_int
    L.vx_frame_wait.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32]; L.vx_frame_wait.restype = C.c_int
    L.vx_frame_gate.argtypes = [P, C.c_uint32, C.c_uint32]; L.vx_frame_gate.restype = C.c_int
    L.vx_frame_sync_errors.argtypes = [P, C.POINTER(C.c_uint32)]; L.vx_frame_sync_errors.restype = C.c_int
    try:
        L.vx_svo_scatter_errors.argtypes = [P, C.POINTER(C.c_uint32)]; L.vx_svo_scatter_errors.restype = C.c_int
        L.vx_frame_flags_reset.argtypes = [P]; L.vx_frame_flags_reset.restype = C.c_int
        L.vx_read_hit_records.argtypes = [P, P]; L.vx_read_hit_records.restype = C.c_int
        u32, u64, i32p = C.c_uint32, C.c_uint64, C.POINTER(C.c_int)
        L.vx_group_create.argtypes = [C.POINTER(VxConfig), i32p, u32, C.POINTER(P)]; L.vx_group_create.restype = C.c_int
        L.vx_group_destroy.argtypes = [P]; L.vx_group_destroy.restype = None
        L.vx_group_last_error.argtypes = [P]; L.vx_group_last_error.restype = C.c_char_p
        L.vx_group_size.argtypes = [P]; L.vx_group_size.restype = u32
        L.vx_group_ctx.argtypes = [P, u32]; L.vx_group_ctx.restype = P
        L.vx_group_set_materials.argtypes = [P, C.POINTER(VxMaterial), u32]; L.vx_group_set_materials.restype = C.c_int
        L.vx_group_set_textures.argtypes = [P, P, u32, u32, u32, u32]; L.vx_group_set_textures.restype = C.c_int
        L.vx_group_set_option.argtypes = [P, u32, u64]; L.vx_group_set_option.restype = C.c_int
        L.vx_group_svo_host_mirror.argtypes = [P]; L.vx_group_svo_host_mirror.restype = P
        L.vx_group_svo_set_hot_range.argtypes = [P, u64, u64]; L.vx_group_svo_set_hot_range.restype = C.c_int
        L.vx_group_svo_commit.argtypes = [P, C.c_float, C.POINTER(VxRange), u32, u64, u32]; L.vx_group_svo_commit.restype = C.c_int
        L.vx_group_stats.argtypes = [P, C.POINTER(VxStats)]; L.vx_group_stats.restype = C.c_int
        L.vx_group_render.argtypes = [P, C.POINTER(VxRenderParams), u32, u32]; L.vx_group_render.restype = C.c_int
        L.vx_group_wait.argtypes = [P]; L.vx_group_wait.restype = C.c_int
        L.vx_group_read_frame_rgba8.argtypes = [P, P]; L.vx_group_read_frame_rgba8.restype = C.c_int
        L.vx_group_read_frame_rgba32f.argtypes = [P, P]; L.vx_group_read_frame_rgba32f.restype = C.c_int
        L.vx_group_host_frame.argtypes = [P, u64]; L.vx_group_host_frame.restype = P
        L.vx_group_render_read_rgba8.argtypes = [P, C.POINTER(VxRenderParams), u32, u32, P, u32]; L.vx_group_render_read_rgba8.restype = C.c_int
        L.vx_group_raycast.argtypes = [P, P, u64, P]; L.vx_group_raycast.restype = C.c_int
        L.vx_probe_read_bandwidth.argtypes = [P, u64, u32, C.POINTER(C.c_float)]; L.vx_probe_read_bandwidth.restype = C.c_int
        L.vx_render_read_rgba8_begin.argtypes = [P, C.POINTER(VxRenderParams), u32, u32, C.POINTER(VxShard), P, u32]; L.vx_render_read_rgba8_begin.restype = C.c_int
        L.vx_render_read_rgba8_end.argtypes = [P]; L.vx_render_read_rgba8_end.restype = C.c_int
    except AttributeError:
        if not os.environ.get("VOXELRT_AB_VARIANT"):   # only tools/ab_kernels.py may load an older build of the library
            raise
    _lib = L
    return L


def host():
    """libvoxelrs_host.so (C++ host mirror) with argtypes set. With VOXELRS_WORLD_ONLY=1 in the environment (bench.py's reference
    arm) the world-producer half alone, libvoxelrs_world.so, is loaded instead: World / registries work, nothing that needs the GPU
    library exists, and libvoxelrt.so is never mapped into the process."""
    global _host
    if _host is not None:
        return _host
    world_only = os.environ.get("VOXELRS_WORLD_ONLY") == "1"
    if not world_only:
        lib()  # dependency, loaded RTLD_GLOBAL first
    H = _load("libvoxelrs_world.so" if world_only else "libvoxelrs_host.so")
    P, u8p, u32, u64, i32, f = C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_float
    sig = {
        "vxh_last_error": ([], C.c_char_p),
        "vxh_world_new": ([u32, i32, i32, i32, u32, C.c_int], P),
        "vxh_world_free": ([P], None),
        "vxh_world_set_format": ([P, C.c_int], None),
        "vxh_world_format": ([P], C.c_int),
        "vxh_world_header_bytes": ([P], u64),
        "vxh_world_generate": ([P, i32, i32, C.c_int], u64),
        "vxh_world_height_at": ([P, i32, i32], i32),
        "vxh_world_set_terrain": ([P, C.c_int], None),
        "vxh_kat_worldgen_noise": ([u32, C.c_float, C.c_int, C.c_double, C.c_double], C.c_double),
        "vxh_kat_spline": ([C.POINTER(C.c_float), C.c_int, C.c_double], C.c_double),
        "vxh_world_chunk_count": ([P], u64),
        "vxh_world_set_leaf_blocks": ([P, u32, u32, u32, u64, P, u32, C.c_uint8, C.c_int], C.c_int),
        "vxh_world_set_leaf_dense": ([P, u32, u32, u32, u64, P, C.c_uint8], C.c_int),
        "vxh_world_edit_block": ([P, i32, i32, i32, u32], C.c_int),
        "vxh_world_serialize": ([P], None),
        "vxh_world_depth": ([P], u32),
        "vxh_world_size_bytes": ([P], u64),
        "vxh_world_write_to": ([P, u8p], u64),
        "vxh_world_write_changes_to": ([P, u8p, u64, C.c_int], C.c_int),
        "vxh_world_dirty_ranges": ([P, P, u32], u32),
        "vxh_world_mark_all_dirty": ([P], None),
        "vxh_world_root_range": ([P, C.POINTER(u64), C.POINTER(u64)], None),
        "vxh_world_root_info": ([P, C.POINTER(u64), P], None),
        "vxh_world_cnv_block_pos": ([P, P, P], None),
        "vxh_world_cnv_svo_pos": ([P, P, P], None),
        "vxh_world_cnv_chunk_pos": ([P, i32, i32, i32, P], C.c_int),
        "vxh_calculate_lod": ([i32, i32, i32, i32, i32, i32], C.c_uint8),
        "vxh_kat_block_octree": ([P, u32, C.c_uint8, C.c_int, C.c_uint8, P, u64, P], u64),
        "vxh_kat_csvo_octant": ([P, u32, C.c_uint8, C.c_int, C.c_uint8, P, u64, P, u32, C.POINTER(u32)], u64),
        "vxh_csvo_dense_equals_generic": ([P, C.c_uint8], C.c_int),
        "vxh_serialize_dense_batch": ([P, u32, P, C.c_int], u64),
        "vxh_world_chunk_list": ([P, P, u64], u64),
        "vxh_world_fill_chunk": ([P, i32, i32, i32, P], C.c_int),
        "vxh_world_chunk_range": ([P, i32, i32, i32, C.POINTER(VxRange)], C.c_int),
        "vxh_world_csvo_root_offset": ([P], u64),
        "vxh_world_range_bytes": ([P, P, u64], u64),
        "vxh_serialize_dense": ([P, C.c_uint8, P, u64, P], u64),
        "vxh_serialize_filled": ([P, C.c_uint8, P, u64, P], u64),
        "vxh_esvo32_new": ([], P), "vxh_esvo32_free": ([P], None),
        "vxh_esvo32_set_leaf": ([P, u32, u32, u32, u32, C.c_int, P], None),
        "vxh_esvo32_move_leaf": ([P, u32, u32, u32, u32, u32, P, C.POINTER(u32)], C.c_int),
        "vxh_esvo32_remove_leaf": ([P, u32, u32, C.POINTER(u32)], C.c_int),
        "vxh_esvo32_serialize": ([P], None),
        "vxh_esvo32_root_info": ([P, C.POINTER(u64), P], None),
        "vxh_esvo32_bytes": ([P, P, u64], u64),
        "vxh_esvo32_ranges": ([P, C.c_int, P, u32], u32),
        "vxh_esvo32_range_of": ([P, u64, C.POINTER(VxRange)], C.c_int),
        "vxh_esvo32_clear_updated": ([P], None),
        "vxh_esvo32_write_to": ([P, P], u64),
        "vxh_esvo32_write_changes_to": ([P, P, u64, C.c_int], C.c_int),
        "vxh_rangebuf_new": ([], P), "vxh_rangebuf_free": ([P], None),
        "vxh_rangebuf_insert": ([P, u64, P, u64], u64), "vxh_rangebuf_remove": ([P, u64], None),
        "vxh_rangebuf_bytes": ([P, P, u64], u64), "vxh_rangebuf_ranges": ([P, C.c_int, P, u32], u32),
        "vxh_rangebuf_with_capacity": ([u64], P), "vxh_rangebuf_ids": ([P, P, u32], u32), "vxh_merge_ranges": ([P, u32], u32),
        "vxh_kat_shift_chunks": ([P, i32, i32, i32, u32, P, P, u32], u32), "vxh_esvo32_get_leaf": ([P, u32, u32, u32], C.c_int64),
        "vxh_world_set_center": ([P, i32, i32, i32], C.c_int),
        "vxh_world_load_chunk": ([P, i32, i32, i32], C.c_int), "vxh_world_remove_chunk": ([P, i32, i32, i32], None),
        "vxh_chunkloader_new": ([u32, i32, i32], P), "vxh_chunkloader_free": ([P], None), "vxh_chunkloader_set_radius": ([P, u32], None),
        "vxh_chunkloader_is_loaded": ([P, i32, i32, i32], C.c_int), "vxh_chunkloader_add_loaded": ([P, i32, i32, i32, u32], None),
        "vxh_chunkloader_loaded_count": ([P], u64), "vxh_chunkloader_update": ([P, f, f, f, P, u64], u64),
        "vxh_octree_new": ([], P), "vxh_octree_free": ([P], None),
        "vxh_octree_set_leaf": ([P, u32, u32, u32, u32, P], None), "vxh_octree_move_leaf": ([P, u32, u32, u32, u32, u32, P], None),
        "vxh_octree_remove_leaf": ([P, u32, u32, u32, P], None), "vxh_octree_remove_leaf_by_id": ([P, u32, u32], C.c_int64),
        "vxh_octree_get_leaf": ([P, u32, u32, u32], C.c_int64), "vxh_octree_compact": ([P], None),
        "vxh_octree_construct": ([P, u32, P, u32], None), "vxh_octree_dump": ([P, P, P, u32, P, u32], None),
        "vxh_picker_serialize": ([P, u32, P, u32, P, u64], u64),
        "vxh_picker_deserialize": ([P, u32, P, u32, P, u64, P, P], None),
        "vxh_registry_new": ([], P), "vxh_registry_free": ([P], None),
        "vxh_registry_add_texture": ([P, C.c_char_p, u32, u32, P], C.c_int),
        "vxh_registry_set_mip_levels": ([P, C.c_uint8], None),
        "vxh_registry_add_material": ([P, u32, f, f, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int], C.c_int),
        "vxh_registry_materials": ([P, P, u32], u32),
        "vxh_registry_textures": ([P, P, u64, P], u64),
        "vxh_look_to_rh_inverted": ([P, P, P, P], None),
        "vxh_svo_new": ([P, u64, u32, u32, u64, C.c_int, u32], P), "vxh_svo_free": ([P], None),
        "vxh_svo_ctx": ([P], P),
        "vxh_svo_update": ([P, P], C.c_int),
        "vxh_svo_stats": ([P, P], None),
        "vxh_svo_render": ([P, C.POINTER(VxhRenderParams), u32, u32, C.POINTER(VxShard)], C.c_int),
        "vxh_worldsvo_render": ([P, P, C.POINTER(VxhRenderParams), u32, u32, C.POINTER(VxShard)], C.c_int),
        "vxh_svo_raycast": ([P, P, u32, P, u32, P, P], C.c_int),
        "vxh_worldsvo_raycast": ([P, P, P, u32, P, u32, P, P], C.c_int),
    }
    for name, (args, res) in sig.items():
        if world_only and (name.startswith("vxh_svo_") or name.startswith("vxh_worldsvo_")):
            continue
        fn = getattr(H, name)
        fn.argtypes = args
        fn.restype = res
    _host = H
    return H


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class VxError(RuntimeError):
    pass


# ------------------------------------------------------------------- world --

class World:
    """systems::worldsvo::Svo's CPU half: Esvo of SerializedChunks + SVO coordinate space (+ synthetic terrain)."""

    def __init__(self, radius=0, center=(0, 0, 0), seed=1, no_lod=False, fmt=FORMAT_ESVO, terrain="standin"):
        """terrain: "reference" = the reference's generator (noise 0.8.2 Perlin + splines, gamelogic/worldgen.rs), "standin" = hash-gradient
        noise of the same shape (the world the round-1 profiles were taken on)."""
        self.h = host().vxh_world_new(radius, center[0], center[1], center[2], seed, int(no_lod))
        host().vxh_world_set_terrain(self.h, {"standin": 0, "reference": 1}[terrain])
        self.radius, self.center, self.fmt = radius, tuple(center), fmt
        host().vxh_world_set_format(self.h, fmt)

    @property
    def header_bytes(self):
        """Bytes in front of the RangeBuffer image inside the GPU buffer (24 ESVO, 8 CSVO)."""
        return host().vxh_world_header_bytes(self.h)

    @property
    def svo_flags(self):
        """VxConfig.flags bits a graphics::Svo for this world needs."""
        return VX_FLAG_SVO_CSVO if self.fmt == FORMAT_CSVO else 0

    def __del__(self):
        if getattr(self, "h", None):
            host().vxh_world_free(self.h)
            self.h = None

    def generate(self, y0=0, y1=8, threads=0):
        return host().vxh_world_generate(self.h, y0, y1, threads or (os.cpu_count() or 1))

    def set_leaf_blocks(self, svo_pos, blocks, uid=1, lod=5, compact=False):
        """blocks: iterable of (x, y, z, id) set through Chunk::set_block on pooled storage (expand_to(5))."""
        arr = np.ascontiguousarray(np.array(list(blocks), dtype=np.uint32).reshape(-1, 4))
        rc = host().vxh_world_set_leaf_blocks(self.h, svo_pos[0], svo_pos[1], svo_pos[2], uid, _ptr(arr), len(arr), lod, int(compact))
        if rc:
            raise VxError(host().vxh_last_error().decode())

    def set_leaf_dense(self, svo_pos, blocks32, uid=1, lod=5):
        arr = np.ascontiguousarray(blocks32, dtype=np.uint32).reshape(-1)
        assert arr.size == 32 ** 3
        rc = host().vxh_world_set_leaf_dense(self.h, svo_pos[0], svo_pos[1], svo_pos[2], uid, _ptr(arr), lod)
        if rc:
            raise VxError(host().vxh_last_error().decode())

    def edit_block(self, wx, wy, wz, block_id):
        host().vxh_world_edit_block(self.h, wx, wy, wz, block_id)

    def serialize(self):
        host().vxh_world_serialize(self.h)

    def set_center(self, center):
        """The player entered another chunk (systems::worldsvo::Svo::update, worldsvo.rs:133-137): the SVO window is re-centred and
        every loaded chunk shifted to its new place in SVO space (shift_chunks, :158-196); chunks outside the window are dropped.
        Chunk records stay where they are in the buffer. Returns True if the centre changed."""
        changed = bool(host().vxh_world_set_center(self.h, int(center[0]), int(center[1]), int(center[2])))
        self.center = tuple(int(v) for v in center)
        return changed

    def load_chunk(self, c):
        """(Re)generates and inserts one chunk with the LOD rule (a storage / generator result arriving, worldsvo.rs:90-99)."""
        return bool(host().vxh_world_load_chunk(self.h, int(c[0]), int(c[1]), int(c[2])))

    def remove_chunk(self, c):
        host().vxh_world_remove_chunk(self.h, int(c[0]), int(c[1]), int(c[2]))

    def write_changes_to(self, gpu_buffer, reset=True):
        """graphics::Svo::update's host half on a caller-held image of the GPU buffer (svo.rs:171-189): octree_scale at byte 0, then
        Esvo::write_changes_to(dst = buffer + 4) = preamble + the dirty ranges only. Returns False where the reference's assert fires."""
        gpu_buffer[:4] = np.frombuffer(np.float32(2.0 ** -self.depth).tobytes(), dtype=np.uint8)
        return host().vxh_world_write_changes_to(self.h, C.c_void_p(gpu_buffer.ctypes.data + 4), len(gpu_buffer) - 4, int(reset)) == 0

    @property
    def depth(self):
        return host().vxh_world_depth(self.h)

    @property
    def size_bytes(self):
        return host().vxh_world_size_bytes(self.h)

    @property
    def chunk_count(self):
        return host().vxh_world_chunk_count(self.h)

    def height_at(self, x, z):
        return host().vxh_world_height_at(self.h, x, z)

    def chunks(self):
        """World chunk coordinates of the loaded chunks."""
        n = host().vxh_world_chunk_list(self.h, None, 0)
        a = np.zeros((n, 3), dtype=np.int32)
        host().vxh_world_chunk_list(self.h, _ptr(a), n)
        return a

    def chunk_blocks(self, c):
        """(dense 32^3 uint32 block array, lod) the chunk was serialized from (terrain + edits)."""
        b = np.zeros(32768, dtype=np.uint32)
        lod = host().vxh_world_fill_chunk(self.h, int(c[0]), int(c[1]), int(c[2]), _ptr(b))
        return b, lod

    def chunk_range(self, c):
        r = VxRange()
        ok = host().vxh_world_chunk_range(self.h, int(c[0]), int(c[1]), int(c[2]), C.byref(r))
        return (r.offset, r.length) if ok else None

    def range_bytes(self):
        n = host().vxh_world_range_bytes(self.h, None, 0)
        out = np.zeros(n, dtype=np.uint8)
        host().vxh_world_range_bytes(self.h, _ptr(out), n)
        return out

    def dirty_ranges(self):
        n = host().vxh_world_dirty_ranges(self.h, None, 0)
        arr = (VxRange * max(n, 1))()
        host().vxh_world_dirty_ranges(self.h, arr, n)
        return [(arr[i].offset, arr[i].length) for i in range(n)]

    def mark_all_dirty(self):
        host().vxh_world_mark_all_dirty(self.h)

    def root_range(self):
        a, b = C.c_uint64(), C.c_uint64()
        host().vxh_world_root_range(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def root_info(self):
        off = C.c_uint64()
        m = (C.c_uint8 * 3)()
        host().vxh_world_root_info(self.h, C.byref(off), m)
        return off.value, m[0], m[1], m[2]

    def gpu_buffer(self):
        """Bytes exactly as graphics::Svo::update lays them out: f32 2^-depth, preamble, RangeBuffer (svo.rs:173-181)."""
        hb = self.header_bytes
        buf = np.zeros(hb + self.size_bytes + (4 if self.fmt == FORMAT_CSVO else 0), dtype=np.uint8)   # CSVO: one spare word for read_uint at the end
        buf[:4] = np.frombuffer(np.float32(2.0 ** -self.depth).tobytes(), dtype=np.uint8)
        n = host().vxh_world_write_to(self.h, C.c_void_p(buf.ctypes.data + 4))
        assert n == hb - 4 + self.size_bytes or n == 0
        return buf

    def cnv_block_pos(self, p):
        a, o = np.array(p, dtype=np.float32), np.zeros(3, dtype=np.float32)
        host().vxh_world_cnv_block_pos(self.h, _ptr(a), _ptr(o))
        return o

    def cnv_svo_pos(self, p):
        a, o = np.array(p, dtype=np.float32), np.zeros(3, dtype=np.float32)
        host().vxh_world_cnv_svo_pos(self.h, _ptr(a), _ptr(o))
        return o

    def cnv_chunk_pos(self, c):
        o = np.zeros(3, dtype=np.uint32)
        ok = host().vxh_world_cnv_chunk_pos(self.h, c[0], c[1], c[2], _ptr(o))
        return tuple(int(v) for v in o) if ok else None


class ChunkLoader:
    """systems::chunkloader::ChunkLoader (src/systems/chunkloader.rs): update(pos) -> [(kind, (x, y, z), lod)] with kind "load" /
    "unload" / "lod", nearest chunks first."""
    KINDS = ("load", "unload", "lod")

    def __init__(self, radius, start_y=0, end_y=8):
        self.h = host().vxh_chunkloader_new(radius, start_y, end_y)
        if not self.h:
            raise VxError("ChunkLoader: start_y < end_y required")   # the reference's assert!, chunkloader.rs:35
        self.radius = radius

    def __del__(self):
        if getattr(self, "h", None):
            host().vxh_chunkloader_free(self.h)
            self.h = None

    def update(self, pos):
        r = self.radius
        cap = (2 * r + 1) ** 2 * 64 + 64
        out = np.zeros((cap, 5), dtype=np.int32)
        n = host().vxh_chunkloader_update(self.h, float(pos[0]), float(pos[1]), float(pos[2]), _ptr(out), cap)
        assert n <= cap
        return [(self.KINDS[k], (int(x), int(y), int(z)), int(lod)) for k, x, y, z, lod in out[:n]]

    def is_loaded(self, c):
        return bool(host().vxh_chunkloader_is_loaded(self.h, int(c[0]), int(c[1]), int(c[2])))

    @property
    def loaded_count(self):
        return host().vxh_chunkloader_loaded_count(self.h)


def follow(world, loader, cam_pos):
    """One streaming step of gamelogic::World::update (src/gamelogic/world.rs:116-210) without the job system: the chunk loader's events
    for the new camera position are applied to the world (load = generate + serialize with the LOD rule, unload = remove, LOD change =
    serialize again) and the SVO window is re-centred on the camera's chunk. Returns the events. Call world.serialize() afterwards."""
    events = loader.update(cam_pos)
    center = tuple(int(v) >> 5 for v in cam_pos)            # ChunkPos::from(camera.position), chunk.rs:150-152,175-178
    world.set_center(center)
    for kind, c, lod in events:
        if kind == "unload":
            world.remove_chunk(c)
        else:
            world.load_chunk(c)
    return events


# ---------------------------------------------------------------- registry --

class Registry:
    """graphics::svo_registry::VoxelRegistry (src/graphics/svo_registry.rs:99-165)."""

    def __init__(self, mip_levels=6):
        self.h = host().vxh_registry_new()
        host().vxh_registry_set_mip_levels(self.h, mip_levels)

    def __del__(self):
        if getattr(self, "h", None):
            host().vxh_registry_free(self.h)
            self.h = None

    def add_texture(self, name, rgba_top_down):
        img = np.ascontiguousarray(rgba_top_down, dtype=np.uint8)
        hgt, wid = img.shape[0], img.shape[1]
        if host().vxh_registry_add_texture(self.h, name.encode(), wid, hgt, _ptr(img)):
            raise VxError(host().vxh_last_error().decode())
        return self

    def add_material(self, block, specular=(0.0, 0.0), top=None, side=None, bottom=None, all_sides=None, with_normals=False):
        if all_sides is not None:
            top = side = bottom = all_sides
        enc = lambda s: s.encode() if s is not None else None
        host().vxh_registry_add_material(self.h, block, specular[0], specular[1], enc(top), enc(side), enc(bottom), int(with_normals))
        return self

    def materials(self):
        n = host().vxh_registry_materials(self.h, None, 0)
        arr = (VxMaterial * max(n, 1))()
        host().vxh_registry_materials(self.h, arr, n)
        return np.frombuffer(bytes(arr), dtype=np.dtype([("specular_pow", "<f4"), ("specular_strength", "<f4"), ("tex", "<i4", 6)]))[:n].copy()

    def textures(self):
        dims = (C.c_uint32 * 4)()
        n = host().vxh_registry_textures(self.h, None, 0, dims)
        buf = np.zeros(n, dtype=np.uint8)
        host().vxh_registry_textures(self.h, _ptr(buf), n, dims)
        w, h, layers, mips = dims
        return buf.reshape(layers, h, w, 4), mips


# the 25 textures + 13 materials of src/gamelogic/content.rs:20-60, in registration order
CONTENT_TEXTURES = [
    ("dirt", "dirt"), ("dirt_normal", "dirt_n"), ("grass_side", "grass_side"), ("grass_side_normal", "grass_side_n"),
    ("grass_top", "grass_top"), ("grass_top_normal", "grass_top_n"), ("stone", "stone"), ("stone_normal", "stone_n"),
    ("stone_bricks", "stone_bricks"), ("stone_bricks_normal", "stone_bricks_n"), ("glass", "glass"), ("gravel", "gravel"),
    ("gravel_normal", "gravel_n"), ("sand", "sand"), ("sand_normal", "sand_n"), ("water", "water"), ("oak_log", "oak_log"),
    ("oak_log_normal", "oak_log_n"), ("oak_log_top", "oak_log_top"), ("oak_log_top_normal", "oak_log_top_n"),
    ("oak_leaves", "oak_leaves"), ("oak_planks", "oak_planks"), ("oak_planks_normal", "oak_planks_n"),
    ("cobblestone", "cobblestone"), ("cobblestone_normal", "cobblestone_n"),
]


def content_registry(atlas):
    """blocks::new_registry() of src/gamelogic/content.rs:20-60. atlas: {png stem: HxWx4 uint8, top-down}."""
    r = Registry(6)
    for name, stem in CONTENT_TEXTURES:
        r.add_texture(name, atlas[stem])
    r.add_material(0)
    r.add_material(1, (14.0, 0.4), top="grass_top", side="grass_side", bottom="dirt", with_normals=True)
    r.add_material(2, (14.0, 0.4), all_sides="dirt", with_normals=True)
    r.add_material(3, (70.0, 0.4), all_sides="stone", with_normals=True)
    r.add_material(4, (70.0, 0.4), all_sides="stone_bricks", with_normals=True)
    r.add_material(5, (70.0, 0.4), all_sides="glass")
    r.add_material(6, (70.0, 0.4), all_sides="gravel", with_normals=True)
    r.add_material(7, (70.0, 0.4), all_sides="sand", with_normals=True)
    r.add_material(8, (70.0, 0.4), all_sides="water")
    r.add_material(9, (70.0, 0.4), side="oak_log", top="oak_log_top", bottom="oak_log_top", with_normals=True)
    r.add_material(10, (70.0, 0.4), all_sides="oak_leaves")
    r.add_material(11, (70.0, 0.4), all_sides="oak_planks", with_normals=True)
    r.add_material(12, (70.0, 0.4), all_sides="cobblestone", with_normals=True)
    return r


def load_atlas(path=None):
    """The reference's 25 block textures as committed under tests/golden/atlas.npz (see make_fixtures.py)."""
    path = path or os.path.join(ROOT, "tests", "golden", "atlas.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


# --------------------------------------------------------------------- Svo --

def render_params(cam_pos, cam_fwd, cam_up=(0, 1, 0), fov_y_deg=72.0, aspect=1.0, ambient=0.3, light_dir=None, selected_voxel=None,
                  render_shadows=True, shadow_distance=500.0):
    """RenderParams as gamelogic::World::render builds them (src/gamelogic/world.rs:269-283)."""
    p = VxhRenderParams()
    if light_dir is None:
        l = np.float32(-1.0) / np.sqrt(np.float32(3.0))
        light_dir = (l, l, l)
    p.ambient_intensity = ambient
    p.light_dir = _f3(light_dir); p.cam_pos = _f3(cam_pos); p.cam_fwd = _f3(cam_fwd); p.cam_up = _f3(cam_up)
    p.fov_y_rad = float(np.float32(np.deg2rad(np.float32(fov_y_deg))))
    p.aspect_ratio = aspect
    p.has_selected_voxel = int(selected_voxel is not None)
    p.selected_voxel = _f3(selected_voxel if selected_voxel is not None else (0, 0, 0))
    p.render_shadows = int(render_shadows)
    p.shadow_distance = shadow_distance
    return p


def to_vx_render_params(p):
    """What graphics::Svo::render uploads as uniforms (svo.rs:197-215) — used to feed the oracle the same inputs."""
    q = VxRenderParams()
    eye, d, up = np.array(p.cam_pos, np.float32), np.array(p.cam_fwd, np.float32), np.array(p.cam_up, np.float32)
    out = np.zeros(16, np.float32)
    host().vxh_look_to_rh_inverted(_ptr(eye), _ptr(d), _ptr(up), _ptr(out))
    q.view = (C.c_float * 16)(*out)
    q.fov_y_rad, q.aspect_ratio, q.ambient_intensity = p.fov_y_rad, p.aspect_ratio, p.ambient_intensity
    q.light_dir = p.light_dir; q.cam_pos = p.cam_pos
    nan = float("nan")
    q.highlight_pos = _f3(p.selected_voxel) if p.has_selected_voxel else _f3((nan, nan, nan))
    q.render_shadows = p.render_shadows
    q.shadow_distance = p.shadow_distance
    return q


class SvoGroup:
    """graphics::Svo spread over several GPUs of ONE process (vx_group_*, include/voxelrt.h): the reference engine's shape — a single
    process, one Svo — with the SVO replicated per device and frames cut into image-space shards. Straight over the C ABI."""

    def __init__(self, registry, devices, size_mb=10, max_width=1920, max_height=1080, max_rays=100, flags=0):
        cfg = VxConfig(0, flags, size_mb * 1000 * 1000, max_width, max_height, max_rays)
        devs = (C.c_int * len(devices))(*devices)
        g = C.c_void_p()
        rc = lib().vx_group_create(C.byref(cfg), devs, len(devices), C.byref(g))
        if rc:
            raise VxError(f"rc={rc}: {lib().vx_group_last_error(None).decode()}")
        self.g, self.devices, self.capacity = g, list(devices), size_mb * 1000 * 1000
        self.width = self.height = 0
        tex, mips = registry.textures()
        mats = registry.materials()
        arr = (VxMaterial * len(mats)).from_buffer_copy(mats.tobytes())
        self._check(lib().vx_group_set_textures(self.g, _ptr(np.ascontiguousarray(tex)), tex.shape[2], tex.shape[1], tex.shape[0], mips))
        self._check(lib().vx_group_set_materials(self.g, arr, len(mats)))

    def close(self):
        if getattr(self, "g", None):
            lib().vx_group_destroy(self.g)
            self.g = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise VxError(f"rc={rc}: {lib().vx_group_last_error(self.g).decode()}")

    def __len__(self):
        return lib().vx_group_size(self.g)

    def ctx(self, i):
        return C.c_void_p(lib().vx_group_ctx(self.g, i))

    def set_option(self, opt, value):
        self._check(lib().vx_group_set_option(self.g, opt, value))

    def update(self, world):
        """graphics::Svo::update (svo.rs:171-189) for every replica: the serializer's write_changes_to goes into device 0's pinned
        mirror, vx_group_svo_commit moves the dirty ranges (H2D to device 0 -> NCCL broadcast -> scatter kernels)."""
        hb = world.header_bytes
        mirror = np.ctypeslib.as_array((C.c_uint8 * self.capacity).from_address(lib().vx_group_svo_host_mirror(self.g)))
        ranges = world.dirty_ranges()
        if not world.write_changes_to(mirror):
            raise VxError("dst is not large enough (esvo.rs:328-331)")
        arr = (VxRange * max(len(ranges), 1))(*[VxRange(o, l) for o, l in ranges])
        off, ln = world.root_range()
        lib().vx_group_svo_set_hot_range(self.g, off, ln)
        self._check(lib().vx_group_svo_commit(self.g, float(np.float32(2.0 ** -world.depth)), arr, len(ranges), world.size_bytes, world.depth))
        return hb

    def render_raw(self, vx_params, width, height):
        self._check(lib().vx_group_render(self.g, C.byref(vx_params), width, height))
        self.width, self.height = width, height

    def wait(self):
        self._check(lib().vx_group_wait(self.g))

    def read_rgba32f(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._check(lib().vx_group_read_frame_rgba32f(self.g, _ptr(out)))
        return out

    def read_rgba8(self):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._check(lib().vx_group_read_frame_rgba8(self.g, _ptr(out)))
        return out

    def host_frame(self, width, height):
        """(height, width, 4) uint8 view of the group's pinned portable host frame."""
        p = lib().vx_group_host_frame(self.g, width * height * 4)
        if not p:
            raise VxError(lib().vx_group_last_error(self.g).decode())
        return np.ctypeslib.as_array((C.c_uint8 * (width * height * 4)).from_address(p)).reshape(height, width, 4)

    def render_read_rgba8(self, vx_params, width, height, out_ptr, bands=2):
        self._check(lib().vx_group_render_read_rgba8(self.g, C.byref(vx_params), width, height, C.c_void_p(out_ptr), bands))
        self.width, self.height = width, height

    def raycast_tasks(self, tasks):
        tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
        res = np.zeros(len(tasks), dtype=RESULT_DTYPE)
        self._check(lib().vx_group_raycast(self.g, _ptr(tasks), len(tasks), _ptr(res)))
        return res

    def raycast_ptr(self, tasks_ptr, n, results_ptr):
        self._check(lib().vx_group_raycast(self.g, C.c_void_p(tasks_ptr), n, C.c_void_p(results_ptr)))


class Svo:
    """graphics::Svo (src/graphics/svo.rs:56-255) through the C++ host mirror and the C ABI."""

    def __init__(self, registry, size_mb=10, max_width=1920, max_height=1080, max_rays=100, device=0, flags=0):
        self.h = host().vxh_svo_new(registry.h, size_mb, max_width, max_height, max_rays, device, flags)
        if not self.h:
            raise VxError(host().vxh_last_error().decode())
        self.ctx = C.c_void_p(host().vxh_svo_ctx(self.h))
        self.width = self.height = 0

    def close(self):
        if getattr(self, "h", None):
            host().vxh_svo_free(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise VxError(f"rc={rc}: {lib().vx_last_error(self.ctx).decode()} | {host().vxh_last_error().decode()}")

    def set_option(self, opt, value):
        self._check(lib().vx_set_option(self.ctx, opt, value))

    def update(self, world):
        self._check(host().vxh_svo_update(self.h, world.h))

    def get_stats(self):
        o = (C.c_uint64 * 3)()
        host().vxh_svo_stats(self.h, o)
        return {"used_bytes": o[0], "capacity_bytes": o[1], "depth": o[2]}

    def render(self, params, width, height, shard=None, world=None):
        """world given => systems::worldsvo::Svo::render (camera in world space, worldsvo.rs:397-409)."""
        sh = VxShard(*shard) if shard else None
        shp = C.byref(sh) if sh else None
        if world is not None:
            self._check(host().vxh_worldsvo_render(self.h, world.h, C.byref(params), width, height, shp))
        else:
            self._check(host().vxh_svo_render(self.h, C.byref(params), width, height, shp))
        self.width, self.height = width, height

    def wait(self):
        self._check(lib().vx_render_wait(self.ctx))

    def read_rgba32f(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._check(lib().vx_read_frame_rgba32f(self.ctx, _ptr(out)))
        return out

    def read_hit_records(self):
        """OctreeResult of every pixel's primary ray of the last render (VxHitRecord, row 0 = bottom)."""
        out = np.zeros((self.height, self.width), dtype=HIT_DTYPE)
        self._check(lib().vx_read_hit_records(self.ctx, _ptr(out)))
        return out

    def read_rgba8(self):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        self._check(lib().vx_read_frame_rgba8(self.ctx, _ptr(out)))
        return out

    def raycast(self, rays=(), aabbs=(), world=None):
        """rays: (pos3, dir3, max_dst); aabbs: (pos3, offset3, extents3). Returns (ray results [n,8], aabb results [m,6])."""
        r = np.ascontiguousarray(np.array(rays, dtype=np.float32).reshape(-1, 7))
        a = np.ascontiguousarray(np.array(aabbs, dtype=np.float32).reshape(-1, 9))
        ro, ao = np.zeros((len(r), 8), np.float32), np.zeros((len(a), 6), np.float32)
        if world is not None:
            self._check(host().vxh_worldsvo_raycast(self.h, world.h, _ptr(r), len(r), _ptr(a), len(a), _ptr(ro), _ptr(ao)))
        else:
            self._check(host().vxh_svo_raycast(self.h, _ptr(r), len(r), _ptr(a), len(a), _ptr(ro), _ptr(ao)))
        return ro, ao

    def raycast_tasks(self, tasks):
        """Raw vx_raycast on a TASK_DTYPE array; returns a RESULT_DTYPE array."""
        tasks = np.ascontiguousarray(tasks)
        assert tasks.dtype == TASK_DTYPE
        res = np.zeros(len(tasks), dtype=RESULT_DTYPE)
        self._check(lib().vx_raycast(self.ctx, _ptr(tasks), len(tasks), _ptr(res)))
        return res

    def debug_cast(self, pos, direction, max_dst, cast_translucent, frames_cap=100):
        res = VxOctreeResult()
        frames = (VxDebugFrame * frames_cap)()
        n = C.c_uint32()
        d = np.array(direction, dtype=np.float32)
        d = d / np.sqrt(np.float32((d * d).sum(dtype=np.float32)))   # svo_shader_tests.rs:246 dir.normalize()
        self._check(lib().vx_debug_cast(self.ctx, C.byref(_f3(pos)), C.byref(_f3(d)), max_dst, int(cast_translucent), C.byref(res),
                                        frames, frames_cap, C.byref(n)))
        return res, [frames[i] for i in range(min(n.value, frames_cap))], n.value

    # ---- raw C-ABI helpers used by bench.py and the multi-GPU path ----
    def render_raw(self, vx_params, width, height, shard=None, out=None):
        """vx_render with an already-built VxRenderParams (no host-mirror work in the call)."""
        sh = VxShard(*shard) if shard else None
        self._check(lib().vx_render(self.ctx, C.byref(vx_params), width, height, C.byref(sh) if sh else None,
                                    _ptr(out) if out is not None else None))
        self.width, self.height = width, height

    def render_read_rgba8(self, vx_params, width, height, out_ptr, bands=4, shard=None):
        """vx_render_read_rgba8 into host memory at out_ptr (pinned for overlap)."""
        sh = VxShard(*shard) if shard else None
        self._check(lib().vx_render_read_rgba8(self.ctx, C.byref(vx_params), width, height, C.byref(sh) if sh else None,
                                               C.c_void_p(out_ptr), bands))
        self.width, self.height = width, height

    def render_read_rgba8_begin(self, vx_params, width, height, out_ptr, bands=2, shard=None):
        sh = VxShard(*shard) if shard else None
        self._check(lib().vx_render_read_rgba8_begin(self.ctx, C.byref(vx_params), width, height, C.byref(sh) if sh else None,
                                                     C.c_void_p(out_ptr), bands))
        self.width, self.height = width, height

    def render_read_rgba8_end(self):
        self._check(lib().vx_render_read_rgba8_end(self.ctx))

    def commit(self, octree_scale, ranges, used_bytes, depth):
        """vx_svo_commit: ranges = [(offset, length)] relative to the RangeBuffer, bytes already in the host mirror."""
        arr = (VxRange * max(len(ranges), 1))(*[VxRange(o, l) for o, l in ranges])
        self._check(lib().vx_svo_commit(self.ctx, octree_scale, arr, len(ranges), used_bytes, depth))

    def host_mirror(self, nbytes):
        """numpy view of the first nbytes of the pinned host mirror of the GPU world buffer."""
        p = lib().vx_svo_host_mirror(self.ctx)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def pack_dirty(self, ranges, out):
        arr = (VxRange * max(len(ranges), 1))(*[VxRange(o, l) for o, l in ranges])
        n = lib().vx_svo_pack_dirty(self.ctx, arr, len(ranges), _ptr(out) if out is not None else None, out.nbytes if out is not None else 0)
        if n < 0:
            self._check(int(n))
        return int(n)

    def commit_packed_device(self, dev_ptr, n_ranges, payload_bytes, used_bytes, depth):
        self._check(lib().vx_svo_commit_packed_device(self.ctx, C.c_void_p(dev_ptr), n_ranges, payload_bytes, used_bytes, depth))

    def shard_bytes(self, width, height, shard):
        sh = VxShard(*shard)
        return lib().vx_shard_bytes(width, height, C.byref(sh))

    def pack_shard(self, shard, dev_ptr):
        sh = VxShard(*shard)
        self._check(lib().vx_pack_shard(self.ctx, C.byref(sh), C.c_void_p(dev_ptr)))

    def unpack_shard(self, shard, dev_ptr):
        sh = VxShard(*shard)
        self._check(lib().vx_unpack_shard(self.ctx, C.byref(sh), C.c_void_p(dev_ptr)))

    def set_streams(self, render=None, upload=None, picker=None):
        self._check(lib().vx_set_streams(self.ctx, C.c_void_p(render), C.c_void_p(upload), C.c_void_p(picker)))

    def stream(self, which=0):
        p = C.c_void_p()
        self._check(lib().vx_stream(self.ctx, which, C.byref(p)))
        return p.value

    def frame_ipc_handle(self):
        buf = (C.c_uint8 * 64)()
        self._check(lib().vx_frame_ipc_handle(self.ctx, buf))
        return bytes(buf)

    def serialize_chunks(self, blocks, lods=None, want_records=True, blocks_ptr=None, n_chunks=None):
        """vx_serialize_chunks_esvo. blocks: (n, 32768) uint32 host array, or pass blocks_ptr (device pointer) + n_chunks.
        Returns (infos [CHUNK_INFO_DTYPE], records uint8 array or None, kernel_ms)."""
        if blocks_ptr is None:
            blocks = np.ascontiguousarray(blocks, dtype=np.uint32).reshape(-1, 32768)
            n_chunks, blocks_ptr = len(blocks), blocks.ctypes.data
        infos = np.zeros(n_chunks, dtype=CHUNK_INFO_DTYPE)
        lods_a = np.ascontiguousarray(lods, dtype=np.uint8) if lods is not None else None
        total = C.c_uint64()
        cap = n_chunks * 4681 * 48 if want_records else 0
        rec = np.zeros(cap, dtype=np.uint8) if want_records else None
        self._check(lib().vx_serialize_chunks_esvo(self.ctx, C.c_void_p(blocks_ptr), n_chunks, _ptr(lods_a) if lods_a is not None else None,
                                                   _ptr(infos), _ptr(rec) if rec is not None else None, cap, C.byref(total)))
        ms = C.c_float()
        self._check(lib().vx_serialize_chunks_result(self.ctx, None, C.byref(ms)))
        return infos, (rec[:total.value] if rec is not None else None), ms.value

    def serialize_records_ptr(self):
        """Device pointer of the records of the last serialize_chunks call."""
        p = C.c_void_p()
        self._check(lib().vx_serialize_chunks_result(self.ctx, C.byref(p), None))
        return p.value

    def svo_write_device(self, range_offset, src_dev_ptr, length):
        self._check(lib().vx_svo_write_device(self.ctx, range_offset, C.c_void_p(src_dev_ptr), length))

    def frame8_ipc_handle(self):
        buf = (C.c_uint8 * 64)()
        self._check(lib().vx_frame8_ipc_handle(self.ctx, buf))
        return bytes(buf)

    def open_peer_frame8(self, handle):
        self._check(lib().vx_open_peer_frame8(self.ctx, (C.c_uint8 * 64)(*handle)))

    def open_peer_frame(self, handle):
        buf = (C.c_uint8 * 64)(*handle)
        self._check(lib().vx_open_peer_frame(self.ctx, buf))

    def close_peer_frame(self):
        self._check(lib().vx_close_peer_frame(self.ctx))

    def sync_ipc_handle(self):
        buf = (C.c_uint8 * 64)()
        self._check(lib().vx_sync_ipc_handle(self.ctx, buf))
        return bytes(buf)

    def open_peer_sync(self, handle):
        self._check(lib().vx_open_peer_sync(self.ctx, (C.c_uint8 * 64)(*handle)))

    def close_peer_sync(self):
        self._check(lib().vx_close_peer_sync(self.ctx))

    def frame_signal(self, slot, value):
        self._check(lib().vx_frame_signal(self.ctx, slot, value))

    def frame_wait(self, first_slot, n_slots, value):
        self._check(lib().vx_frame_wait(self.ctx, first_slot, n_slots, value))

    def frame_gate(self, slot, value):
        self._check(lib().vx_frame_gate(self.ctx, slot, value))

    def frame_sync_errors(self):
        n = C.c_uint32()
        self._check(lib().vx_frame_sync_errors(self.ctx, C.byref(n)))
        return n.value

    def probe_read_bandwidth(self, nbytes, passes):
        """GB/s of a streaming read of an nbytes scratch buffer (fits the L2: L2 bandwidth; several times the L2: HBM)."""
        g = C.c_float()
        self._check(lib().vx_probe_read_bandwidth(self.ctx, nbytes, passes, C.byref(g)))
        return g.value

    def scatter_errors(self):
        n = C.c_uint32()
        self._check(lib().vx_svo_scatter_errors(self.ctx, C.byref(n)))
        return n.value

    def frame_flags_reset(self):
        self._check(lib().vx_frame_flags_reset(self.ctx))

    def frame_device_ptr(self):
        p, w, h = C.c_void_p(), C.c_uint32(), C.c_uint32()
        self._check(lib().vx_frame_device_ptr(self.ctx, C.byref(p), C.byref(w), C.byref(h)))
        return p.value, w.value, h.value

    def raycast_device(self, tasks_ptr, n, results_ptr):
        self._check(lib().vx_raycast_device(self.ctx, C.c_void_p(tasks_ptr), n, C.c_void_p(results_ptr)))

    def raycast_wait(self):
        self._check(lib().vx_raycast_wait(self.ctx))

    def frame_stats(self, which=0):
        st = VxFrameStats()
        self._check(lib().vx_frame_stats(self.ctx, which, C.byref(st)))
        return st.as_dict()

    def launch_count(self):
        return lib().vx_launch_count(self.ctx)


from . import sharded  # noqa: E402,F401  (multi-GPU tile shards: voxelrs_b200.sharded.ShardedFrame)
