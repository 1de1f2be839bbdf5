"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (voxel-rs_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_lib = None


class OctreeResult(C.Structure):
    _fields_ = [("t", C.c_float), ("value", C.c_uint32), ("face_id", C.c_int32), ("pos", C.c_float * 3), ("uv", C.c_float * 2),
                ("color", C.c_float * 4), ("lod", C.c_float), ("inside_voxel", C.c_uint32)]

    def as_dict(self):
        return {"t": self.t, "value": self.value, "face_id": self.face_id, "pos": tuple(self.pos), "uv": tuple(self.uv),
                "color": tuple(self.color), "lod": self.lod, "inside_voxel": bool(self.inside_voxel)}


class DebugFrame(C.Structure):
    _fields_ = [("t_min", C.c_float), ("ptr", C.c_uint32), ("idx", C.c_uint32), ("parent_octant_idx", C.c_uint32),
                ("scale", C.c_int32), ("is_child", C.c_int32), ("is_leaf", C.c_int32), ("crossed_boundary", C.c_int32),
                ("next_ptr", C.c_uint32)]

    def as_tuple(self):
        return (self.t_min, self.ptr, self.idx, self.parent_octant_idx, self.scale, self.is_child, self.is_leaf)


class Counters(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("steps", C.c_uint64), ("pushes", C.c_uint64),
                ("leaf_tests", C.c_uint64), ("tex_fetches", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class RenderParams(C.Structure):   # same layout as VxRenderParams
    _fields_ = [("view", C.c_float * 16), ("fov_y_rad", C.c_float), ("aspect_ratio", C.c_float),
                ("ambient_intensity", C.c_float), ("light_dir", C.c_float * 3), ("cam_pos", C.c_float * 3),
                ("highlight_pos", C.c_float * 3), ("render_shadows", C.c_uint32), ("shadow_distance", C.c_float)]


HIT_DTYPE = np.dtype([("t", "<f4"), ("value", "<u4"), ("face_id", "<i4"), ("pos", "<f4", 3), ("near_boundary", "<u4")])


def build(force=False):
    so = os.path.join(_DIR, "liboracle.so")
    src = os.path.join(_DIR, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.run(["make", "-C", _DIR, "-B", "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    P = C.c_void_p
    L.vxo_texture_create.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]; L.vxo_texture_create.restype = P
    L.vxo_texture_destroy.argtypes = [P]; L.vxo_texture_destroy.restype = None
    L.vxo_texture_levels.argtypes = [P]; L.vxo_texture_levels.restype = C.c_uint32
    L.vxo_texture_level.argtypes = [P, C.c_uint32, P]; L.vxo_texture_level.restype = C.c_uint64
    scene = [P, C.c_uint64, P, C.c_uint32, P, C.c_int]
    L.vxo_debug_cast.argtypes = scene + [C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_float, C.c_uint32,
                                         C.POINTER(OctreeResult), C.POINTER(DebugFrame), C.c_uint32, C.POINTER(C.c_uint32)]
    L.vxo_debug_cast.restype = None
    L.vxo_raycast.argtypes = scene + [P, C.c_uint64, P, C.POINTER(Counters), C.c_int]; L.vxo_raycast.restype = None
    L.vxo_render.argtypes = scene + [C.POINTER(RenderParams), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, P, C.POINTER(Counters), C.c_int]
    L.vxo_render.restype = None
    L.vxo_primary_hits.argtypes = scene + [C.POINTER(RenderParams), C.c_uint32, C.c_uint32, P, C.c_int]; L.vxo_primary_hits.restype = None
    L.vxo_render_steps.argtypes = scene + [C.POINTER(RenderParams), C.c_uint32, C.c_uint32, P, P, C.c_int]; L.vxo_render_steps.restype = None
    L.vxo_to_rgba8.argtypes = [P, C.c_uint64, P]; L.vxo_to_rgba8.restype = None
    L.vxo_max_threads.argtypes = []; L.vxo_max_threads.restype = C.c_int
    L.vxo_set_clip.argtypes = [C.c_int]; L.vxo_set_clip.restype = None
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """World bytes (GPU-buffer layout) + material table + texture array, held for oracle calls."""

    def __init__(self, world_bytes, materials, textures_level0, mip_levels, fmt=0):
        """materials: bytes-like of n*32 (VxMaterial records); textures_level0: [layers, h, w, 4] uint8, v-flipped as uploaded."""
        self.fmt = int(fmt)   # 0 = ESVO, 1 = CSVO (the shader's SVO_TYPE)
        self.world = np.ascontiguousarray(world_bytes, dtype=np.uint8)
        self.materials = np.ascontiguousarray(np.frombuffer(bytes(materials), dtype=np.uint8))
        self.n_materials = len(self.materials) // 32
        tex = np.ascontiguousarray(textures_level0, dtype=np.uint8)
        layers, h, w, _ = tex.shape
        self.tex = lib().vxo_texture_create(_ptr(tex), w, h, layers, mip_levels)
        self.tex_shape = (layers, h, w)

    def __del__(self):
        if getattr(self, "tex", None):
            lib().vxo_texture_destroy(self.tex)
            self.tex = None

    def _args(self):
        return (_ptr(self.world), len(self.world), _ptr(self.materials), self.n_materials, self.tex, self.fmt)

    def mip_level(self, level):
        n = lib().vxo_texture_level(self.tex, level, None)
        out = np.zeros(n, np.uint8)
        lib().vxo_texture_level(self.tex, level, _ptr(out))
        layers, h, w = self.tex_shape
        return out.reshape(layers, max(h >> level, 1), max(w >> level, 1), 4)

    def debug_cast(self, pos, direction, max_dst, cast_translucent, frames_cap=100):
        d = np.array(direction, dtype=np.float32)
        d = d / np.sqrt(np.float32((d * d).sum(dtype=np.float32)))   # svo_shader_tests.rs:246
        res = OctreeResult()
        frames = (DebugFrame * frames_cap)()
        n = C.c_uint32()
        lib().vxo_debug_cast(*self._args(), C.byref((C.c_float * 3)(*[float(x) for x in pos])), C.byref((C.c_float * 3)(*[float(x) for x in d])),
                             max_dst, int(cast_translucent), C.byref(res), frames, frames_cap, C.byref(n))
        return res, [frames[i] for i in range(min(n.value, frames_cap))], n.value

    def raycast(self, tasks, threads=0):
        """tasks: structured array of 48-byte picker tasks; returns (results same-shaped 48-byte records, counters)."""
        tasks = np.ascontiguousarray(tasks)
        assert tasks.dtype.itemsize == 48
        res = np.zeros(len(tasks), dtype=np.dtype({"names": ["dst", "inside_voxel", "pos", "normal"],
                                                   "formats": ["<f4", "<u4", ("<f4", 3), ("<f4", 3)], "offsets": [0, 4, 16, 32], "itemsize": 48}))
        cnt = Counters()
        lib().vxo_raycast(*self._args(), _ptr(tasks), len(tasks), _ptr(res), C.byref(cnt), threads)
        return res, cnt.as_dict()

    def render(self, params, width, height, y0=0, y1=None, threads=0, out=None):
        """params: any struct with VxRenderParams layout. Returns (RGBA32F [h,w,4] with rows outside [y0,y1) untouched, counters)."""
        y1 = height if y1 is None else y1
        if out is None:
            out = np.zeros((height, width, 4), dtype=np.float32)
        cnt = Counters()
        p = RenderParams.from_buffer_copy(bytes(params))
        lib().vxo_render(*self._args(), C.byref(p), width, height, y0, y1, _ptr(out), C.byref(cnt), threads)
        return out, cnt.as_dict()

    def render_steps(self, params, width, height, threads=0):
        """Per pixel: loop iterations of the primary ray and of the shadow ray (0 = none) — for offline scheduling analysis."""
        a, b = np.zeros((height, width), dtype=np.uint32), np.zeros((height, width), dtype=np.uint32)
        p = RenderParams.from_buffer_copy(bytes(params))
        lib().vxo_render_steps(*self._args(), C.byref(p), width, height, _ptr(a), _ptr(b), threads)
        return a, b

    def primary_hits(self, params, width, height, threads=0):
        out = np.zeros((height, width), dtype=HIT_DTYPE)
        p = RenderParams.from_buffer_copy(bytes(params))
        lib().vxo_primary_hits(*self._args(), C.byref(p), width, height, _ptr(out), threads)
        return out


def to_rgba8(rgba32f):
    a = np.ascontiguousarray(rgba32f, dtype=np.float32)
    out = np.zeros(a.shape, dtype=np.uint8)
    lib().vxo_to_rgba8(_ptr(a), a.size // 4, _ptr(out))
    return out


def set_clip(on):
    """Product extension (off by default = the shader as written): stop a ray once it has left the occupied box of the world, like
    the CUDA kernels do (vx_set_option 12). Results are unchanged by construction; only the iteration counters follow it."""
    lib().vxo_set_clip(1 if on else 0)


def max_threads():
    return lib().vxo_max_threads()
