// voxelrt.cu — kernels and C ABI of libvoxelrt (see include/voxelrt.h for the reference seams).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -shared -Xcompiler -fPIC
// (--fmad=false is part of the numeric contract, see traverse.cuh).
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/voxelrt.h"
#include "traverse.cuh"

namespace vx {

// ================================================================ kernels ==

// Decodes a 2-D Morton code (x in even bits).
__device__ __forceinline__ uint32_t compact1by1(uint32_t v) {
    v &= 0x55555555u;
    v = (v ^ (v >> 1)) & 0x33333333u;
    v = (v ^ (v >> 2)) & 0x0f0f0f0fu;
    v = (v ^ (v >> 4)) & 0x00ff00ffu;
    v = (v ^ (v >> 8)) & 0x0000ffffu;
    return v;
}

#define VX_TILES_PER_FETCH 4u

struct RenderArgs {
    Scene scene;
    RenderUniforms u;
    float4* frame;            // RGBA32F, row 0 = bottom (world.glsl:140)
    Counters* counters;
    unsigned int* work_counter;   // persistent kernels: next unclaimed work item
    uint32_t tiles_x, tiles_y;    // frame size in 8x4-pixel warp tiles
    uint32_t macro_x, macro_y;    // frame size in 4x4-tile (32x16 pixel) macro blocks
    uint32_t shard_rank, shard_size;
};

// Tile order: macro blocks of 4x4 warp tiles row-major over the frame, Morton order inside a block,
// so consecutive tile indices are spatial neighbours (coherent rays for neighbouring warps) and only
// edge blocks contain tiles outside the frame. Shards own whole macro blocks, interleaved.
__device__ __forceinline__ bool tile_coords(const RenderArgs& a, uint32_t t, uint32_t& tx, uint32_t& ty) {
    const uint32_t macro = t >> 4, within = t & 15u;
    const uint32_t mx = macro % a.macro_x, my = macro / a.macro_x;
    tx = mx * 4 + compact1by1(within); ty = my * 4 + compact1by1(within >> 1);
    if (a.shard_size > 1 && (macro % a.shard_size) != a.shard_rank) return false;
    return tx < a.tiles_x && ty < a.tiles_y;
}

__device__ __forceinline__ Stack make_stack(const Scene& s, uint32_t* smem) {
    Stack st;
    st.stride = blockDim.x; st.levels = s.stack_levels;
    st.rec = smem;
    st.desc = smem + (size_t)st.levels * st.stride;
    st.t_max = reinterpret_cast<float*>(smem + 2 * (size_t)st.levels * st.stride);
    return st;
}

__device__ __forceinline__ void flush_counters(Counters* g, const Counters& c) {
    // warp-aggregate then one atomic per warp and counter
    unsigned long long v[6] = {c.primary_rays, c.shadow_rays, c.steps, c.pushes, c.leaf_tests, c.tex_fetches};
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(g);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(dst + k, x);
    }
}

// ---- simple render kernel: one thread per pixel, whole pipeline inline (A/B baseline) -------------
// Block = 128 threads = 4 warps, each warp an 8x4 pixel tile; blockIdx enumerates groups of 4 tiles in
// tile order (tile_coords).
template <bool VEC, bool COUNT>
__global__ void __launch_bounds__(128) render_simple_kernel(RenderArgs a) {
    extern __shared__ uint32_t smem[];
    const Stack st = make_stack(a.scene, smem);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t tx, ty;
    const bool have_tile = tile_coords(a, blockIdx.x * 4 + warp, tx, ty);
    const uint32_t gx = tx * 8 + (lane & 7), gy = ty * 4 + (lane >> 3);
    Counters cnt = {0, 0, 0, 0, 0, 0};
    if (have_tile && gx < a.u.width && gy < a.u.height) {
        const float octree_scale = __uint_as_float(__ldg(a.scene.desc - 1));
        const float inv_scale = 1.0f / octree_scale;
        float ox, oy, oz, dx, dy, dz;
        primary_ray(a.u, gx, gy, ox, oy, oz, dx, dy, dz);
        Ray r; Hit h;
        ray_init(r, a.scene, octree_scale, ox, oy, oz, dx, dy, dz, -1.0f);
        cnt.primary_rays = 1;
        int status;
        while ((status = ray_step<true, VEC, COUNT>(r, a.scene, st, inv_scale, h, cnt)) == RAY_CONTINUE) {}
        float4 color;
        if (status == RAY_HIT) {
            Shade sh;
            shade_hit(a.scene, a.u, h, sh, cnt.tex_fetches);
            float shadow = 1.0f;
            if (!sh.done && sh.want_shadow) {
                cnt.shadow_rays = 1;
                Hit h2;
                ray_init(r, a.scene, octree_scale, sh.sox, sh.soy, sh.soz, -a.u.lx, -a.u.ly, -a.u.lz, -1.0f);
                while ((status = ray_step<true, VEC, COUNT>(r, a.scene, st, inv_scale, h2, cnt)) == RAY_CONTINUE) {}
                shadow = (status == RAY_HIT) ? 0.0f : 1.0f;
            }
            color = shade_finish(a.u, sh, shadow);
        } else {
            color = sky_color(dx, dy, dz);
        }
        a.frame[(size_t)gy * a.u.width + gx] = color;
    }
    flush_counters(a.counters, cnt);
}

// ---- persistent render kernel -----------------------------------------------------------------------
// One resident CTA set (grid = SMs x CTAs/SM), every warp loops: lane 0 claims the next Morton-ordered
// 8x4 pixel tile with one atomicAdd (warp-level work fetch), all 32 lanes trace their primary rays in
// lock-step through the shared step machine; a lane whose primary ray hits is shaded in place and
// re-armed with its SHADOW ray, so primary and shadow rays of one tile share the same traversal loop
// (no second pass, no per-pixel state in memory). When fewer than REFILL lanes are still busy and the
// warp holds a whole tile of finished lanes... the tile is written and the next one fetched.
template <bool VEC, bool COUNT>
__global__ void __launch_bounds__(128) render_persistent_kernel(RenderArgs a) {
    extern __shared__ uint32_t smem[];
    const Stack st = make_stack(a.scene, smem);
    const uint32_t lane = threadIdx.x & 31;
    const float octree_scale = __uint_as_float(__ldg(a.scene.desc - 1));
    const float inv_scale = 1.0f / octree_scale;
    const uint32_t n_tiles = a.macro_x * a.macro_y * 16u;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    uint32_t batch = 0, batch_left = 0;   // tiles are claimed VX_TILES_PER_FETCH at a time (one atomic per 4 tiles)

    for (;;) {
        // ---- warp-level work fetch: lane 0 claims the next run of tiles, the warp shares it by shuffle
        if (batch_left == 0) {
            if (lane == 0) batch = atomicAdd(a.work_counter, (unsigned)VX_TILES_PER_FETCH);
            batch = __shfl_sync(0xffffffffu, batch, 0);
            batch_left = VX_TILES_PER_FETCH;
        }
        const uint32_t t = batch + (VX_TILES_PER_FETCH - batch_left);
        --batch_left;
        if (t >= n_tiles) break;
        uint32_t tx, ty;
        if (!tile_coords(a, t, tx, ty)) continue;
        const uint32_t gx = tx * 8 + (lane & 7), gy = ty * 4 + (lane >> 3);
        const bool live = gx < a.u.width && gy < a.u.height;

        Ray r; Hit h; Shade sh;
        float ox, oy, oz, dx = 0, dy = 0, dz = 0;
        // phase: 0 = primary in flight, 1 = shadow in flight, 2 = finished
        int phase = 2;
        float4 color = make_float4(0, 0, 0, 0);
        if (live) {
            primary_ray(a.u, gx, gy, ox, oy, oz, dx, dy, dz);
            ray_init(r, a.scene, octree_scale, ox, oy, oz, dx, dy, dz, -1.0f);
            phase = 0;
            cnt.primary_rays++;
        }
        // ---- lock-step traversal; warp vote ends the loop when every lane is finished
        while (__any_sync(0xffffffffu, phase != 2)) {
            if (phase != 2) {
                const int status = ray_step<true, VEC, COUNT>(r, a.scene, st, inv_scale, h, cnt);
                if (status != RAY_CONTINUE) {
                    if (phase == 0) {
                        if (status == RAY_HIT) {
                            shade_hit(a.scene, a.u, h, sh, cnt.tex_fetches);
                            if (!sh.done && sh.want_shadow) {
                                ray_init(r, a.scene, octree_scale, sh.sox, sh.soy, sh.soz, -a.u.lx, -a.u.ly, -a.u.lz, -1.0f);
                                phase = 1;
                                cnt.shadow_rays++;
                            } else {
                                color = shade_finish(a.u, sh, 1.0f);
                                phase = 2;
                            }
                        } else {
                            color = sky_color(dx, dy, dz);
                            phase = 2;
                        }
                    } else {
                        color = shade_finish(a.u, sh, status == RAY_HIT ? 0.0f : 1.0f);
                        phase = 2;
                    }
                }
            }
        }
        if (live) __stcs(a.frame + (size_t)gy * a.u.width + gx, color);   // streaming store: the frame is write-once
    }
    flush_counters(a.counters, cnt);
}

// ---- picker kernel: picker.glsl main(), one thread per task ------------------------------------------
struct RaycastArgs {
    Scene scene;
    const float4* tasks;      // VxPickerTask = 3 x float4
    float4* results;          // VxPickerResult = 3 x float4
    unsigned long long n;
    Counters* counters;
};

template <bool VEC, bool COUNT>
__global__ void __launch_bounds__(128) raycast_kernel(RaycastArgs a) {
    extern __shared__ uint32_t smem[];
    const Stack st = make_stack(a.scene, smem);
    const float octree_scale = __uint_as_float(__ldg(a.scene.desc - 1));
    const float inv_scale = 1.0f / octree_scale;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const float4 t0 = __ldg(a.tasks + 3 * i), t1 = __ldg(a.tasks + 3 * i + 1), t2 = __ldg(a.tasks + 3 * i + 2);
        Ray r; Hit h;
        ray_init(r, a.scene, octree_scale, t1.x, t1.y, t1.z, t2.x, t2.y, t2.z, t0.x);
        cnt.primary_rays++;
        int status;
        while ((status = ray_step<false, VEC, COUNT>(r, a.scene, st, inv_scale, h, cnt)) == RAY_CONTINUE) {}
        float4 o0 = make_float4(-1.0f, 0.0f, 0.0f, 0.0f), o1 = make_float4(0, 0, 0, 0), o2 = make_float4(0, 0, 0, 0);
        if (status == RAY_HIT && h.t > 0.0f) {                         // picker.glsl:40
            o0.x = h.t; o0.y = __uint_as_float(r.inside_voxel);
            o1 = make_float4(h.posx, h.posy, h.posz, 0.0f);
            const int axis = h.face_id >> 1;
            const float sgn = (h.face_id & 1) ? 1.0f : -1.0f;          // FACE_NORMALS, svo.glsl:2-9
            o2 = make_float4(axis == 0 ? sgn : 0.0f, axis == 1 ? sgn : 0.0f, axis == 2 ? sgn : 0.0f, 0.0f);
        }
        __stcs(a.results + 3 * i, o0); __stcs(a.results + 3 * i + 1, o1); __stcs(a.results + 3 * i + 2, o2);
    }
    flush_counters(a.counters, cnt);
}

// ---- debug cast: svo.test.glsl main(), one thread, records every iteration ----------------------------
struct DebugArgs {
    Scene scene;
    float pos[3], dir[3];
    float max_dst;
    uint32_t cast_translucent;
    VxOctreeResult* result;
    VxDebugFrame* frames;
    uint32_t frames_cap;
    uint32_t* n_frames;
};

// The step machine does not carry the shader's (ptr, parent_octant_idx); the debug kernel shadows
// them (plus their stacks) next to it to emit reference-format frames.
template <bool TRANSLUCENT>
__device__ void debug_cast_impl(const DebugArgs& a, const Stack& st) {
    const Scene& s = a.scene;
    const float octree_scale = __uint_as_float(__ldg(s.desc - 1));
    const float inv_scale = 1.0f / octree_scale;
    Ray r; Hit h;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    ray_init(r, s, octree_scale, a.pos[0], a.pos[1], a.pos[2], a.dir[0], a.dir[1], a.dir[2], a.max_dst);
    uint32_t ptr = 0, pidx = 0;
    uint32_t ptr_stack[VX_MAX_SCALE + 1], pidx_stack[VX_MAX_SCALE + 1];
    for (int i = 0; i <= VX_MAX_SCALE; ++i) { ptr_stack[i] = 0; pidx_stack[i] = 0; }
    uint32_t n = 0;
    int status;
    for (;;) {
        // the frame the shader would emit at :175 for this iteration (if it gets that far)
        const bool will_run = !(r.max_dst >= 0.0f && r.t_min > r.max_dst) && r.steps < VX_MAX_STEPS;
        const uint32_t oi = (uint32_t)(r.idx ^ r.octant_mask);
        const int scale_before = r.scale;
        const uint32_t rec_before = r.rec;
        if (will_run) {
            if (n < a.frames_cap) {
                VxDebugFrame& f = a.frames[n];
                f.t_min = r.t_min * inv_scale; f.ptr = ptr; f.idx = oi; f.parent_octant_idx = pidx; f.scale = r.scale;
                f.is_child = (r.desc & ((1u << oi) << 8)) != 0; f.is_leaf = (r.desc & (1u << oi)) != 0;
                f.crossed_boundary = 0; f.next_ptr = 0;
            }
            ++n;
        }
        const float h_before = r.h;
        const float tcx = r.px * r.tcx - r.tbx, tcy = r.py * r.tcy - r.tby, tcz = r.pz * r.tcz - r.tbz;
        const float tc_max = gl_min(gl_min(tcx, tcy), tcz);
        status = ray_step<TRANSLUCENT, false, false>(r, s, st, inv_scale, h, cnt);
        if (status != RAY_CONTINUE) break;
        if (r.scale == scale_before - 1) {            // PUSH happened
            if (tc_max < h_before) { ptr_stack[scale_before] = ptr; pidx_stack[scale_before] = pidx; }
            ptr = rec_before; pidx = oi;
        } else if (r.scale > scale_before) {          // POP happened
            ptr = ptr_stack[r.scale]; pidx = pidx_stack[r.scale];
        }
    }
    VxOctreeResult& o = *a.result;
    o.t = -1.0f; o.value = 0; o.face_id = 0; o.pos[0] = o.pos[1] = o.pos[2] = 0; o.uv[0] = o.uv[1] = 0;
    o.color[0] = o.color[1] = o.color[2] = o.color[3] = 0; o.lod = 0; o.inside_voxel = r.inside_voxel;
    if (status == RAY_HIT) {
        o.t = h.t; o.value = h.value; o.face_id = h.face_id; o.pos[0] = h.posx; o.pos[1] = h.posy; o.pos[2] = h.posz;
        o.uv[0] = h.u; o.uv[1] = h.v; o.lod = h.lod;
        if (TRANSLUCENT) { o.color[0] = h.r; o.color[1] = h.g; o.color[2] = h.b; o.color[3] = h.a; }
    }
    *a.n_frames = n;
}

__global__ void debug_cast_kernel(DebugArgs a) {
    extern __shared__ uint32_t smem[];
    const Stack st = make_stack(a.scene, smem);
    if (a.cast_translucent) {
        debug_cast_impl<true>(a, st);
    } else {
        debug_cast_impl<false>(a, st);
        // cast_translucent=false still samples the texture in the shader (svo.esvo.glsl:237) and reports
        // its colour; reproduce that for the debug record only.
        VxOctreeResult& o = *a.result;
        if (o.t >= 0.0f) {
            const Scene& s = a.scene;
            const Material* m = s.materials + (o.value < s.n_materials ? o.value : s.n_materials - 1);
            int tex_id = m->tex_side;
            if (o.face_id == 3) tex_id = m->tex_top; else if (o.face_id == 2) tex_id = m->tex_bottom;
            float sm = gl_clamp((o.t - 15.0f) / (25.0f - 15.0f), 0.0f, 1.0f);
            sm = (sm * sm) * (3.0f - 2.0f * sm);
            const float tex_lod = (sm * (o.t - 15.0f)) * 0.05f;
            unsigned long long nf = 0;
            const float4 c = texture_lod(s, o.uv[0], o.uv[1], tex_id, tex_lod, nf);
            o.color[0] = c.x; o.color[1] = c.y; o.color[2] = c.z; o.color[3] = c.w; o.lod = tex_lod;
        }
    }
}

// ---- small utility kernels -----------------------------------------------------------------------------

// glGenerateMipmap stand-in: level l+1 texel = rounded mean of the 2x2 block below (texture_array.rs:258-260)
__global__ void mip_kernel(const uint32_t* src, uint32_t* dst, uint32_t pw, uint32_t ph, uint32_t cw, uint32_t ch, uint32_t layers) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cw * ch * layers) return;
    const uint32_t x = i % cw, y = (i / cw) % ch, layer = i / (cw * ch);
    const uint32_t x0 = 2 * x, x1 = (2 * x + 1 < pw) ? 2 * x + 1 : pw - 1, y0 = 2 * y, y1 = (2 * y + 1 < ph) ? 2 * y + 1 : ph - 1;
    const uint32_t* b = src + (size_t)layer * pw * ph;
    const uint32_t t00 = b[y0 * pw + x0], t10 = b[y0 * pw + x1], t01 = b[y1 * pw + x0], t11 = b[y1 * pw + x1];
    uint32_t out = 0;
    for (int c = 0; c < 4; ++c) {
        const uint32_t sum = ((t00 >> (8 * c)) & 0xff) + ((t10 >> (8 * c)) & 0xff) + ((t01 >> (8 * c)) & 0xff) + ((t11 >> (8 * c)) & 0xff);
        out |= ((sum + 2) >> 2) << (8 * c);
    }
    dst[i] = out;
}

// glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE) of the RGBA32F attachment (framebuffer.rs:97-105)
__global__ void rgba8_kernel(const float4* frame, uint32_t* out, unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = frame[i];
    const float v[4] = {c.x, c.y, c.z, c.w};
    uint32_t p = 0;
    for (int k = 0; k < 4; ++k) {
        float f = v[k];
        if (!(f == f)) f = 0.0f;
        f = gl_clamp(f, 0.0f, 1.0f);
        p |= (uint32_t)(int)(f * 255.0f + 0.5f) << (8 * k);
    }
    out[i] = p;
}

// Applies a packed dirty set (n VxRange headers, then payload) to the world buffer (replica update).
__global__ void scatter_ranges_kernel(uint8_t* world, const uint8_t* packed, uint32_t n_ranges, unsigned long long payload_bytes) {
    const VxRange* hdr = reinterpret_cast<const VxRange*>(packed);
    const uint8_t* payload = packed + (size_t)n_ranges * sizeof(VxRange);
    // payload = [24 head bytes][range 0 bytes][range 1 bytes]...; one block sweeps the whole payload
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < payload_bytes; i += stride) {
        if (i < 24) { world[i] = payload[i]; continue; }
        // find the range containing payload byte i (ranges are few; linear scan)
        unsigned long long off = 24;
        for (uint32_t k = 0; k < n_ranges; ++k) {
            const unsigned long long len = hdr[k].length;
            if (i < off + len) { world[24 + hdr[k].offset + (i - off)] = payload[i]; break; }
            off += len;
        }
    }
}

}  // namespace vx

// ================================================================== host ==

using namespace vx;

static thread_local std::string g_create_error;

struct VxCtx {
    VxConfig cfg{};
    int sm_count = 0;
    std::string err;

    cudaStream_t s_render = nullptr, s_upload = nullptr, s_picker = nullptr;
    cudaEvent_t e_upload = nullptr, e_render = nullptr, e_picker = nullptr;
    cudaEvent_t t0_render = nullptr, t1_render = nullptr, t0_picker = nullptr, t1_picker = nullptr;

    uint8_t* d_world_raw = nullptr;   // allocation; GL byte 0 lives at d_world_raw + 8 so that records are 16-B aligned
    uint8_t* d_world = nullptr;
    uint8_t* h_mirror = nullptr;      // pinned, capacity bytes
    uint8_t* h_stage = nullptr;       // pinned staging ring for async dirty uploads
    size_t stage_cap = 0;
    uint64_t hot_off = 0, hot_len = 0;
    bool have_svo = false;

    Material* d_materials = nullptr;
    uint32_t n_materials = 0;
    uint32_t* d_texels = nullptr;
    uint32_t tex_w = 0, tex_h = 0, tex_layers = 0, tex_levels = 0;
    uint32_t tex_off[16] = {};

    float4* d_frame = nullptr;
    uint32_t* d_frame8 = nullptr;
    uint32_t frame_w = 0, frame_h = 0;

    float4* d_tasks = nullptr;
    float4* d_results = nullptr;

    Counters* d_counters = nullptr;   // [0] render, [1] raycast
    unsigned int* d_work = nullptr;
    VxFrameStats last_render{}, last_raycast{};
    bool render_timed = false, raycast_timed = false;

    VxStats stats{};
    uint64_t launches = 0;

    // options (vx_set_option)
    uint64_t opt_simple = 0, opt_vec = 1, opt_count = 0, opt_ctas_per_sm = 0, opt_l2_window = 1;
};

static int fail(VxCtx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}
#define CU(ctx, call)                                                                                                   \
    do {                                                                                                                \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess) return fail(ctx, VX_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static Scene make_scene(const VxCtx* c) {
    Scene s{};
    s.desc = reinterpret_cast<const uint32_t*>(c->d_world + 4);
    s.desc_words = (uint32_t)((c->cfg.svo_capacity_bytes - 4) / 4);
    s.materials = c->d_materials; s.n_materials = c->n_materials;
    s.texels = c->d_texels; s.tex_w = c->tex_w; s.tex_h = c->tex_h; s.tex_layers = c->tex_layers; s.tex_levels = c->tex_levels;
    for (int i = 0; i < 16; ++i) s.tex_off[i] = c->tex_off[i];
    uint32_t levels = c->stats.depth + 1;
    s.stack_levels = levels < 2 ? 2 : (levels > VX_MAX_SCALE ? VX_MAX_SCALE : levels);
    return s;
}
static size_t stack_smem_bytes(const Scene& s, uint32_t threads) { return (size_t)3 * s.stack_levels * threads * 4; }

extern "C" {

const char* vx_build_info(void) { return "libvoxelrt sm_100a --fmad=false " __DATE__ " " __TIME__; }

const char* vx_last_error(const VxCtx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

uint64_t vx_launch_count(const VxCtx* ctx) { return ctx ? ctx->launches : 0; }

int vx_set_option(VxCtx* ctx, uint32_t option, uint64_t value);

int vx_create(const VxConfig* cfg, VxCtx** out) {
    if (!cfg || !out) return fail(nullptr, VX_E_ARG, "vx_create: null argument");
    if (cfg->svo_capacity_bytes < 64) return fail(nullptr, VX_E_ARG, "vx_create: svo_capacity_bytes too small");
    if (cfg->svo_capacity_bytes > (1ull << 34)) return fail(nullptr, VX_E_ARG, "vx_create: svo_capacity_bytes > 16 GiB (u32 word pointers)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VX_E_CUDA, "vx_create: no CUDA device (%s) — libvoxelrt has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, VX_E_ARG, "vx_create: device %d out of range (%d devices)", cfg->device, ndev);
    VxCtx* c = new VxCtx();
    c->cfg = *cfg;
    *out = nullptr;
#define CUC(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            int rc_ = fail(nullptr, VX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));               \
            vx_destroy(c);                                                                             \
            return rc_;                                                                                \
        }                                                                                              \
    } while (0)
    CUC(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, cfg->device));
    c->sm_count = prop.multiProcessorCount;
    CUC(cudaStreamCreateWithFlags(&c->s_render, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->s_upload, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->s_picker, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&c->e_upload, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->e_render, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->e_picker, cudaEventDisableTiming));
    CUC(cudaEventCreate(&c->t0_render)); CUC(cudaEventCreate(&c->t1_render));
    CUC(cudaEventCreate(&c->t0_picker)); CUC(cudaEventCreate(&c->t1_picker));
    const size_t cap = (size_t)cfg->svo_capacity_bytes;
    CUC(cudaMalloc(&c->d_world_raw, cap + 64));
    c->d_world = c->d_world_raw + 8;
    CUC(cudaMemsetAsync(c->d_world_raw, 0, cap + 64, c->s_upload));
    CUC(cudaHostAlloc(&c->h_mirror, cap, cudaHostAllocDefault));
    std::memset(c->h_mirror, 0, cap < (1u << 20) ? cap : (1u << 20));
    c->stage_cap = cap < (64u << 20) ? cap : (64u << 20);
    CUC(cudaHostAlloc(&c->h_stage, c->stage_cap, cudaHostAllocDefault));
    if (cfg->max_width && cfg->max_height) {
        const size_t px = (size_t)cfg->max_width * cfg->max_height;
        CUC(cudaMalloc(&c->d_frame, px * sizeof(float4)));
        CUC(cudaMalloc(&c->d_frame8, px * 4));
    }
    if (cfg->max_rays) {
        CUC(cudaMalloc(&c->d_tasks, (size_t)cfg->max_rays * 48));
        CUC(cudaMalloc(&c->d_results, (size_t)cfg->max_rays * 48));
    }
    CUC(cudaMalloc(&c->d_counters, 2 * sizeof(Counters)));
    CUC(cudaMemsetAsync(c->d_counters, 0, 2 * sizeof(Counters), c->s_upload));
    CUC(cudaMalloc(&c->d_work, 64));
    CUC(cudaEventRecord(c->e_upload, c->s_upload));
    CUC(cudaStreamSynchronize(c->s_upload));
    c->stats.capacity_bytes = cap;
    c->opt_simple = (cfg->flags & VX_FLAG_KERNEL_SIMPLE) ? 1 : 0;
    c->opt_l2_window = (cfg->flags & VX_FLAG_NO_L2_WINDOW) ? 0 : 1;
#undef CUC
    *out = c;
    return VX_OK;
}

void vx_destroy(VxCtx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    if (c->d_world_raw) cudaFree(c->d_world_raw);
    if (c->h_mirror) cudaFreeHost(c->h_mirror);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_materials) cudaFree(c->d_materials);
    if (c->d_texels) cudaFree(c->d_texels);
    if (c->d_frame) cudaFree(c->d_frame);
    if (c->d_frame8) cudaFree(c->d_frame8);
    if (c->d_tasks) cudaFree(c->d_tasks);
    if (c->d_results) cudaFree(c->d_results);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->d_work) cudaFree(c->d_work);
    cudaEvent_t evs[] = {c->e_upload, c->e_render, c->e_picker, c->t0_render, c->t1_render, c->t0_picker, c->t1_picker};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    cudaStream_t ss[] = {c->s_render, c->s_upload, c->s_picker};
    for (cudaStream_t s : ss) if (s) cudaStreamDestroy(s);
    delete c;
}

// Runtime knobs for A/B measurements (not part of the reference surface).
//   1 = simple kernels (0/1)   2 = 128-bit node fetches (0/1)   3 = count steps/pushes/leaf tests (0/1)
//   4 = CTAs per SM for persistent kernels (0 = occupancy query)   5 = L2 access-policy window (0/1)
int vx_set_option(VxCtx* ctx, uint32_t option, uint64_t value) {
    if (!ctx) return VX_E_ARG;
    switch (option) {
        case 1: ctx->opt_simple = value; break;
        case 2: ctx->opt_vec = value; break;
        case 3: ctx->opt_count = value; break;
        case 4: ctx->opt_ctas_per_sm = value; break;
        case 5: ctx->opt_l2_window = value; break;
        default: return fail(ctx, VX_E_ARG, "vx_set_option: unknown option %u", option);
    }
    return VX_OK;
}

int vx_set_materials(VxCtx* c, const VxMaterial* materials, uint32_t count) {
    if (!c || !materials || count == 0) return fail(c, VX_E_ARG, "vx_set_materials: null/empty");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaDeviceSynchronize());
    if (c->d_materials) cudaFree(c->d_materials);
    c->d_materials = nullptr;
    CU(c, cudaMalloc(&c->d_materials, (size_t)count * sizeof(Material)));
    CU(c, cudaMemcpy(c->d_materials, materials, (size_t)count * sizeof(Material), cudaMemcpyHostToDevice));
    c->n_materials = count;
    return VX_OK;
}

int vx_set_textures(VxCtx* c, const uint8_t* rgba8, uint32_t width, uint32_t height, uint32_t layers, uint32_t mip_levels) {
    if (!c || !rgba8 || !width || !height || !layers) return fail(c, VX_E_ARG, "vx_set_textures: null/empty");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaDeviceSynchronize());
    uint32_t m = width < height ? width : height, il = 0;
    while ((m >> (il + 1)) != 0) ++il;                       // ilog2(min(w,h)), texture_array.rs:105
    uint32_t levels = mip_levels < il ? mip_levels : il;
    if (levels < 1) levels = 1;
    if (levels > 16) levels = 16;
    size_t total = 0;
    uint32_t off[16] = {};
    for (uint32_t l = 0; l < levels; ++l) {
        uint32_t wl = width >> l, hl = height >> l;
        wl = wl ? wl : 1; hl = hl ? hl : 1;
        off[l] = (uint32_t)total;
        total += (size_t)wl * hl * layers;
    }
    if (c->d_texels) cudaFree(c->d_texels);
    c->d_texels = nullptr;
    CU(c, cudaMalloc(&c->d_texels, total * 4));
    CU(c, cudaMemcpy(c->d_texels, rgba8, (size_t)width * height * layers * 4, cudaMemcpyHostToDevice));
    for (uint32_t l = 1; l < levels; ++l) {
        uint32_t pw = width >> (l - 1), ph = height >> (l - 1), cw = width >> l, ch = height >> l;
        pw = pw ? pw : 1; ph = ph ? ph : 1; cw = cw ? cw : 1; ch = ch ? ch : 1;
        const uint32_t n = cw * ch * layers;
        mip_kernel<<<(n + 255) / 256, 256, 0, c->s_upload>>>(c->d_texels + off[l - 1], c->d_texels + off[l], pw, ph, cw, ch, layers);
        c->launches++;
    }
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->s_upload));
    c->tex_w = width; c->tex_h = height; c->tex_layers = layers; c->tex_levels = levels;
    for (int i = 0; i < 16; ++i) c->tex_off[i] = off[i];
    return VX_OK;
}

uint8_t* vx_svo_host_mirror(VxCtx* c) { return c ? c->h_mirror : nullptr; }

int vx_svo_set_hot_range(VxCtx* c, uint64_t offset, uint64_t length) {
    if (!c) return VX_E_ARG;
    c->hot_off = offset; c->hot_len = length;
    return VX_OK;
}

// Persisting-L2 access-policy window over [preamble .. world-root octree] on the render/picker streams.
static void install_l2_window(VxCtx* c) {
    cudaStreamAttrValue attr{};
    if (c->opt_l2_window && c->hot_len) {
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, c->cfg.device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->cfg.device);
        size_t bytes = (size_t)c->hot_len;
        if (max_win > 0 && bytes > (size_t)max_win) bytes = (size_t)max_win;
        if (max_persist > 0) {
            size_t carve = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
        }
        attr.accessPolicyWindow.base_ptr = c->d_world + 24 + c->hot_off;
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
    }
    cudaStreamSetAttribute(c->s_render, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaStreamSetAttribute(c->s_picker, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();   // a refused window is a lost optimisation, not an error
}

int vx_svo_commit(VxCtx* c, float octree_scale, const VxRange* dirty, uint32_t n_dirty, uint64_t used_bytes, uint32_t depth) {
    if (!c || (n_dirty && !dirty)) return fail(c, VX_E_ARG, "vx_svo_commit: null argument");
    const uint64_t cap = c->cfg.svo_capacity_bytes;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_dirty; ++i) {
        // reference: assert!(start + length < dst_len) with dst_len = capacity - 1 relative to byte 4 (esvo.rs:328-331, svo.rs:180-181)
        if (dirty[i].offset + dirty[i].length + 24 > cap)
            return fail(c, VX_E_CAPACITY, "dst is not large enough: len=%llu range_start=%llu range_length=%llu", (unsigned long long)cap,
                        (unsigned long long)dirty[i].offset, (unsigned long long)dirty[i].length);
        total += dirty[i].length;
    }
    CU(c, cudaSetDevice(c->cfg.device));
    std::memcpy(c->h_mirror, &octree_scale, 4);                                  // svo.rs:173-175
    const bool staged = total + 24 <= c->stage_cap;
    // the staging block is reused: the previous upload must have drained it (normally long done)
    if (n_dirty && staged) CU(c, cudaStreamSynchronize(c->s_upload));
    // do not tear a frame / ray batch in flight (render_fence.wait(), svo.rs:178) — on the GPU timeline, not the CPU's
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_render, 0));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_picker, 0));
    if (n_dirty) {
        if (staged) {
            // the caller may rewrite the mirror as soon as we return: snapshot the dirty bytes into the pinned
            // staging block and copy from there asynchronously
            size_t off = 0;
            std::memcpy(c->h_stage, c->h_mirror, 24);
            CU(c, cudaMemcpyAsync(c->d_world, c->h_stage, 24, cudaMemcpyHostToDevice, c->s_upload));
            off = 24;
            for (uint32_t i = 0; i < n_dirty; ++i) {
                std::memcpy(c->h_stage + off, c->h_mirror + 24 + dirty[i].offset, dirty[i].length);
                CU(c, cudaMemcpyAsync(c->d_world + 24 + dirty[i].offset, c->h_stage + off, dirty[i].length, cudaMemcpyHostToDevice, c->s_upload));
                off += dirty[i].length;
            }
        } else {
            CU(c, cudaMemcpyAsync(c->d_world, c->h_mirror, 24, cudaMemcpyHostToDevice, c->s_upload));
            for (uint32_t i = 0; i < n_dirty; ++i)
                CU(c, cudaMemcpyAsync(c->d_world + 24 + dirty[i].offset, c->h_mirror + 24 + dirty[i].offset, dirty[i].length,
                                      cudaMemcpyHostToDevice, c->s_upload));
            CU(c, cudaStreamSynchronize(c->s_upload));   // bulk (re)load straight from the mirror: must finish before the caller reuses it
        }
        c->have_svo = true;
    }
    CU(c, cudaEventRecord(c->e_upload, c->s_upload));
    c->stats.used_bytes = used_bytes; c->stats.depth = depth;
    install_l2_window(c);
    return VX_OK;
}

int64_t vx_svo_pack_dirty(VxCtx* c, const VxRange* dirty, uint32_t n_dirty, void* out, uint64_t out_cap) {
    if (!c || (n_dirty && !dirty)) return VX_E_ARG;
    uint64_t need = (uint64_t)n_dirty * sizeof(VxRange) + 24;
    for (uint32_t i = 0; i < n_dirty; ++i) {
        if (dirty[i].offset + dirty[i].length + 24 > c->cfg.svo_capacity_bytes) return fail(c, VX_E_CAPACITY, "vx_svo_pack_dirty: range outside buffer");
        need += dirty[i].length;
    }
    if (!out) return (int64_t)need;
    if (need > out_cap) return fail(c, VX_E_CAPACITY, "vx_svo_pack_dirty: need %llu bytes, have %llu", (unsigned long long)need, (unsigned long long)out_cap);
    uint8_t* p = (uint8_t*)out;
    std::memcpy(p, dirty, (size_t)n_dirty * sizeof(VxRange));
    p += (size_t)n_dirty * sizeof(VxRange);
    std::memcpy(p, c->h_mirror, 24);
    p += 24;
    for (uint32_t i = 0; i < n_dirty; ++i) { std::memcpy(p, c->h_mirror + 24 + dirty[i].offset, dirty[i].length); p += dirty[i].length; }
    return (int64_t)need;
}

int vx_svo_commit_packed_device(VxCtx* c, const void* packed_dev, uint32_t n_dirty, uint64_t payload_bytes, uint64_t used_bytes, uint32_t depth) {
    if (!c || !packed_dev) return fail(c, VX_E_ARG, "vx_svo_commit_packed_device: null argument");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_render, 0));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_picker, 0));
    const unsigned long long pb = payload_bytes;   // 24 head bytes + range bytes
    const int blocks = (int)((pb + 255) / 256 < 4096 ? (pb + 255) / 256 : 4096);
    scatter_ranges_kernel<<<blocks > 0 ? blocks : 1, 256, 0, c->s_upload>>>(c->d_world, (const uint8_t*)packed_dev, n_dirty, pb);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->e_upload, c->s_upload));
    c->stats.used_bytes = used_bytes; c->stats.depth = depth;
    c->have_svo = true;
    install_l2_window(c);
    return VX_OK;
}

int vx_stats(const VxCtx* c, VxStats* out) {
    if (!c || !out) return VX_E_ARG;
    *out = c->stats;
    return VX_OK;
}

static int check_scene(VxCtx* c, const char* who) {
    if (!c->have_svo) return fail(c, VX_E_STATE, "%s: no SVO committed (vx_svo_commit)", who);
    if (!c->d_materials) return fail(c, VX_E_STATE, "%s: no materials (vx_set_materials)", who);
    if (!c->d_texels) return fail(c, VX_E_STATE, "%s: no textures (vx_set_textures)", who);
    return VX_OK;
}

static int persistent_grid(VxCtx* c, const void* kernel, int threads, size_t smem, int* grid) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return fail(c, VX_E_CUDA, "occupancy query: %s", cudaGetErrorString(e));
    if (per_sm < 1) per_sm = 1;
    if (c->opt_ctas_per_sm && (int)c->opt_ctas_per_sm < per_sm) per_sm = (int)c->opt_ctas_per_sm;
    *grid = per_sm * c->sm_count;
    return VX_OK;
}

int vx_render(VxCtx* c, const VxRenderParams* p, uint32_t width, uint32_t height, const VxShard* shard, float* rgba32f_out) {
    if (!c || !p || !width || !height) return fail(c, VX_E_ARG, "vx_render: null/empty argument");
    if (!c->d_frame || (uint64_t)width * height > (uint64_t)c->cfg.max_width * c->cfg.max_height)
        return fail(c, VX_E_CAPACITY, "vx_render: %ux%u exceeds the %ux%u framebuffer reserved at vx_create", width, height, c->cfg.max_width,
                    c->cfg.max_height);
    if (shard && (shard->world_size == 0 || shard->rank >= shard->world_size)) return fail(c, VX_E_ARG, "vx_render: bad shard");
    int rc = check_scene(c, "vx_render");
    if (rc) return rc;
    CU(c, cudaSetDevice(c->cfg.device));

    RenderArgs a{};
    a.scene = make_scene(c);
    std::memcpy(a.u.view, p->view, sizeof(a.u.view));
    a.u.tan_half_fov = tanf(p->fov_y_rad * 0.5f);                         // world.glsl:115, hoisted
    a.u.aspect = p->aspect_ratio; a.u.ambient = p->ambient_intensity;
    a.u.lx = p->light_dir[0]; a.u.ly = p->light_dir[1]; a.u.lz = p->light_dir[2];
    a.u.cx = p->cam_pos[0]; a.u.cy = p->cam_pos[1]; a.u.cz = p->cam_pos[2];
    a.u.hx = p->highlight_pos[0]; a.u.hy = p->highlight_pos[1]; a.u.hz = p->highlight_pos[2];
    a.u.render_shadows = p->render_shadows; a.u.shadow_distance = p->shadow_distance;
    a.u.width = width; a.u.height = height;
    a.frame = c->d_frame;
    a.counters = c->d_counters;
    a.work_counter = c->d_work;
    a.tiles_x = (width + 7) / 8; a.tiles_y = (height + 3) / 4;
    a.macro_x = (a.tiles_x + 3) / 4; a.macro_y = (a.tiles_y + 3) / 4;
    a.shard_rank = shard ? shard->rank : 0; a.shard_size = shard ? shard->world_size : 1;

    const int threads = 128;
    const size_t smem = stack_smem_bytes(a.scene, threads);
    CU(c, cudaStreamWaitEvent(c->s_render, c->e_upload, 0));
    CU(c, cudaMemsetAsync(c->d_counters, 0, sizeof(Counters), c->s_render));
    CU(c, cudaMemsetAsync(c->d_work, 0, sizeof(unsigned int), c->s_render));
    const bool vec = c->opt_vec != 0, count = c->opt_count != 0;
    CU(c, cudaEventRecord(c->t0_render, c->s_render));
    if (c->opt_simple) {
        const uint32_t blocks = a.macro_x * a.macro_y * 4;
        auto k = vec ? (count ? render_simple_kernel<true, true> : render_simple_kernel<true, false>)
                     : (count ? render_simple_kernel<false, true> : render_simple_kernel<false, false>);
        CU(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<blocks, threads, smem, c->s_render>>>(a);
    } else {
        auto k = vec ? (count ? render_persistent_kernel<true, true> : render_persistent_kernel<true, false>)
                     : (count ? render_persistent_kernel<false, true> : render_persistent_kernel<false, false>);
        CU(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = 0;
        rc = persistent_grid(c, (const void*)k, threads, smem, &grid);
        if (rc) return rc;
        k<<<grid, threads, smem, c->s_render>>>(a);
    }
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->t1_render, c->s_render));
    CU(c, cudaEventRecord(c->e_render, c->s_render));
    c->frame_w = width; c->frame_h = height;
    c->render_timed = true;
    if (rgba32f_out) {
        CU(c, cudaMemcpyAsync(rgba32f_out, c->d_frame, (size_t)width * height * sizeof(float4), cudaMemcpyDeviceToHost, c->s_render));
        CU(c, cudaStreamSynchronize(c->s_render));
    }
    return VX_OK;
}

int vx_render_wait(VxCtx* c) {
    if (!c) return VX_E_ARG;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->s_render));
    return VX_OK;
}

int vx_read_frame_rgba32f(VxCtx* c, float* out) {
    if (!c || !out || !c->frame_w) return fail(c, VX_E_ARG, "vx_read_frame_rgba32f: nothing rendered / null");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaMemcpyAsync(out, c->d_frame, (size_t)c->frame_w * c->frame_h * sizeof(float4), cudaMemcpyDeviceToHost, c->s_render));
    CU(c, cudaStreamSynchronize(c->s_render));
    return VX_OK;
}

int vx_read_frame_rgba8(VxCtx* c, uint8_t* out) {
    if (!c || !out || !c->frame_w) return fail(c, VX_E_ARG, "vx_read_frame_rgba8: nothing rendered / null");
    CU(c, cudaSetDevice(c->cfg.device));
    const unsigned long long n = (unsigned long long)c->frame_w * c->frame_h;
    rgba8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->s_render>>>(c->d_frame, c->d_frame8, n);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(out, c->d_frame8, n * 4, cudaMemcpyDeviceToHost, c->s_render));
    CU(c, cudaStreamSynchronize(c->s_render));
    return VX_OK;
}

int vx_frame_device_ptr(VxCtx* c, void** out_ptr, uint32_t* width, uint32_t* height) {
    if (!c || !out_ptr) return VX_E_ARG;
    *out_ptr = c->d_frame;
    if (width) *width = c->frame_w;
    if (height) *height = c->frame_h;
    return VX_OK;
}

static int launch_raycast(VxCtx* c, const float4* tasks_dev, uint64_t n, float4* results_dev) {
    RaycastArgs a{};
    a.scene = make_scene(c);
    a.tasks = tasks_dev; a.results = results_dev; a.n = n;
    a.counters = c->d_counters + 1;
    const int threads = 128;
    const size_t smem = stack_smem_bytes(a.scene, threads);
    const bool vec = c->opt_vec != 0, count = c->opt_count != 0;
    auto k = vec ? (count ? raycast_kernel<true, true> : raycast_kernel<true, false>) : (count ? raycast_kernel<false, true> : raycast_kernel<false, false>);
    CU(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = 0;
    int rc = persistent_grid(c, (const void*)k, threads, smem, &grid);
    if (rc) return rc;
    const uint64_t need = (n + threads - 1) / threads;
    if ((uint64_t)grid > need) grid = (int)need;
    CU(c, cudaMemsetAsync(c->d_counters + 1, 0, sizeof(Counters), c->s_picker));
    CU(c, cudaEventRecord(c->t0_picker, c->s_picker));
    k<<<grid, threads, smem, c->s_picker>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->t1_picker, c->s_picker));
    CU(c, cudaEventRecord(c->e_picker, c->s_picker));
    c->raycast_timed = true;
    return VX_OK;
}

int vx_raycast(VxCtx* c, const VxPickerTask* tasks, uint64_t n, VxPickerResult* results) {
    if (!c || (n && (!tasks || !results))) return fail(c, VX_E_ARG, "vx_raycast: null argument");
    if (n == 0) return VX_OK;
    if (n > c->cfg.max_rays) return fail(c, VX_E_CAPACITY, "vx_raycast: %llu rays exceed max_rays=%llu", (unsigned long long)n, (unsigned long long)c->cfg.max_rays);
    int rc = check_scene(c, "vx_raycast");
    if (rc) return rc;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamWaitEvent(c->s_picker, c->e_upload, 0));
    CU(c, cudaMemcpyAsync(c->d_tasks, tasks, (size_t)n * 48, cudaMemcpyHostToDevice, c->s_picker));
    rc = launch_raycast(c, c->d_tasks, n, c->d_results);
    if (rc) return rc;
    CU(c, cudaMemcpyAsync(results, c->d_results, (size_t)n * 48, cudaMemcpyDeviceToHost, c->s_picker));
    CU(c, cudaStreamSynchronize(c->s_picker));   // picker_fence.place(); .wait()  (svo.rs:248-249)
    return VX_OK;
}

int vx_raycast_device(VxCtx* c, const VxPickerTask* tasks_dev, uint64_t n, VxPickerResult* results_dev) {
    if (!c || !tasks_dev || !results_dev || !n) return fail(c, VX_E_ARG, "vx_raycast_device: null/empty argument");
    int rc = check_scene(c, "vx_raycast_device");
    if (rc) return rc;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamWaitEvent(c->s_picker, c->e_upload, 0));
    return launch_raycast(c, (const float4*)tasks_dev, n, (float4*)results_dev);
}

int vx_raycast_wait(VxCtx* c) {
    if (!c) return VX_E_ARG;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->s_picker));
    return VX_OK;
}

int vx_debug_cast(VxCtx* c, const float pos[3], const float dir[3], float max_dst, uint32_t cast_translucent, VxOctreeResult* result,
                  VxDebugFrame* frames, uint32_t frames_cap, uint32_t* n_frames) {
    if (!c || !pos || !dir || !result) return fail(c, VX_E_ARG, "vx_debug_cast: null argument");
    int rc = check_scene(c, "vx_debug_cast");
    if (rc) return rc;
    CU(c, cudaSetDevice(c->cfg.device));
    VxOctreeResult* d_res = nullptr; VxDebugFrame* d_frames = nullptr; uint32_t* d_n = nullptr;
    const uint32_t cap = frames ? frames_cap : 0;
    CU(c, cudaMalloc(&d_res, sizeof(VxOctreeResult)));
    CU(c, cudaMalloc(&d_frames, sizeof(VxDebugFrame) * (cap ? cap : 1)));
    CU(c, cudaMalloc(&d_n, 4));
    DebugArgs a{};
    a.scene = make_scene(c);
    for (int k = 0; k < 3; ++k) { a.pos[k] = pos[k]; a.dir[k] = dir[k]; }
    a.max_dst = max_dst; a.cast_translucent = cast_translucent;
    a.result = d_res; a.frames = d_frames; a.frames_cap = cap; a.n_frames = d_n;
    CU(c, cudaStreamWaitEvent(c->s_picker, c->e_upload, 0));
    debug_cast_kernel<<<1, 1, stack_smem_bytes(a.scene, 1), c->s_picker>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    uint32_t n = 0;
    CU(c, cudaMemcpyAsync(result, d_res, sizeof(VxOctreeResult), cudaMemcpyDeviceToHost, c->s_picker));
    CU(c, cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, c->s_picker));
    if (cap) CU(c, cudaMemcpyAsync(frames, d_frames, sizeof(VxDebugFrame) * cap, cudaMemcpyDeviceToHost, c->s_picker));
    CU(c, cudaStreamSynchronize(c->s_picker));
    if (n_frames) *n_frames = n;
    cudaFree(d_res); cudaFree(d_frames); cudaFree(d_n);
    return VX_OK;
}

int vx_frame_stats(VxCtx* c, int which, VxFrameStats* out) {
    if (!c || !out || which < 0 || which > 1) return fail(c, VX_E_ARG, "vx_frame_stats: bad argument");
    CU(c, cudaSetDevice(c->cfg.device));
    cudaStream_t s = which == 0 ? c->s_render : c->s_picker;
    CU(c, cudaStreamSynchronize(s));
    Counters h{};
    CU(c, cudaMemcpy(&h, c->d_counters + which, sizeof(Counters), cudaMemcpyDeviceToHost));
    VxFrameStats st{};
    st.primary_rays = h.primary_rays; st.shadow_rays = h.shadow_rays; st.steps = h.steps; st.pushes = h.pushes;
    st.leaf_tests = h.leaf_tests; st.tex_fetches = h.tex_fetches;
    const bool timed = which == 0 ? c->render_timed : c->raycast_timed;
    if (timed) {
        float ms = 0;
        CU(c, cudaEventElapsedTime(&ms, which == 0 ? c->t0_render : c->t0_picker, which == 0 ? c->t1_render : c->t1_picker));
        st.kernel_ms = ms;
    }
    *out = st;
    return VX_OK;
}

}  // extern "C"
