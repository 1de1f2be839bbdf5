// kernels.cuh — the __global__ entry points of libvoxelrt (sm_100a).
//
//   render_persistent_kernel  world.glsl main(): persistent warps, warp-level work fetch, per-lane ray refill
//   render_simple_kernel      same pixels, one thread per pixel start to finish (A/B baseline)
//   raycast_kernel            picker.glsl main(): batched rays, same refill scheme
//   debug_cast_kernel         svo.test.glsl main(): one ray, per-iteration frames
//   mip / rgba8 / shard copy / dirty-range scatter / texture-opacity utility kernels
#pragma once
#include "../../include/voxelrt.h"
#include "traverse.cuh"

namespace vx {

// Decodes a 2-D Morton code (x in even bits).
__device__ __forceinline__ uint32_t compact1by1(uint32_t v) {
    v &= 0x55555555u;
    v = (v ^ (v >> 1)) & 0x33333333u;
    v = (v ^ (v >> 2)) & 0x0f0f0f0fu;
    v = (v ^ (v >> 4)) & 0x00ff00ffu;
    v = (v ^ (v >> 8)) & 0x0000ffffu;
    return v;
}

struct RenderArgs {
    Scene scene;
    RenderUniforms u;
    float4* frame;                // RGBA32F, row 0 = bottom (world.glsl:140)
    Counters* counters;
    unsigned int* work_counter;   // persistent kernels: next unclaimed strip
    uint32_t tiles_x, tiles_y;    // frame size in 8x4-pixel warp tiles
    uint32_t macro_x, macro_y;    // frame size in 4x4-tile (32x16 pixel) macro blocks
    uint32_t shard_rank, shard_size;
    uint32_t refill_threshold;    // leave the traversal loop when fewer lanes than this are still walking
};

// Work units. The frame is cut into macro blocks of 32x16 pixels (row-major over the frame; a shard owns every
// shard_size-th block). A macro block is 4 strips of 32x4 pixels, a strip is 4 warp tiles of 8x4 pixels laid side by
// side, and pixel p of a strip is lane p%32 of tile p/32: consecutive work indices, consecutive pixels of a warp and
// consecutive tiles of a strip are all spatial neighbours (coherent rays, shared nodes in L1).
__device__ __forceinline__ bool strip_origin(const RenderArgs& a, uint32_t strip, uint32_t& x0, uint32_t& y0) {
    const uint32_t macro = strip >> 2;
    if (a.shard_size > 1 && (macro % a.shard_size) != a.shard_rank) return false;
    x0 = (macro % a.macro_x) * 32;
    y0 = (macro / a.macro_x) * 16 + (strip & 3u) * 4;
    return x0 < a.u.width && y0 < a.u.height;
}
__device__ __forceinline__ void strip_pixel(uint32_t x0, uint32_t y0, uint32_t p, uint32_t& gx, uint32_t& gy) {
    gx = x0 + (p >> 5) * 8 + (p & 7u);
    gy = y0 + ((p >> 3) & 3u);
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

// texels the shader's textureLod reads for this lod (for the counters when the fetch itself is skipped)
__device__ __forceinline__ uint32_t logical_texels(const TexInfo* ti, float lod) {
    if (!(lod > 0.0f)) return 1;
    const uint32_t levels = __ldg(&ti->levels);
    const float l = gl_min(lod, (float)(levels - 1));
    const float fl = floorf(l);
    const uint32_t d1 = (uint32_t)fl, d2 = (d1 + 1 < levels) ? d1 + 1 : levels - 1;
    return (d2 == d1 || l - fl == 0.0f) ? 4 : 8;
}

// Evaluates the leaf a render ray stopped at (svo.esvo.glsl:185-265 with cast_translucent = true).
// need_color = false (shadow rays): only "alpha > 0" matters, and for layers whose every texel is opaque that is
// known without touching the texture.
template <bool COUNT>
__device__ __forceinline__ bool render_leaf(Ray& r, const Scene& s, const float* unorm, float inv_scale, bool need_color, Leaf& g, float& tex_lod,
                                            float4& color, Counters& cnt) {
    leaf_geom<COUNT>(r, s, inv_scale, g, cnt);
    int tex_id;
    leaf_texture(s, g, tex_id, tex_lod);
    bool alpha_pos;
    if (need_color || !layer_is_opaque(s.tex, tex_id)) {
        uint32_t nf = 0;
        color = texture_lod(s.tex, unorm, g.u, g.v, tex_id, tex_lod, &nf);          // :237
        if (COUNT) cnt.tex_fetches += nf;
        alpha_pos = color.w > 0.0f;
    } else {
        if (COUNT) cnt.tex_fetches += logical_texels(s.tex, tex_lod);
        alpha_pos = true;
    }
    const bool first_of_kind = r.adjacent_leaf_count == 0 || g.value != r.last_leaf_value;   // :241
    if (alpha_pos && first_of_kind) return true;                                     // :242
    ++r.adjacent_leaf_count;                                                         // :264-265
    r.last_leaf_value = g.value;
    return false;
}

// ---- persistent render kernel ----------------------------------------------------------------------------------------
// grid = SMs x resident CTAs/SM, 4 warps per CTA. Every warp runs this loop until the frame is done:
//   refill   idle lanes take the next pixels of the warp's current strip (32x4 px); when the strip is used up lane 0
//            claims the next one with a single atomicAdd and broadcasts it by shuffle (warp-level work fetch);
//   walk     all lanes step their ray (primary or shadow — same code, no divergence between the two kinds) in lock-step;
//            a warp vote after every step leaves the loop once fewer than `refill_threshold` lanes are still walking;
//   events   lanes that stopped at a leaf candidate evaluate it together (value, face, uv, material, texture, alpha);
//            accepted primary hits are shaded in place and re-armed as shadow rays, finished pixels are written with a
//            streaming 16-byte store and the lane becomes idle again.
#define PH_IDLE 0
#define PH_PRIMARY 1
#define PH_SHADOW 2

// MINB = resident CTAs per SM the register allocator must allow (5 -> 96 regs, 6 -> 80, 8 -> 64): occupancy against
// spills is an empirical trade, so the variants are all built and chosen at run time (vx_set_option 4).
template <bool VEC, bool COUNT, int MINB>
__global__ void __launch_bounds__(128, MINB) render_persistent_kernel(RenderArgs a) {
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene, smem_raw);
    const uint32_t lane = threadIdx.x & 31, tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    const float octree_scale = __uint_as_float(__ldg(a.scene.desc - 1));
    const float inv_scale = 1.0f / octree_scale;
    const uint32_t n_strips = a.macro_x * a.macro_y * 4u;
    float* cold = sm.cold + tid;   // cold[k * nthr]: 0-3 colour, 4 lit, 5-7 primary direction
    Counters cnt = {0, 0, 0, 0, 0, 0};

    // warp-uniform work state
    uint32_t strip_x0 = 0, strip_y0 = 0, next_px = 128;
    bool more_work = true;
    // lane state
    int phase = PH_IDLE, ev = RAY_CONTINUE;
    bool after_leaf = false;
    uint32_t pix = 0;
    Ray r;

    for (;;) {
        // ---------------------------------------------------------------- refill
        unsigned want = __ballot_sync(0xffffffffu, phase == PH_IDLE);
        while (want && more_work) {
            if (next_px >= 128) {
                uint32_t strip = 0;
                if (lane == 0) strip = atomicAdd(a.work_counter, 1u);
                strip = __shfl_sync(0xffffffffu, strip, 0);
                if (strip >= n_strips) { more_work = false; break; }
                if (!strip_origin(a, strip, strip_x0, strip_y0)) continue;
                next_px = 0;
            }
            const uint32_t n_take = min((uint32_t)__popc(want), 128u - next_px);
            const uint32_t my_rank = __popc(want & lanemask_lt);
            if (((want >> lane) & 1u) && my_rank < n_take) {
                uint32_t gx, gy;
                strip_pixel(strip_x0, strip_y0, next_px + my_rank, gx, gy);
                if (gx < a.u.width && gy < a.u.height) {
                    float ox, oy, oz, dx, dy, dz;
                    primary_ray(a.u, gx, gy, ox, oy, oz, dx, dy, dz);
                    cold[5 * nthr] = dx; cold[6 * nthr] = dy; cold[7 * nthr] = dz;
                    ray_init(r, a.scene, octree_scale, ox, oy, oz, dx, dy, dz, -1.0f);
                    pix = gy * a.u.width + gx;
                    phase = PH_PRIMARY; ev = RAY_CONTINUE; after_leaf = false;
                    cnt.primary_rays++;
                }
            }
            next_px += n_take;
            want = __ballot_sync(0xffffffffu, phase == PH_IDLE);
        }
        const unsigned busy = __ballot_sync(0xffffffffu, phase != PH_IDLE);
        if (!busy) break;
        const int thresh = min((int)a.refill_threshold, __popc(busy));

        // ---------------------------------------------------------------- walk
        for (;;) {
            if (phase != PH_IDLE && ev == RAY_CONTINUE) {
                ev = ray_step<false, VEC, COUNT>(r, a.scene, sm.stack, cnt, after_leaf);
                after_leaf = false;
            }
            if (__popc(__ballot_sync(0xffffffffu, phase != PH_IDLE && ev == RAY_CONTINUE)) < thresh) break;
        }

        // ---------------------------------------------------------------- events
        if (ev == RAY_LEAF) {
            Leaf g; float tex_lod; float4 c;
            if (render_leaf<COUNT>(r, a.scene, sm.unorm, inv_scale, phase == PH_PRIMARY, g, tex_lod, c, cnt)) {
                if (phase == PH_PRIMARY) {
                    float px, py, pz;
                    leaf_pos(r, g, inv_scale, px, py, pz);
                    Shade sh;
                    sh.r = c.x; sh.g = c.y; sh.b = c.z; sh.a = c.w;
                    uint32_t nf = 0;
                    shade_hit(a.scene, sm.unorm, a.u, g, tex_lod, px, py, pz, sh, &nf);
                    if (COUNT) cnt.tex_fetches += nf;
                    if (sh.done) {
                        __stcs(a.frame + pix, make_float4(sh.r, sh.g, sh.b, sh.a));
                        phase = PH_IDLE;
                    } else if (sh.want_shadow) {
                        cold[0] = sh.r; cold[nthr] = sh.g; cold[2 * nthr] = sh.b; cold[3 * nthr] = sh.a; cold[4 * nthr] = sh.lit;
                        ray_init(r, a.scene, octree_scale, sh.sox, sh.soy, sh.soz, -a.u.lx, -a.u.ly, -a.u.lz, -1.0f);
                        phase = PH_SHADOW;
                        cnt.shadow_rays++;
                    } else {
                        __stcs(a.frame + pix, shade_finish(a.u, sh.r, sh.g, sh.b, sh.a, sh.lit, 1.0f));
                        phase = PH_IDLE;
                    }
                } else {   // the shadow ray is blocked
                    __stcs(a.frame + pix, shade_finish(a.u, cold[0], cold[nthr], cold[2 * nthr], cold[3 * nthr], cold[4 * nthr], 0.0f));
                    phase = PH_IDLE;
                }
            } else {
                after_leaf = true;   // translucent / repeated leaf: resume this iteration at ADVANCE
            }
            ev = RAY_CONTINUE;
        } else if (ev == RAY_MISS) {
            if (phase == PH_PRIMARY) __stcs(a.frame + pix, sky_color(cold[5 * nthr], cold[6 * nthr], cold[7 * nthr]));
            else __stcs(a.frame + pix, shade_finish(a.u, cold[0], cold[nthr], cold[2 * nthr], cold[3 * nthr], cold[4 * nthr], 1.0f));
            phase = PH_IDLE;
            ev = RAY_CONTINUE;
        }
    }
    flush_counters(a.counters, cnt);
}

// ---- simple render kernel: one thread per pixel, whole pipeline sequentially (A/B baseline) --------------------------
// Block = 128 threads = 4 warps = the 4 tiles of one 32x4 strip; blockIdx = strip index.
template <bool VEC, bool COUNT>
__device__ __forceinline__ bool trace_render_ray(Ray& r, const Scene& s, const Smem& sm, float inv_scale, bool need_color, Leaf& g, float& tex_lod,
                                                 float4& c, Counters& cnt) {
    bool after_leaf = false;
    for (;;) {
        const int ev = ray_step<false, VEC, COUNT>(r, s, sm.stack, cnt, after_leaf);
        after_leaf = false;
        if (ev == RAY_CONTINUE) continue;
        if (ev == RAY_MISS) return false;
        if (render_leaf<COUNT>(r, s, sm.unorm, inv_scale, need_color, g, tex_lod, c, cnt)) return true;
        after_leaf = true;
    }
}

template <bool VEC, bool COUNT>
__global__ void __launch_bounds__(128) render_simple_kernel(RenderArgs a) {
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene, smem_raw);
    uint32_t x0, y0, gx = 0, gy = 0;
    const bool have = strip_origin(a, blockIdx.x, x0, y0);
    if (have) strip_pixel(x0, y0, threadIdx.x, gx, gy);
    Counters cnt = {0, 0, 0, 0, 0, 0};
    if (have && gx < a.u.width && gy < a.u.height) {
        const float octree_scale = __uint_as_float(__ldg(a.scene.desc - 1));
        const float inv_scale = 1.0f / octree_scale;
        float ox, oy, oz, dx, dy, dz;
        primary_ray(a.u, gx, gy, ox, oy, oz, dx, dy, dz);
        Ray r; Leaf g; float tex_lod; float4 c;
        ray_init(r, a.scene, octree_scale, ox, oy, oz, dx, dy, dz, -1.0f);
        cnt.primary_rays = 1;
        float4 color;
        if (trace_render_ray<VEC, COUNT>(r, a.scene, sm, inv_scale, true, g, tex_lod, c, cnt)) {
            float px, py, pz;
            leaf_pos(r, g, inv_scale, px, py, pz);
            Shade sh;
            sh.r = c.x; sh.g = c.y; sh.b = c.z; sh.a = c.w;
            uint32_t nf = 0;
            shade_hit(a.scene, sm.unorm, a.u, g, tex_lod, px, py, pz, sh, &nf);
            if (COUNT) cnt.tex_fetches += nf;
            if (sh.done) {
                color = make_float4(sh.r, sh.g, sh.b, sh.a);
            } else {
                float shadow = 1.0f;
                if (sh.want_shadow) {
                    cnt.shadow_rays = 1;
                    ray_init(r, a.scene, octree_scale, sh.sox, sh.soy, sh.soz, -a.u.lx, -a.u.ly, -a.u.lz, -1.0f);
                    Leaf g2; float lod2; float4 c2;
                    shadow = trace_render_ray<VEC, COUNT>(r, a.scene, sm, inv_scale, false, g2, lod2, c2, cnt) ? 0.0f : 1.0f;
                }
                color = shade_finish(a.u, sh.r, sh.g, sh.b, sh.a, sh.lit, shadow);
            }
        } else {
            color = sky_color(dx, dy, dz);
        }
        a.frame[(size_t)gy * a.u.width + gx] = color;
    }
    flush_counters(a.counters, cnt);
}

// ---- picker kernel: picker.glsl main() -------------------------------------------------------------------------------
// Same scheme as the render kernel: warps claim runs of 128 consecutive tasks, lanes refill from the run as their ray
// ends, so a warp is not held hostage by its longest ray (16 M random rays differ in length by 100x).
struct RaycastArgs {
    Scene scene;
    const float4* tasks;      // VxPickerTask = 3 x float4
    float4* results;          // VxPickerResult = 3 x float4
    unsigned long long n;
    Counters* counters;
    unsigned long long* work_counter;
    uint32_t refill_threshold;
};

__device__ __forceinline__ void write_picker_result(float4* results, unsigned long long i, const Ray& r, const Leaf* g, float inv_scale) {
    float4 o0 = make_float4(-1.0f, 0.0f, 0.0f, 0.0f), o1 = make_float4(0, 0, 0, 0), o2 = make_float4(0, 0, 0, 0);
    if (g && g->dst > 0.0f) {                                          // picker.glsl:40 (res.t > 0)
        float px, py, pz;
        leaf_pos(r, *g, inv_scale, px, py, pz);
        o0.x = g->dst; o0.y = __uint_as_float(r.inside_voxel);
        o1 = make_float4(px, py, pz, 0.0f);
        const int axis = g->face_id >> 1;
        const float sgn = (g->face_id & 1) ? 1.0f : -1.0f;             // FACE_NORMALS, svo.glsl:2-9
        o2 = make_float4(axis == 0 ? sgn : 0.0f, axis == 1 ? sgn : 0.0f, axis == 2 ? sgn : 0.0f, 0.0f);
    }
    __stcs(results + 3 * i, o0); __stcs(results + 3 * i + 1, o1); __stcs(results + 3 * i + 2, o2);
}

template <bool VEC, bool COUNT>
__global__ void __launch_bounds__(128) raycast_kernel(RaycastArgs a) {
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene, smem_raw);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    const float octree_scale = __uint_as_float(__ldg(a.scene.desc - 1));
    const float inv_scale = 1.0f / octree_scale;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    unsigned long long run_base = 0;
    uint32_t next = 128, run_len = 128;
    bool more_work = true;
    bool active = false;
    int ev = RAY_CONTINUE;
    unsigned long long my_task = 0;
    Ray r;

    for (;;) {
        unsigned want = __ballot_sync(0xffffffffu, !active);
        while (want && more_work) {
            if (next >= run_len) {
                if (lane == 0) run_base = atomicAdd(a.work_counter, 128ull);
                run_base = __shfl_sync(0xffffffffu, run_base, 0);
                if (run_base >= a.n) { more_work = false; break; }
                run_len = (uint32_t)min(128ull, a.n - run_base);
                next = 0;
            }
            const uint32_t n_take = min((uint32_t)__popc(want), run_len - next);
            const uint32_t my_rank = __popc(want & lanemask_lt);
            if (((want >> lane) & 1u) && my_rank < n_take) {
                my_task = run_base + next + my_rank;
                const float4 t0 = __ldg(a.tasks + 3 * my_task), t1 = __ldg(a.tasks + 3 * my_task + 1), t2 = __ldg(a.tasks + 3 * my_task + 2);
                ray_init(r, a.scene, octree_scale, t1.x, t1.y, t1.z, t2.x, t2.y, t2.z, t0.x);
                active = true; ev = RAY_CONTINUE;
                cnt.primary_rays++;
            }
            next += n_take;
            want = __ballot_sync(0xffffffffu, !active);
        }
        const unsigned busy = __ballot_sync(0xffffffffu, active);
        if (!busy) break;
        const int thresh = min((int)a.refill_threshold, __popc(busy));
        for (;;) {
            if (active && ev == RAY_CONTINUE) ev = ray_step<true, VEC, COUNT>(r, a.scene, sm.stack, cnt);
            if (__popc(__ballot_sync(0xffffffffu, active && ev == RAY_CONTINUE)) < thresh) break;
        }
        if (ev == RAY_LEAF) {
            // cast_translucent = false: the first leaf is the hit whatever its texel is (svo.esvo.glsl:241-242); the picker
            // never reads the colour (picker.glsl:40-44), so the texture is not sampled at all.
            Leaf g;
            leaf_geom<COUNT>(r, a.scene, inv_scale, g, cnt);
            write_picker_result(a.results, my_task, r, &g, inv_scale);
            active = false; ev = RAY_CONTINUE;
        } else if (ev == RAY_MISS) {
            write_picker_result(a.results, my_task, r, nullptr, inv_scale);
            active = false; ev = RAY_CONTINUE;
        }
    }
    flush_counters(a.counters, cnt);
}

// ---- debug cast: svo.test.glsl main(), one thread, records every iteration --------------------------------------------
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

// The step machine does not carry the shader's (ptr, parent_octant_idx); the debug kernel shadows them (plus their
// stacks) next to it to emit reference-format frames.
__global__ void debug_cast_kernel(DebugArgs a) {
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene, smem_raw);
    const Scene& s = a.scene;
    const float octree_scale = __uint_as_float(__ldg(s.desc - 1));
    const float inv_scale = 1.0f / octree_scale;
    Ray r;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    ray_init(r, s, octree_scale, a.pos[0], a.pos[1], a.pos[2], a.dir[0], a.dir[1], a.dir[2], a.max_dst);
    uint32_t ptr = 0, pidx = 0;
    uint32_t ptr_stack[VX_MAX_SCALE + 1], pidx_stack[VX_MAX_SCALE + 1];
    for (int i = 0; i <= VX_MAX_SCALE; ++i) { ptr_stack[i] = 0; pidx_stack[i] = 0; }
    uint32_t n = 0;
    bool after_leaf = false, hit = false;
    Leaf g; float tex_lod = 0.0f; float4 color = make_float4(0, 0, 0, 0);
    for (;;) {
        const uint32_t oi = (uint32_t)((r.idx ^ (r.idx >> 4)) & 7);
        const int scale_before = r.scale;
        const uint32_t rec_before = r.rec;
        if (!after_leaf) {
            // the frame the shader emits at :175 for this iteration (if it gets past :152-156)
            const bool will_run = !(r.max_dst >= 0.0f && r.t_min > r.max_dst) && r.steps < VX_MAX_STEPS;
            if (will_run) {
                if (n < a.frames_cap) {
                    VxDebugFrame& f = a.frames[n];
                    f.t_min = r.t_min * inv_scale; f.ptr = ptr; f.idx = oi; f.parent_octant_idx = pidx; f.scale = r.scale;
                    f.is_child = (r.desc & ((1u << oi) << 8)) != 0; f.is_leaf = (r.desc & (1u << oi)) != 0;
                    f.crossed_boundary = 0; f.next_ptr = 0;
                }
                ++n;
            }
        }
        const float h_before = r.h;
        const float tcx = __fmaf_rn(r.px, r.tcx, -r.tbx), tcy = __fmaf_rn(r.py, r.tcy, -r.tby), tcz = __fmaf_rn(r.pz, r.tcz, -r.tbz);
        const float tc_max = tmin2(tmin2(tcx, tcy), tcz);
        const int ev = ray_step<true, false, false>(r, s, sm.stack, cnt, after_leaf);
        after_leaf = false;
        if (ev == RAY_MISS) break;
        if (ev == RAY_LEAF) {
            leaf_geom<false>(r, s, inv_scale, g, cnt);
            int tex_id;
            leaf_texture(s, g, tex_id, tex_lod);
            uint32_t nf = 0;
            color = texture_lod(s.tex, sm.unorm, g.u, g.v, tex_id, tex_lod, &nf);   // the shader samples in both modes (:237)
            const bool first_of_kind = r.adjacent_leaf_count == 0 || g.value != r.last_leaf_value;
            if ((color.w > 0.0f || !a.cast_translucent) && first_of_kind) { hit = true; break; }
            ++r.adjacent_leaf_count; r.last_leaf_value = g.value;
            after_leaf = true;
            continue;
        }
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
    if (hit) {
        o.t = g.dst; o.value = g.value; o.face_id = g.face_id;
        leaf_pos(r, g, inv_scale, o.pos[0], o.pos[1], o.pos[2]);
        o.uv[0] = g.u; o.uv[1] = g.v; o.lod = tex_lod;
        o.color[0] = color.x; o.color[1] = color.y; o.color[2] = color.z; o.color[3] = color.w;
    }
    *a.n_frames = n;
}

// ---- small utility kernels ---------------------------------------------------------------------------------------------

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

// Clears bit L of *opaque_layers when any texel (any level) of layer L has alpha == 0. One launch per level.
__global__ void opaque_kernel(const uint32_t* texels, uint32_t per_layer, uint32_t layers, unsigned long long* opaque_layers) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_layer * layers) return;
    const uint32_t layer = i / per_layer;
    if ((texels[i] >> 24) == 0 && layer < 64) atomicAnd(opaque_layers, ~(1ull << layer));
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

// Shard <-> contiguous buffer. One CTA of 128 threads moves one 32x16-pixel macro block (4 pixels per thread as
// 16-byte vectors, coalesced on both sides); block b of the grid handles the b-th macro block OWNED by the shard.
template <bool PACK>
__global__ void __launch_bounds__(128) shard_copy_kernel(float4* frame, float4* packed, uint32_t width, uint32_t height, uint32_t macro_x,
                                                         uint32_t n_macros, uint32_t rank, uint32_t size) {
    const uint32_t macro = blockIdx.x * size + rank;
    if (macro >= n_macros) return;
    const uint32_t x0 = (macro % macro_x) * 32, y0 = (macro / macro_x) * 16;
    float4* p = packed + (size_t)blockIdx.x * 512;
    for (uint32_t i = threadIdx.x; i < 512; i += 128) {
        const uint32_t x = x0 + (i & 31), y = y0 + (i >> 5);
        if (x < width && y < height) {
            if (PACK) p[i] = frame[(size_t)y * width + x];
            else frame[(size_t)y * width + x] = p[i];
        } else if (PACK) {
            p[i] = make_float4(0, 0, 0, 0);
        }
    }
}

// Applies a packed dirty set (n VxRange headers, then [24 head bytes][range 0 bytes][range 1 bytes]...) to the world
// buffer of a replica. Byte-granular because ranges are only 4-byte aligned relative to each other.
__global__ void scatter_ranges_kernel(uint8_t* world, const uint8_t* packed, uint32_t n_ranges, unsigned long long payload_bytes) {
    const VxRange* hdr = reinterpret_cast<const VxRange*>(packed);
    const uint32_t* payload = reinterpret_cast<const uint32_t*>(packed + (size_t)n_ranges * sizeof(VxRange));
    uint32_t* w32 = reinterpret_cast<uint32_t*>(world);   // world + 0 is 8-byte aligned; every range offset/length is a multiple of 4
    const unsigned long long words = payload_bytes / 4;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) {
        if (i < 6) { w32[i] = payload[i]; continue; }
        unsigned long long off = 6;
        for (uint32_t k = 0; k < n_ranges; ++k) {
            const unsigned long long len = hdr[k].length / 4;
            if (i < off + len) { w32[6 + hdr[k].offset / 4 + (i - off)] = payload[i]; break; }
            off += len;
        }
    }
}

}  // namespace vx
