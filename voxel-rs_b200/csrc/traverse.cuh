// traverse.cuh — device-side ESVO ray traversal + shading for sm_100a.
//
// Behavioural contract = voxel-rs shaders (paths relative to tim-oster/voxel-rs):
//   intersect_octree  assets/shaders/svo.esvo.glsl:50-393
//   trace_ray         assets/shaders/world.glsl:27-90
//   get_sky_color     assets/shaders/world.glsl:92-108
//   textureLod state  src/graphics/texture_array.rs:200-203
// Not a transliteration:
//  * the traversal is a per-lane step machine (ray_init / ray_step) that persistent warps drive in
//    lock-step, swapping finished rays for new ones between steps;
//  * ray_step only walks the tree (PUSH / ADVANCE / POP). When it reaches a leaf candidate it returns
//    RAY_LEAF and the caller evaluates the leaf (value, face, uv, texture, translucency) OUTSIDE the
//    hot loop, where the lanes of a warp that sit on a leaf do it together; a rejected (translucent)
//    leaf re-enters the same iteration at its ADVANCE phase (`after_leaf`);
//  * node state is (rec, desc) = (record of the CURRENT octant, 16-bit child/leaf masks of its
//    children) instead of the shader's (ptr, parent_octant_idx). It is a pure function of
//    (ptr, parent_octant_idx) over an immutable buffer, so PUSH/POP/HIT visit exactly the same nodes
//    in exactly the same iterations, but an iteration that is not a PUSH or a leaf test touches no
//    memory (the shader re-reads its descriptor word every iteration, svo.esvo.glsl:168), and a
//    PUSH issues its two loads (child masks + child pointer) independently instead of back to back.
//
// Numerics (DESIGN.md "Numerics"): this TU is compiled with --fmad=false; float ops are IEEE single,
// round-to-nearest, in the shader's operation order, with an explicit fused multiply-add in exactly the
// three places the shader's comment asks for one (svo.esvo.glsl:97-99): t_corner (:159), the leaf
// entry corner (:197) and t_center (:275). Geometry results are bit-exact against oracle/oracle.cpp.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vx {

#define VX_MAX_STEPS 1000
#define VX_MAX_SCALE 23
#define VX_EPSILON 0.00000011920929f

// GLSL min/max as specified (min(x,y) = y<x ? y : x); used where a NaN / signed-zero choice could reach an output.
__device__ __forceinline__ float gl_min(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float gl_max(float x, float y) { return (x < y) ? y : x; }
__device__ __forceinline__ float gl_clamp(float x, float lo, float hi) { return gl_min(gl_max(x, lo), hi); }
__device__ __forceinline__ float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
// Inside the traversal the t values are finite and only ever compared, so the single-instruction FMNMX gives the
// same decisions as the GLSL form (they differ only in which zero / NaN is returned).
__device__ __forceinline__ float tmin2(float x, float y) { return fminf(x, y); }
__device__ __forceinline__ float tmax2(float x, float y) { return fmaxf(x, y); }

struct Material {   // = VxMaterial / svo.glsl:48-59
    float specular_pow, specular_strength;
    int tex_top, tex_side, tex_bottom, tex_top_normal, tex_side_normal, tex_bottom_normal;
};

// Texture array description, kept in device memory so that the (single, out-of-line) sampler can take a pointer.
struct TexInfo {
    const uint32_t* texels;  // RGBA8 texels, all mip levels of all layers; level l starts at off[l]
    uint32_t w, h, layers, levels;
    uint32_t off[16];
    unsigned long long opaque_layers;   // bit L set: every texel of every level of layer L has alpha > 0 (L < 64)
};

// Everything a kernel needs to read the scene. Passed by value as a kernel parameter.
struct Scene {
    const uint32_t* __restrict__ desc;   // descriptors[] = world buffer + 4 bytes; desc[ptr] 16-B aligned when ptr % 4 == 1
    uint32_t desc_words;                 // capacity in words (loads are clamped to it)
    const Material* __restrict__ materials;
    uint32_t n_materials;
    const TexInfo* __restrict__ tex;
    uint32_t stack_levels;               // entries of the per-thread traversal stack (= SVO depth + 1, <= 23)
};

struct Counters {   // = VxFrameStats counters
    unsigned long long primary_rays, shadow_rays, steps, pushes, leaf_tests, tex_fetches;
};

// Per-thread traversal stack in SHARED memory, one column per thread: slot s of thread t lives at
// base[s * stride + t]. Consecutive lanes hit consecutive banks whatever their (divergent) slot is.
struct Stack {
    uint32_t* rec;     // record of the octant at that level
    uint32_t* desc;    // its child/leaf masks
    float* t_max;
    uint32_t stride;   // = blockDim.x
    uint32_t levels;
    __device__ __forceinline__ uint32_t slot(int scale) const {
        uint32_t s = (uint32_t)(VX_MAX_SCALE - 1 - scale);
        return (s < levels ? s : levels - 1) * stride + threadIdx.x;
    }
};

// Shared-memory scratch of one CTA: traversal stacks, the unorm8 -> float table and per-thread cold state.
struct Smem {
    Stack stack;
    const float* unorm;   // unorm[b] = b / 255.0f (IEEE division done once per CTA instead of once per channel fetch)
    float* cold;          // COLD_WORDS floats per thread, [k * blockDim.x + tid]
};
#define VX_COLD_WORDS 8
__host__ __device__ inline size_t smem_bytes(uint32_t stack_levels, uint32_t threads) {
    return ((size_t)3 * stack_levels * threads + 256 + (size_t)VX_COLD_WORDS * threads) * 4;
}
__device__ __forceinline__ Smem make_smem(const Scene& s, uint32_t* base) {
    Smem m;
    const uint32_t n = blockDim.x;
    m.stack.stride = n; m.stack.levels = s.stack_levels;
    m.stack.rec = base;
    m.stack.desc = base + (size_t)s.stack_levels * n;
    m.stack.t_max = reinterpret_cast<float*>(base + 2 * (size_t)s.stack_levels * n);
    float* lut = reinterpret_cast<float*>(base + 3 * (size_t)s.stack_levels * n);
    for (uint32_t i = threadIdx.x; i < 256; i += n) lut[i] = (float)i / 255.0f;
    m.unorm = lut;
    m.cold = lut + 256;
    __syncthreads();
    return m;
}

enum : int { RAY_CONTINUE = 0, RAY_LEAF = 1, RAY_MISS = 2 };

// Per-ray registers.
struct Ray {
    float rox, roy, roz;      // origin in [1,2) space
    float rdx, rdy, rdz;      // direction (epsilon-clamped)
    float tcx, tcy, tcz;      // t_coef
    float tbx, tby, tbz;      // t_bias
    float px, py, pz;         // pos
    float t_min, t_max, h;
    float scale_exp2;
    float max_dst;            // already scaled; < 0 = unlimited
    int scale;
    int idx;                  // bits 0-2: idx, bits 4-6: octant_mask
    int steps;
    uint32_t rec, desc;
    uint32_t last_leaf_value;
    int adjacent_leaf_count;
    uint32_t inside_voxel;
};

// Geometry of a leaf candidate (svo.esvo.glsl:190-224, 233)
struct Leaf {
    uint32_t value;
    int face_id;
    float u, v;
    float dst;
    float qx, qy, qz, se;   // un-mirrored voxel corner and edge length in [1,2) space
};

__device__ __forceinline__ uint32_t ld_desc(const Scene& s, uint32_t i) {
    i = i < s.desc_words ? i : s.desc_words - 1;   // robust-buffer-access style clamp
    return __ldg(s.desc + i);
}

// 128-bit read-only load of 4 consecutive descriptor words; i must be 4-word aligned relative to a
// record start (ptr % 4 == 1), which holds for every octant record the reference serializer emits.
__device__ __forceinline__ uint4 ld_desc4(const Scene& s, uint32_t i) {
    i = (i + 3 < s.desc_words) ? i : s.desc_words - 4;
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(s.desc + i));
    return r;
}

__device__ __forceinline__ uint32_t sel4(const uint4& v, uint32_t k) {
    uint32_t lo = (k & 1) ? v.y : v.x, hi = (k & 1) ? v.w : v.z;
    return (k & 2) ? hi : lo;
}

// Child masks + child record pointer of child `oi` of the octant whose record is `rec`
// (= the shader's descriptors[ptr + idx/2] field and get_octant_ptr, svo.esvo.glsl:9-16,168-171).
template <bool VEC>
__device__ __forceinline__ void fetch_child(const Scene& s, uint32_t rec, uint32_t oi, uint32_t& out_desc, uint32_t& out_rec) {
    uint32_t wh, wb;
    if (VEC && (rec & 3u) == 1u) {
        uint4 hdr = ld_desc4(s, rec);
        uint4 body = ld_desc4(s, rec + 4 + (oi & 4u));
        wh = sel4(hdr, oi >> 1);
        wb = sel4(body, oi & 3u);
    } else {
        wh = ld_desc(s, rec + (oi >> 1));
        wb = ld_desc(s, rec + 4 + oi);
    }
    out_desc = (oi & 1u) ? (wh >> 16) : (wh & 0xffffu);
    out_rec = (wb & 0x80000000u) ? (rec + 4 + oi + (wb & 0x7fffffffu)) : wb;
}

// ---------------------------------------------------------------- textures --

__device__ __forceinline__ int ifloor_clamped(float x) {
    x = gl_min(gl_max(x, -16777216.0f), 16777216.0f);
    if (!(x == x)) return 0;
    return (int)floorf(x);
}
__device__ __forceinline__ int imod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }
__device__ __forceinline__ int iclamp(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }

__device__ __forceinline__ float4 fetch_texel(const TexInfo* ti, const float* unorm, uint32_t level, uint32_t wl, uint32_t hl, int layer, int i, int j) {
    const uint32_t t = __ldg(ti->texels + ti->off[level] + ((uint32_t)layer * hl + (uint32_t)j) * wl + (uint32_t)i);
    return make_float4(unorm[t & 0xffu], unorm[(t >> 8) & 0xffu], unorm[(t >> 16) & 0xffu], unorm[t >> 24]);
}

__device__ __forceinline__ float4 sample_linear(const TexInfo* ti, const float* unorm, uint32_t level, int layer, float u, float v) {
    int wl = (int)(ti->w >> level), hl = (int)(ti->h >> level);
    wl = wl ? wl : 1; hl = hl ? hl : 1;
    const float uu = u * (float)wl - 0.5f, vv = v * (float)hl - 0.5f;
    const int i0 = ifloor_clamped(uu), j0 = ifloor_clamped(vv);
    float a = uu - floorf(uu), b = vv - floorf(vv);
    if (!(a == a)) a = 0.0f;
    if (!(b == b)) b = 0.0f;
    const int i0c = iclamp(i0, 0, wl - 1), i1c = iclamp(i0 + 1, 0, wl - 1);   // WRAP_S clamp
    const int j0w = imod(j0, hl), j1w = imod(j0 + 1, hl);                      // WRAP_T repeat
    const float4 t00 = fetch_texel(ti, unorm, level, wl, hl, layer, i0c, j0w), t10 = fetch_texel(ti, unorm, level, wl, hl, layer, i1c, j0w);
    const float4 t01 = fetch_texel(ti, unorm, level, wl, hl, layer, i0c, j1w), t11 = fetch_texel(ti, unorm, level, wl, hl, layer, i1c, j1w);
    return make_float4(mixf(mixf(t00.x, t10.x, a), mixf(t01.x, t11.x, a), b), mixf(mixf(t00.y, t10.y, a), mixf(t01.y, t11.y, a), b),
                       mixf(mixf(t00.z, t10.z, a), mixf(t01.z, t11.z, a), b), mixf(mixf(t00.w, t10.w, a), mixf(t01.w, t11.w, a), b));
}

// textureLod with MIN=LINEAR_MIPMAP_LINEAR, MAG=NEAREST: lod <= 0 -> NEAREST level 0, else trilinear.
// Deliberately ONE out-of-line copy: it is called a handful of times per pixel (leaf colour, shadow alpha, normal
// map) against hundreds of traversal steps, and inlining it three times is what blew the instruction cache.
// Returns the number of texels read in *fetches (added).
__device__ __noinline__ float4 texture_lod(const TexInfo* ti, const float* unorm, float u, float v, int tex_id, float lod, uint32_t* fetches) {
    const int layer = iclamp(tex_id, 0, (int)ti->layers - 1);
    if (!(lod > 0.0f)) {
        const int i = iclamp(ifloor_clamped(u * (float)ti->w), 0, (int)ti->w - 1);
        const int j = imod(ifloor_clamped(v * (float)ti->h), (int)ti->h);
        *fetches += 1;
        return fetch_texel(ti, unorm, 0, ti->w, ti->h, layer, i, j);
    }
    const float maxl = (float)(ti->levels - 1);
    const float l = gl_min(lod, maxl);
    const float fl = floorf(l);
    const uint32_t d1 = (uint32_t)fl;
    const uint32_t d2 = (d1 + 1 < ti->levels) ? d1 + 1 : ti->levels - 1;
    const float f = l - fl;
    const float4 c1 = sample_linear(ti, unorm, d1, layer, u, v);
    *fetches += 4;
    if (d2 == d1 || f == 0.0f) return c1;
    const float4 c2 = sample_linear(ti, unorm, d2, layer, u, v);
    *fetches += 4;
    return make_float4(mixf(c1.x, c2.x, f), mixf(c1.y, c2.y, f), mixf(c1.z, c2.z, f), mixf(c1.w, c2.w, f));
}

// -------------------------------------------------------------- traversal --

// svo.esvo.glsl:52-149. (ox,oy,oz) in SVO voxel space.
__device__ __forceinline__ void ray_init(Ray& r, const Scene& s, float octree_scale, float ox, float oy, float oz, float dx, float dy,
                                         float dz, float max_dst) {
    r.rox = ox * octree_scale + 1.0f; r.roy = oy * octree_scale + 1.0f; r.roz = oz * octree_scale + 1.0f;
    r.max_dst = max_dst * octree_scale;

    const int sign_mask = (int)0x80000000u;
    const int eps_bits = __float_as_int(VX_EPSILON) & ~sign_mask;
    if (fabsf(dx) < VX_EPSILON) dx = __int_as_float(eps_bits | (__float_as_int(dx) & sign_mask));
    if (fabsf(dy) < VX_EPSILON) dy = __int_as_float(eps_bits | (__float_as_int(dy) & sign_mask));
    if (fabsf(dz) < VX_EPSILON) dz = __int_as_float(eps_bits | (__float_as_int(dz) & sign_mask));
    r.rdx = dx; r.rdy = dy; r.rdz = dz;

    r.tcx = 1.0f / -fabsf(dx); r.tcy = 1.0f / -fabsf(dy); r.tcz = 1.0f / -fabsf(dz);
    r.tbx = r.tcx * r.rox; r.tby = r.tcy * r.roy; r.tbz = r.tcz * r.roz;

    int octant_mask = 0;
    if (dx > 0) { octant_mask ^= 1; r.tbx = 3.0f * r.tcx - r.tbx; }
    if (dy > 0) { octant_mask ^= 2; r.tby = 3.0f * r.tcy - r.tby; }
    if (dz > 0) { octant_mask ^= 4; r.tbz = 3.0f * r.tcz - r.tbz; }

    const float t_min = tmax2(tmax2(2.0f * r.tcx - r.tbx, 2.0f * r.tcy - r.tby), 2.0f * r.tcz - r.tbz);
    r.t_min = tmax2(0.0f, t_min);
    r.t_max = tmin2(tmin2(r.tcx - r.tbx, r.tcy - r.tby), r.tcz - r.tbz);
    r.h = r.t_max;

    int idx = 0;
    r.px = 1.0f; r.py = 1.0f; r.pz = 1.0f;
    if (r.t_min < 1.5f * r.tcx - r.tbx) { idx ^= 1; r.px = 1.5f; }
    if (r.t_min < 1.5f * r.tcy - r.tby) { idx ^= 2; r.py = 1.5f; }
    if (r.t_min < 1.5f * r.tcz - r.tbz) { idx ^= 4; r.pz = 1.5f; }
    r.idx = idx | (octant_mask << 4);

    r.scale = VX_MAX_SCALE - 1;
    r.scale_exp2 = 0.5f;
    r.steps = 0;
    r.last_leaf_value = 0xffffffffu;
    r.adjacent_leaf_count = 0;
    r.inside_voxel = 0;

    // state (ptr=0, parent_octant_idx=0): the preamble's child 0 = world root (esvo.rs:179-188)
    const uint32_t w0 = ld_desc(s, 0), w4 = ld_desc(s, 4);
    r.desc = w0 & 0xffffu;
    r.rec = (w4 & 0x80000000u) ? (4u + (w4 & 0x7fffffffu)) : w4;
}

// One iteration of the loop at svo.esvo.glsl:152-392, minus the leaf evaluation.
//   returns RAY_LEAF when the current child is a leaf with t_min > 0 (:185): the caller evaluates it with leaf_geom()
//   and either finishes the ray or calls ray_step again with after_leaf = true, which resumes THAT iteration at its
//   ADVANCE phase (:324) after the bookkeeping of a rejected translucent leaf (:264-265).
//   LIMITED: the ray has a max_dst (:153); render rays do not.
template <bool LIMITED, bool VEC, bool COUNT>
__device__ __forceinline__ int ray_step(Ray& r, const Scene& s, const Stack& st, Counters& cnt, bool after_leaf = false) {
    if (!after_leaf) {
        if (LIMITED && r.max_dst >= 0.0f && r.t_min > r.max_dst) return RAY_MISS;   // :153
        if (r.steps >= VX_MAX_STEPS) return RAY_MISS;                                // :152
        r.steps++;
        if (COUNT) cnt.steps++;
    }
    const float tcornx = __fmaf_rn(r.px, r.tcx, -r.tbx), tcorny = __fmaf_rn(r.py, r.tcy, -r.tby), tcornz = __fmaf_rn(r.pz, r.tcz, -r.tbz);   // :159
    const float tc_max = tmin2(tmin2(tcornx, tcorny), tcornz);                                                                                // :161

    if (!after_leaf) {
        const uint32_t octant_idx = (uint32_t)((r.idx ^ (r.idx >> 4)) & 7);   // :164  idx ^ octant_mask
        const uint32_t bit = 1u << octant_idx;
        const bool is_child = (r.desc & (bit << 8)) != 0;                     // :172
        const bool is_leaf = (r.desc & bit) != 0;                             // :173

        if (is_child && r.t_min <= r.t_max) {                                 // :178
            if (is_leaf) {
                if (r.t_min > 0.0f) return RAY_LEAF;                          // :185
                if (r.t_min == 0.0f) r.inside_voxel = 1;                      // :180
            }
            // :266-312 — also taken by a leaf at t_min == 0 (origin inside a voxel), see SURVEY Appendix B
            const float half_scale = r.scale_exp2 * 0.5f;                     // :274
            const float tv_max = tmin2(r.t_max, tc_max);                      // :278
            if (r.t_min <= tv_max) {                                          // :280  PUSH
                if (COUNT) cnt.pushes++;
                if (tc_max < r.h) {                                           // :284-288
                    const uint32_t sl = st.slot(r.scale);
                    st.rec[sl] = r.rec; st.desc[sl] = r.desc; st.t_max[sl] = r.t_max;
                }
                r.h = tc_max;                                                 // :289
                uint32_t nd, nr;
                fetch_child<VEC>(s, r.rec, octant_idx, nd, nr);               // :292 (+ the :168 read of the next iterations)
                r.rec = nr; r.desc = nd;
                const float tcx_ = __fmaf_rn(half_scale, r.tcx, tcornx), tcy_ = __fmaf_rn(half_scale, r.tcy, tcorny),
                            tcz_ = __fmaf_rn(half_scale, r.tcz, tcornz);      // :275
                --r.scale; r.scale_exp2 = half_scale;                         // :295-297
                int idx = 0;                                                  // :301-304
                if (r.t_min < tcx_) { idx ^= 1; r.px += half_scale; }
                if (r.t_min < tcy_) { idx ^= 2; r.py += half_scale; }
                if (r.t_min < tcz_) { idx ^= 4; r.pz += half_scale; }
                r.idx = (r.idx & 0x70) | idx;
                r.t_max = tv_max;                                             // :307
                return RAY_CONTINUE;                                          // :310
            }
        } else {
            r.adjacent_leaf_count = 0;                                        // :315-316
            r.last_leaf_value = 0xffffffffu;
        }
    }

    int step_mask = 0;                                                        // :324-327  ADVANCE
    if (tc_max >= tcornx) { step_mask ^= 1; r.px -= r.scale_exp2; }
    if (tc_max >= tcorny) { step_mask ^= 2; r.py -= r.scale_exp2; }
    if (tc_max >= tcornz) { step_mask ^= 4; r.pz -= r.scale_exp2; }
    r.t_min = tc_max;                                                         // :330
    r.idx ^= step_mask;                                                       // :331

    if ((r.idx & step_mask) != 0) {                                           // :335  POP
        uint32_t differing_bits = 0;                                          // :347-350
        if (step_mask & 1) differing_bits |= __float_as_uint(r.px) ^ __float_as_uint(r.px + r.scale_exp2);
        if (step_mask & 2) differing_bits |= __float_as_uint(r.py) ^ __float_as_uint(r.py + r.scale_exp2);
        if (step_mask & 4) differing_bits |= __float_as_uint(r.pz) ^ __float_as_uint(r.pz + r.scale_exp2);
        r.scale = 31 - __clz(differing_bits);                                 // :360 findMSB
        r.scale_exp2 = __int_as_float((r.scale - VX_MAX_SCALE + 127) << 23);  // :361 exp2(scale - 23)
        if (r.scale >= VX_MAX_SCALE) return RAY_MISS;                         // :365
        const uint32_t sl = st.slot(r.scale);                                 // :370-372
        r.rec = st.rec[sl]; r.desc = st.desc[sl]; r.t_max = st.t_max[sl];
        const int shx = __float_as_int(r.px) >> r.scale, shy = __float_as_int(r.py) >> r.scale, shz = __float_as_int(r.pz) >> r.scale;   // :377-382
        r.px = __int_as_float(shx << r.scale); r.py = __int_as_float(shy << r.scale); r.pz = __int_as_float(shz << r.scale);
        r.idx = (r.idx & 0x70) | (shx & 1) | ((shy & 1) << 1) | ((shz & 1) << 2);   // :388
        r.h = 0.0f;                                                           // :390
    }
    return RAY_CONTINUE;
}

// HIT block geometry, svo.esvo.glsl:190-224 + :233, for the leaf candidate ray_step stopped at.
template <bool COUNT>
__device__ __forceinline__ void leaf_geom(const Ray& r, const Scene& s, float inv_octree_scale, Leaf& g, Counters& cnt) {
    if (COUNT) cnt.leaf_tests++;
    const int octant_mask = r.idx >> 4;
    const uint32_t octant_idx = (uint32_t)((r.idx ^ octant_mask) & 7);
    g.value = ld_desc(s, r.rec + 4 + octant_idx);                             // :190-194
    const float se = r.scale_exp2;
    const float tnx = __fmaf_rn(r.px + se, r.tcx, -r.tbx), tny = __fmaf_rn(r.py + se, r.tcy, -r.tby), tnz = __fmaf_rn(r.pz + se, r.tcz, -r.tbz);   // :197
    const float tc_min = tmax2(tmax2(tnx, tny), tnz);                         // :199
    float qx = r.px, qy = r.py, qz = r.pz;                                    // :202-205
    if (octant_mask & 1) qx = 3.0f - se - qx;
    if (octant_mask & 2) qy = 3.0f - se - qy;
    if (octant_mask & 4) qz = 3.0f - se - qz;
    const float inv_se = 1.0f / se;                                           // exact: se is a power of two
    if (tc_min == tnx) {                                                      // :210-224
        g.face_id = (__float_as_int(r.rdx) >> 31) & 1;
        g.u = ((r.roz + r.rdz * tnx) - qz) * inv_se; g.v = ((r.roy + r.rdy * tnx) - qy) * inv_se;
        if (r.rdx > 0) g.u = 1 - g.u;
    } else if (tc_min == tny) {
        g.face_id = 2 | ((__float_as_int(r.rdy) >> 31) & 1);
        g.u = ((r.rox + r.rdx * tny) - qx) * inv_se; g.v = ((r.roz + r.rdz * tny) - qz) * inv_se;
        if (r.rdy > 0) g.v = 1 - g.v;
    } else {
        g.face_id = 4 | ((__float_as_int(r.rdz) >> 31) & 1);
        g.u = ((r.rox + r.rdx * tnz) - qx) * inv_se; g.v = ((r.roy + r.rdy * tnz) - qy) * inv_se;
        if (r.rdz < 0) g.u = 1 - g.u;
    }
    g.dst = r.t_min * inv_octree_scale;                                       // :233 (exact: scale is a power of two)
    g.qx = qx; g.qy = qy; g.qz = qz; g.se = se;
}

// res.pos, svo.esvo.glsl:252-258
__device__ __forceinline__ void leaf_pos(const Ray& r, const Leaf& g, float inv_octree_scale, float& x, float& y, float& z) {
    const float hx = gl_min(gl_max(r.rox + r.t_min * r.rdx, g.qx + VX_EPSILON), g.qx + g.se - VX_EPSILON);
    const float hy = gl_min(gl_max(r.roy + r.t_min * r.rdy, g.qy + VX_EPSILON), g.qy + g.se - VX_EPSILON);
    const float hz = gl_min(gl_max(r.roz + r.t_min * r.rdz, g.qz + VX_EPSILON), g.qz + g.se - VX_EPSILON);
    x = (hx - 1.0f) * inv_octree_scale; y = (hy - 1.0f) * inv_octree_scale; z = (hz - 1.0f) * inv_octree_scale;
}

// Material texture for the hit face + custom LOD, svo.esvo.glsl:227-235
__device__ __forceinline__ void leaf_texture(const Scene& s, const Leaf& g, int& tex_id, float& tex_lod) {
    const Material* m = s.materials + (g.value < s.n_materials ? g.value : s.n_materials - 1);
    tex_id = __ldg(&m->tex_side);
    if (g.face_id == 3) tex_id = __ldg(&m->tex_top);
    else if (g.face_id == 2) tex_id = __ldg(&m->tex_bottom);
    float sm = gl_clamp((g.dst - 15.0f) / (25.0f - 15.0f), 0.0f, 1.0f);
    sm = (sm * sm) * (3.0f - 2.0f * sm);
    tex_lod = (sm * (g.dst - 15.0f)) * 0.05f;
}

__device__ __forceinline__ bool layer_is_opaque(const TexInfo* ti, int tex_id) {
    const int layer = iclamp(tex_id, 0, (int)__ldg(&ti->layers) - 1);
    return layer < 64 && ((__ldg(&ti->opaque_layers) >> layer) & 1ull);
}

// ---------------------------------------------------------------- shading --

struct RenderUniforms {   // world.glsl:12-25, view already inverted by the host (svo.rs:197)
    float view[16];
    float tan_half_fov;     // tan(u_fovy * 0.5), hoisted to the host
    float aspect;
    float ambient;
    float lx, ly, lz;       // u_light_dir
    float cx, cy, cz;       // u_cam_pos
    float hx, hy, hz;       // u_highlight_pos
    uint32_t render_shadows;
    float shadow_distance;
    uint32_t width, height;
};

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) { return (ax * bx + ay * by) + az * bz; }

// world.glsl:112-129: pixel -> primary ray (no half-pixel offset; row 0 = bottom)
__device__ __forceinline__ void primary_ray(const RenderUniforms& u, uint32_t gx, uint32_t gy, float& ox, float& oy, float& oz, float& dx,
                                            float& dy, float& dz) {
    float uvx = (float)gx / (float)u.width, uvy = (float)gy / (float)u.height;
    uvx = uvx * 2.0f - 1.0f; uvy = uvy * 2.0f - 1.0f;
    uvx *= u.aspect;
    uvx *= u.tan_half_fov; uvy *= u.tan_half_fov;
    const float* m = u.view;
    const float rw = m[15];
    ox = m[12] / rw; oy = m[13] / rw; oz = m[14] / rw;
    const float lx = ((m[0] * uvx + m[4] * uvy) + m[8] * -1.0f) + m[12];
    const float ly = ((m[1] * uvx + m[5] * uvy) + m[9] * -1.0f) + m[13];
    const float lz = ((m[2] * uvx + m[6] * uvy) + m[10] * -1.0f) + m[14];
    const float lw = ((m[3] * uvx + m[7] * uvy) + m[11] * -1.0f) + m[15];
    const float vx_ = lx / lw - ox, vy_ = ly / lw - oy, vz_ = lz / lw - oz;
    const float l = sqrtf(dot3(vx_, vy_, vz_, vx_, vy_, vz_));
    dx = vx_ / l; dy = vy_ / l; dz = vz_ / l;
}

// world.glsl:92-108
__device__ __forceinline__ float4 sky_color(float dx, float dy, float dz) {
    const float SKYR = 135.0f / 255.0f, SKYG = 206.0f / 255.0f, SKYB = 235.0f / 255.0f;
    const float HR = mixf(1.0f, SKYR, 0.3f), HG = mixf(1.0f, SKYG, 0.3f), HB = mixf(1.0f, SKYB, 0.3f);
    const float pl = sqrtf(dot3(dx, 0.0f, dz, dx, 0.0f, dz));
    const float px = dx / pl, py = 0.0f / pl, pz = dz / pl;
    const float lrd = sqrtf(dot3(dx, dy, dz, dx, dy, dz)), lp = sqrtf(dot3(px, py, pz, px, py, pz));
    // acos is undefined for |x| > 1 in GLSL; the reference's expected image pins 1+ulp (horizon row) to acos(1) = 0
    const float a = acosf(gl_min(dot3(dx, dy, dz, px, py, pz) / fabsf(lrd) * fabsf(lp), 1.0f));
    float grad = a / 1.570796f;
    grad = 1 - powf(1 - grad, 3.0f);
    return make_float4(mixf(HR, SKYR, grad), mixf(HG, SKYG, grad), mixf(HB, SKYB, grad), 1.0f);
}

// What shading leaves pending while the shadow ray is in flight (world.glsl:70-89).
struct Shade {
    float r, g, b, a;       // res.color
    float lit;              // diffuse + specular
    float sox, soy, soz;    // shadow ray origin
    bool done;              // true: (r,g,b,a) is final (highlight outline) — no lighting
    bool want_shadow;
};

// world.glsl:37-84 up to (not including) the shadow ray. (r,g,b,a) of `o` must hold res.color on entry.
__device__ __forceinline__ void shade_hit(const Scene& s, const float* unorm, const RenderUniforms& u, const Leaf& g, float tex_lod, float posx,
                                          float posy, float posz, Shade& o, uint32_t* fetches) {
    o.done = false; o.want_shadow = false;
    if (floorf(posx) == floorf(u.hx) && floorf(posy) == floorf(u.hy) && floorf(posz) == floorf(u.hz)) {   // :37
        const float thickness = 1.0f / 16.0f;
        const float lx = fabsf(g.u - 0.5f) * 2.0f, ly = fabsf(g.v - 0.5f) * 2.0f;
        if (gl_max(lx, ly) > 1.0f - thickness) { o.r = o.g = o.b = o.a = 1.0f; o.done = true; return; }
    }
    const Material* m = s.materials + (g.value < s.n_materials ? g.value : s.n_materials - 1);   // :48
    int tex_normal_id = __ldg(&m->tex_side_normal);
    if (g.face_id == 3) tex_normal_id = __ldg(&m->tex_top_normal);
    else if (g.face_id == 2) tex_normal_id = __ldg(&m->tex_bottom_normal);

    // FACE_NORMALS / FACE_TANGENTS / FACE_BITANGENTS (svo.glsl:2-29) from the face id
    const int axis = g.face_id >> 1;
    const float sgn = (g.face_id & 1) ? 1.0f : -1.0f;
    float nx = axis == 0 ? sgn : 0.0f, ny = axis == 1 ? sgn : 0.0f, nz = axis == 2 ? sgn : 0.0f;
    float tx, ty = 0.0f, tz, bx = 0.0f, by, bz;
    if (axis == 0) { tx = 0.0f; tz = -sgn; by = 1.0f; bz = 0.0f; }          // x-: (0,0,1)  x+: (0,0,-1); bitangent (0,1,0)
    else if (axis == 1) { tx = 1.0f; tz = 0.0f; by = 0.0f; bz = 1.0f; }     // y: tangent (1,0,0), bitangent (0,0,1)
    else { tx = sgn; tz = 0.0f; by = 1.0f; bz = 0.0f; }                     // z-: (-1,0,0) z+: (1,0,0); bitangent (0,1,0)

    if (tex_normal_id != -1) {                                       // :59-67
        const float4 t = texture_lod(s.tex, unorm, g.u, g.v, tex_normal_id, tex_lod, fetches);
        float ex = t.x * 2 - 1, ey = t.z * 2 - 1, ez = t.y * 2 - 1;  // .xzy
        const float l = sqrtf(dot3(ex, ey, ez, ex, ey, ez));
        ex = ex / l; ey = ey / l; ez = ez / l;
        const float nnx = (ex * tx + ey * nx) + ez * bx, nny = (ex * ty + ey * ny) + ez * by, nnz = (ex * tz + ey * nz) + ez * bz;
        nx = nnx; ny = nny; nz = nnz;
    }
    const float ilx = -u.lx, ily = -u.ly, ilz = -u.lz;
    const float dni = dot3(nx, ny, nz, ilx, ily, ilz);
    const float diffuse = gl_max(dni, 0.0f);                         // :70
    float vx_ = posx - u.cx, vy_ = posy - u.cy, vz_ = posz - u.cz;   // :73
    const float vl = sqrtf(dot3(vx_, vy_, vz_, vx_, vy_, vz_));
    vx_ = vx_ / vl; vy_ = vy_ / vl; vz_ = vz_ / vl;
    const float rx = ilx - (2.0f * dni) * nx, ry = ily - (2.0f * dni) * ny, rz = ilz - (2.0f * dni) * nz;   // :74
    const float specular = powf(gl_max(dot3(vx_, vy_, vz_, rx, ry, rz), 0.0f), __ldg(&m->specular_pow)) * __ldg(&m->specular_strength);   // :75
    o.lit = diffuse + specular;
    if (u.render_shadows && g.dst < u.shadow_distance) {             // :80
        o.want_shadow = true;
        o.sox = posx + nx * 0.001f; o.soy = posy + ny * 0.001f; o.soz = posz + nz * 0.001f;   // :82
    }
}

// world.glsl:87-89
__device__ __forceinline__ float4 shade_finish(const RenderUniforms& u, float r, float g, float b, float a, float lit, float shadow) {
    const float light = gl_clamp(u.ambient + lit * shadow, 0.0f, 1.0f);
    return make_float4(r * light, g * light, b * light, a);
}

}  // namespace vx
