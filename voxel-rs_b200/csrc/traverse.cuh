// traverse.cuh — device-side SVO ray traversal (ESVO and CSVO node formats) + shading for sm_100a.
//
// Behavioural contract = voxel-rs shaders (paths relative to tim-oster/voxel-rs):
//   intersect_octree  assets/shaders/svo.esvo.glsl:50-393  (walk_step<VX_FMT_ESVO>)
//   intersect_octree  assets/shaders/svo.csvo.glsl:171-509 (walk_step<VX_FMT_CSVO>, csvo_* node decode :25-150)
//   trace_ray         assets/shaders/world.glsl:27-90
//   get_sky_color     assets/shaders/world.glsl:92-108
//   textureLod state  src/graphics/texture_array.rs:200-203
// Not a transliteration:
//  * the traversal is a per-lane step machine (walk_init / walk_step) that persistent warps drive in
//    lock-step, swapping finished rays for new ones between steps;
//  * walk_step only walks the tree (PUSH / ADVANCE / POP). When it reaches a leaf candidate it parks the
//    ray (state <= ST_LEAF) and the caller evaluates the leaf (value, face, uv, texture, translucency)
//    OUTSIDE the hot loop, where the lanes of a warp that sit on a leaf do it together; a rejected
//    (translucent) leaf finishes the same iteration at its ADVANCE phase (walk_skip_leaf);
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
    const uint32_t* __restrict__ desc;   // descriptors[]: world buffer + 4 bytes (ESVO, svo.esvo.glsl:3-6) / + 8 bytes (CSVO, svo.csvo.glsl:1-5)
    uint32_t desc_words;                 // capacity in words
    uint32_t format;                     // VX_FMT_ESVO / VX_FMT_CSVO (host-side dispatch only; kernels are compiled per format)
    uint32_t max_rec;                    // desc_words - 12: largest record index whose 12 words are inside the buffer
    const Material* __restrict__ materials;
    uint32_t n_materials;
    const TexInfo* __restrict__ tex;
    const float* __restrict__ unorm;     // 256-entry b / 255.0f table in global memory (copied into shared memory per CTA)
    uint32_t stack_levels;               // entries of the per-thread traversal stack (= SVO depth + 1, <= 23)
    uint32_t stack_max_off;              // (stack_levels - 1) * VX_STACK_STRIDE: byte offset of the last stack slot (Walk::soff clamp)
    const uint32_t* __restrict__ bounds; // occupied box of the SVO in voxel units {min x,y,z, max x,y,z (exclusive)}, written by svo_bounds_kernel
                                         // after every commit; null = rays are not clipped against it
    unsigned long long opaque_materials; // bit m set (m < 64): every texel of the three face textures of material m has alpha > 0,
                                         // so a leaf of that material is accepted by the translucency rule (:241-242) without sampling
};

#define VX_FMT_ESVO 0
#define VX_FMT_CSVO 1

struct Counters {   // = VxFrameStats counters
    unsigned long long primary_rays, shadow_rays, steps, pushes, leaf_tests, tex_fetches;
};

// Shared-memory scratch of one CTA of VX_THREADS threads:
//   stack   per-thread traversal stack, one column per thread: slot L word k of thread t at stack[(L*3 + k) * VX_THREADS + t]
//           (k = 0 record, 1 child/leaf masks, 2 t_max). Consecutive lanes hit consecutive banks whatever their slot is.
//   unorm   unorm[b] = b / 255.0f (IEEE division done once per CTA instead of once per channel fetch)
//   cold    VX_COLD_WORDS floats per thread, [k * VX_THREADS + t]: ray origin/direction and other per-ray values that the
//           hot loop never touches
#define VX_THREADS 128
#define VX_COLD_WORDS 8
struct Smem {
    uint32_t stack;        // shared-space byte address of this thread's stack column
    const float* unorm;
    float* cold;
};
// with_lut = false: a kernel that never samples a texture (the picker) leaves the unorm table out — the 1 KB is what separates
// 8 from 9 resident CTAs per SM at the depth-12 world of BASELINE configs[3].
__host__ __device__ inline size_t smem_bytes(uint32_t stack_levels, bool with_stack = true, bool with_lut = true) {
    return ((with_stack ? (size_t)3 * stack_levels * VX_THREADS : 0) + (with_lut ? 256 : 0) + (size_t)VX_COLD_WORDS * VX_THREADS) * 4;
}
__device__ __forceinline__ Smem make_smem(const float* unorm_table, uint32_t stack_levels, uint32_t* base, bool with_stack = true, bool with_lut = true) {
    Smem m;
    const size_t stack_words = with_stack ? (size_t)3 * stack_levels * VX_THREADS : 0;
    m.stack = (uint32_t)__cvta_generic_to_shared(base + threadIdx.x);
    asm volatile("" : "+r"(m.stack));   // opaque from here on: one live register instead of a per-iteration recomputation
    float* lut = reinterpret_cast<float*>(base + stack_words);
    if (with_lut)
        for (uint32_t i = threadIdx.x; i < 256; i += VX_THREADS) lut[i] = __ldg(unorm_table + i);   // (every kernel that calls this runs VX_THREADS threads)
    m.unorm = lut;
    m.cold = lut + (with_lut ? 256 : 0) + threadIdx.x;
    __syncthreads();
    return m;
}

// ---------------------------------------------------------------- textures --

__device__ __forceinline__ int ifloor_clamped(float x) {
    x = gl_min(gl_max(x, -16777216.0f), 16777216.0f);
    if (!(x == x)) return 0;
    return (int)floorf(x);
}
// GL_REPEAT index: a mod n, non-negative. Texture sizes are powers of two in practice (the reference's atlas is 32x32): then it is one AND
// (two's complement makes it right for negative a too) instead of an integer division — 8 of them per trilinear sample.
__device__ __forceinline__ int imod(int a, int n) {
    if ((n & (n - 1)) == 0) return a & (n - 1);
    int r = a % n;
    return r < 0 ? r + n : r;
}
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
// NCH = how many channels the caller uses (4 = rgba, 3 = rgb: the normal map's alpha is never read, world.glsl:60): the
// unused channel's table look-ups and lerps are dead code in that instantiation.
template <int NCH>
__device__ __noinline__ float4 texture_lod(const TexInfo* ti, const float* unorm, float u, float v, int tex_id, float lod, uint32_t* fetches) {
    const int layer = iclamp(tex_id, 0, (int)ti->layers - 1);
    if (!(lod > 0.0f)) {
        const int i = iclamp(ifloor_clamped(u * (float)ti->w), 0, (int)ti->w - 1);
        const int j = imod(ifloor_clamped(v * (float)ti->h), (int)ti->h);
        *fetches += 1;
        float4 r = fetch_texel(ti, unorm, 0, ti->w, ti->h, layer, i, j);
        if (NCH < 4) r.w = 0.0f;
        return r;
    }
    const float maxl = (float)(ti->levels - 1);
    const float l = gl_min(lod, maxl);
    const float fl = floorf(l);
    const uint32_t d1 = (uint32_t)fl;
    const uint32_t d2 = (d1 + 1 < ti->levels) ? d1 + 1 : ti->levels - 1;
    const float f = l - fl;
    float4 c1 = sample_linear(ti, unorm, d1, layer, u, v);
    if (NCH < 4) c1.w = 0.0f;
    *fetches += 4;
    if (d2 == d1 || f == 0.0f) return c1;
    const float4 c2 = sample_linear(ti, unorm, d2, layer, u, v);
    *fetches += 4;
    return make_float4(mixf(c1.x, c2.x, f), mixf(c1.y, c2.y, f), mixf(c1.z, c2.z, f), NCH < 4 ? 0.0f : mixf(c1.w, c2.w, f));
}


// Hot per-ray state of the traversal (registers). The ray's origin/direction are NOT part of it: they are only needed
// when a leaf is evaluated and live in the per-thread cold area of shared memory (Smem::cold, slots 0-5).
struct Walk {
    float tcx, tcy, tcz;      // t_coef
    float tbx, tby, tbz;      // t_bias
    float px, py, pz;         // pos (mirrored space, [1,2))
    float t_min, t_max, h;
    float se;                 // scale_exp2 = 2^(scale-23); the shader's integer `scale` is its exponent (walk_scale)
    float limit;              // max_dst in [1,2) space, +inf when unlimited (:153)
    uint32_t rec, desc;       // ESVO: record of the current octant, child/leaf masks of its children (bits 0-7 leaf, 8-15 child)
                              // CSVO: byte pointer of the current node, remaining depth (the shader's ptr / depth)
    uint32_t hdr;             // CSVO: the node's header (u16 when depth > 3, else u8), re-read whenever (ptr, depth) changes
    uint32_t mat_ptr, preleaf;   // CSVO: material_section_ptr, pre_leaf_pointer (svo.csvo.glsl:222-223; not stacked, like the shader)
    uint32_t ci;              // slot of the current cell in its parent, UN-mirrored: the shader's `idx ^ octant_mask` (:164), 0..7.
                              // Kept instead of idx: the node decode of every iteration uses it as it is, and idx is one xor away.
    uint32_t flags;           // bits 0-2: octant_mask; bit 8: inside_voxel; bit 9: the previous leaf candidate was rejected
                              // (adjacent_leaf_count > 0, its value is in the caller's last_leaf)
    uint32_t soff;            // byte offset of the traversal-stack slot of the current level, = min(22 - scale, levels - 1) * VX_STACK_STRIDE
    int state;                // see ST_*: > 0 walking (iterations left of MAX_STEPS), 0 budget used up, < 0 stopped
};
#define VX_FLAG_INSIDE 0x100u
#define VX_FLAG_ADJACENT 0x200u
#define VX_STACK_STRIDE (3u * VX_THREADS * 4u)   // bytes between two levels of one thread's stack column
__device__ __forceinline__ int walk_scale(const Walk& w) { return (int)((__float_as_uint(w.se) >> 23) & 0xffu) - 104; }   // se = 2^(scale-23)

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

// ------------------------------------------------------------------------------------ CSVO node decode (svo.csvo.glsl) --
// Byte-packed nodes addressed by byte pointers into descriptors[]. A word that is not inside the buffer reads as 0 (the
// robust-access policy the oracle states too: a ray that starts inside a voxel descends through bytes that are not nodes).
// (desc[desc_words] is a zero guard word: the allocation is 64 bytes larger than the capacity and nothing ever writes there,
// so "outside reads 0" is one min instead of a compare, a predicated load and a select.)
__device__ __forceinline__ uint32_t csvo_word(const Scene& s, uint32_t i) { return __ldg(s.desc + min(i, s.desc_words)); }
__device__ __forceinline__ uint32_t csvo_read_uint(const Scene& s, uint32_t ptr) {     // :25-35, = the 32 bits starting at byte ptr
    const uint32_t i = ptr >> 2;
    return __funnelshift_r(csvo_word(s, i), csvo_word(s, i + 1), (ptr & 3u) * 8u);
}
__device__ __forceinline__ uint32_t csvo_read_ushort(const Scene& s, uint32_t ptr) { return csvo_read_uint(s, ptr) & 0xffffu; }   // :39-42
__device__ __forceinline__ uint32_t csvo_read_byte(const Scene& s, uint32_t ptr) { return (csvo_word(s, ptr >> 2) >> ((ptr & 3u) * 8u)) & 0xffu; }   // :45-49
// bitfieldInsert(0u, 0xffffffffu, 0, bits), with the oracle's definition outside [0, 32]
__device__ __forceinline__ uint32_t low_bits(int bits) { return bits <= 0 ? 0u : (bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u)); }

// Header of the node at (ptr, depth): what read_next_ptr looks at first (:57-58, :108-109).
__device__ __forceinline__ uint32_t csvo_header(const Scene& s, uint32_t ptr, uint32_t depth) {
    return csvo_read_uint(s, ptr) & (depth > 3u ? 0xffffu : 0xffu);   // one code path for both header sizes (no divergence between node kinds)
}
// is the child there? (child_mask != 0, :60-62 / :111)
__device__ __forceinline__ bool csvo_has_child(uint32_t hdr, uint32_t depth, uint32_t idx) {
    return depth > 3u ? ((hdr >> (idx * 2u)) & 3u) != 0u : ((hdr >> idx) & 1u) != 0u;
}
// bytes of the pointers whose 2-bit tags are set in `tags` (sum of (1 << tag) >> 1, :68-87): tag 1 -> 1, 2 -> 2, 3 -> 4
__device__ __forceinline__ uint32_t csvo_tag_bytes(uint32_t tags) {
    const uint32_t lo = tags & 0x5555u, hi = (tags >> 1) & 0x5555u;
    return __popc(lo) + 2u * __popc(hi) + __popc(hi & lo);                     // tag 1: 1 + 0 + 0, tag 2: 0 + 2 + 0, tag 3: 1 + 2 + 1
}
// read_next_ptr (:53-133) for a child that is known to be present.
__device__ __forceinline__ uint32_t csvo_next_ptr(const Scene& s, uint32_t ptr, uint32_t depth, uint32_t hdr, uint32_t idx, bool& crossed_boundary) {
    crossed_boundary = false;
    if (depth > 3u) {                                                          // internal nodes
        const uint32_t child_mask = (hdr >> (idx * 2u)) & 3u;
        const uint32_t offset = csvo_tag_bytes(hdr & ((1u << (idx * 2u)) - 1u));
        const uint32_t ptr_bytes = csvo_tag_bytes(hdr);
        uint32_t ptr_offset = csvo_read_uint(s, ptr + 2u + offset);
        ptr_offset &= low_bits((int)(1u << (child_mask - 1u)) * 8);           // the pointer's own 1 / 2 / 4 bytes
        if (ptr_offset & 0x80000000u) { crossed_boundary = true; return ptr_offset ^ 0x80000000u; }   // absolute: a chunk record
        return ptr + 2u + ptr_bytes + ptr_offset;
    }
    const uint32_t offset = __popc(hdr & ((1u << idx) - 1u));
    if (depth == 3u) return ptr + 1u + __popc(hdr) + csvo_read_byte(s, ptr + 1u + offset);   // pre-leaf nodes
    return ptr + 3u + offset;                                                  // leaf nodes: mask + u16 material offset
}
// read_leaf (:136-150)
__device__ __forceinline__ uint32_t csvo_read_leaf(const Scene& s, uint32_t material_section_ptr, uint32_t pre_leaf_ptr, uint32_t ptr, uint32_t idx) {
    const uint32_t material_section_offset = csvo_read_ushort(s, pre_leaf_ptr + 1u);
    const int bit_mark = (int)(ptr - (pre_leaf_ptr + 3u)) * 8 + (int)idx;
    const uint32_t v0 = csvo_read_uint(s, pre_leaf_ptr + 3u) & low_bits(min(bit_mark, 32));
    const uint32_t v1 = csvo_read_uint(s, pre_leaf_ptr + 7u) & low_bits(max(bit_mark - 32, 0));
    return csvo_read_uint(s, material_section_ptr + material_section_offset * 4u + (__popc(v0) + __popc(v1)) * 4u);
}

// Ray state word (Walk::state): > 0 = walking, value = iterations left of the MAX_STEPS budget (:152);
// 0 = budget used up (a miss, :392); ST_MISS = left the octree / beyond max_dst; ST_IDLE = no ray in this lane;
// <= ST_LEAF = stopped at a leaf candidate with (ST_LEAF - state) iterations of budget left.
enum : int { ST_MISS = -2, ST_IDLE = -3, ST_LEAF = -4 };
__device__ __forceinline__ bool state_at_leaf(int st) { return st <= ST_LEAF; }
__device__ __forceinline__ bool state_missed(int st) { return st == 0 || st == ST_MISS; }

// Occupied box of the world in [1,2) space. A ray is over once it has left the box: nothing it could hit lies outside, so the
// shader's remaining iterations (through empty cells to the edge of the octree, :152-392) can only end in a miss — the traversal
// stops there with that miss. Same output bit for bit, fewer iterations (oracle/oracle.cpp restates it to keep the counters
// comparable; vx_set_option 12 turns it off).
struct Clip {
    float lox, loy, loz, hix, hiy, hiz;
    int mode;   // 0 off, 1 box, 2 the world is empty
};
__device__ __forceinline__ Clip load_clip(const Scene& s, float octree_scale) {
    Clip c;
    c.mode = 0;
    c.lox = c.loy = c.loz = c.hix = c.hiy = c.hiz = 0.0f;
    if (s.bounds) {
        const uint32_t x0 = __ldg(s.bounds), y0 = __ldg(s.bounds + 1), z0 = __ldg(s.bounds + 2);
        const uint32_t x1 = __ldg(s.bounds + 3), y1 = __ldg(s.bounds + 4), z1 = __ldg(s.bounds + 5);
        if (x0 >= x1 || y0 >= y1 || z0 >= z1) { c.mode = 2; return c; }
        c.mode = 1;
        c.lox = (float)x0 * octree_scale + 1.0f; c.loy = (float)y0 * octree_scale + 1.0f; c.loz = (float)z0 * octree_scale + 1.0f;
        c.hix = (float)x1 * octree_scale + 1.0f; c.hiy = (float)y1 * octree_scale + 1.0f; c.hiz = (float)z1 * octree_scale + 1.0f;
    }
    return c;
}

// svo.esvo.glsl:52-149 / svo.csvo.glsl:171-223. (ox,oy,oz) in SVO voxel space. Also returns the [1,2)-space origin and the epsilon-clamped
// direction (the leaf evaluation needs them, :210-224, :252-258).
template <int FMT>
__device__ __forceinline__ void walk_init(Walk& w, const Scene& s, const Clip& clip, float octree_scale, float ox, float oy, float oz, float dx, float dy,
                                          float dz, float max_dst, float& rox, float& roy, float& roz, float& rdx, float& rdy, float& rdz) {
    rox = ox * octree_scale + 1.0f; roy = oy * octree_scale + 1.0f; roz = oz * octree_scale + 1.0f;
    const float md = max_dst * octree_scale;
    w.limit = (md >= 0.0f) ? md : __int_as_float(0x7f800000);   // "max_dst >= 0 && t_min > max_dst" == "t_min > limit"

    const int sign_mask = (int)0x80000000u;
    const int eps_bits = __float_as_int(VX_EPSILON) & ~sign_mask;
    if (fabsf(dx) < VX_EPSILON) dx = __int_as_float(eps_bits | (__float_as_int(dx) & sign_mask));
    if (fabsf(dy) < VX_EPSILON) dy = __int_as_float(eps_bits | (__float_as_int(dy) & sign_mask));
    if (fabsf(dz) < VX_EPSILON) dz = __int_as_float(eps_bits | (__float_as_int(dz) & sign_mask));
    rdx = dx; rdy = dy; rdz = dz;

    w.tcx = 1.0f / -fabsf(dx); w.tcy = 1.0f / -fabsf(dy); w.tcz = 1.0f / -fabsf(dz);
    w.tbx = w.tcx * rox; w.tby = w.tcy * roy; w.tbz = w.tcz * roz;

    if (clip.mode == 1) {
        // exit time of the occupied box (slab test; 1 / d_k = -+t_coef_k), widened by 2^-10 relative + 2^-18 absolute: a voxel at
        // the very face of the box is entered no later than the box is left, whatever the rounding of the two
        const float ix = dx > 0 ? -w.tcx : w.tcx, iy = dy > 0 ? -w.tcy : w.tcy, iz = dz > 0 ? -w.tcz : w.tcz;
        const float ex = fmaxf((clip.lox - rox) * ix, (clip.hix - rox) * ix);
        const float ey = fmaxf((clip.loy - roy) * iy, (clip.hiy - roy) * iy);
        const float ez = fmaxf((clip.loz - roz) * iz, (clip.hiz - roz) * iz);
        const float te = fminf(fminf(ex, ey), ez);
        w.limit = fminf(w.limit, te * 1.0009765625f + 3.814697265625e-06f);
    } else if (clip.mode == 2) {
        w.limit = -1.0f;
    }

    uint32_t octant_mask = 0;
    if (dx > 0) { octant_mask ^= 1; w.tbx = 3.0f * w.tcx - w.tbx; }
    if (dy > 0) { octant_mask ^= 2; w.tby = 3.0f * w.tcy - w.tby; }
    if (dz > 0) { octant_mask ^= 4; w.tbz = 3.0f * w.tcz - w.tbz; }

    const float t_min = tmax2(tmax2(2.0f * w.tcx - w.tbx, 2.0f * w.tcy - w.tby), 2.0f * w.tcz - w.tbz);
    w.t_min = tmax2(0.0f, t_min);
    w.t_max = tmin2(tmin2(w.tcx - w.tbx, w.tcy - w.tby), w.tcz - w.tbz);
    w.h = w.t_max;

    uint32_t idx = 0;
    w.px = 1.0f; w.py = 1.0f; w.pz = 1.0f;
    if (w.t_min < 1.5f * w.tcx - w.tbx) { idx ^= 1; w.px = 1.5f; }
    if (w.t_min < 1.5f * w.tcy - w.tby) { idx ^= 2; w.py = 1.5f; }
    if (w.t_min < 1.5f * w.tcz - w.tbz) { idx ^= 4; w.pz = 1.5f; }
    w.ci = idx ^ octant_mask;
    w.flags = octant_mask;

    w.soff = 0;          // scale = MAX_SCALE - 1
    w.se = 0.5f;
    w.state = (w.t_min > w.limit) ? ST_MISS : VX_MAX_STEPS;   // :153 at the first iteration

    if (FMT == VX_FMT_CSVO) {
        w.rec = __ldg(s.desc - 1);                                             // root_ptr, svo.csvo.glsl:190
        w.desc = 127u - ((__float_as_uint(octree_scale) >> 23) & 0xffu);       // depth from the exponent of octree_scale, :221
        w.hdr = csvo_header(s, w.rec, w.desc);
        w.mat_ptr = 0xffffffffu; w.preleaf = 0xffffffffu;                      // INVALID_PTR, :222-223
        return;
    }
    // state (ptr=0, parent_octant_idx=0): the preamble's child 0 = world root (esvo.rs:179-188)
    const uint32_t w0 = ld_desc(s, 0), w4 = ld_desc(s, 4);
    w.desc = w0 & 0xffffu;
    const uint32_t root = (w4 & 0x80000000u) ? (4u + (w4 & 0x7fffffffu)) : w4;
    w.rec = root < s.max_rec ? root : s.max_rec;
}

// Per-thread traversal stack access. `a` is the SHARED-space byte address of the slot: the thread's column (Smem::stack, kept in
// one register and made opaque to the compiler — it otherwise re-derives it from %tid and the CTA's shared window in every loop
// iteration: 6 instructions, 6 % of the trace kernels, profiles/r01_v3_frame_wavefront.md) plus Walk::soff.
// (VX_HOST_EMULATION: tests/emu compiles these headers with g++ and runs the kernels on CPU fibers against the oracle; the PTX is
// then replaced by the same accesses on the emulated shared window. nvcc never defines it.)
__device__ __forceinline__ void stack_store(uint32_t a, uint32_t rec, uint32_t desc, float t_max) {
#ifndef VX_HOST_EMULATION
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(rec) : "memory");
    asm volatile("st.shared.u32 [%0+512], %1;" ::"r"(a), "r"(desc) : "memory");
    asm volatile("st.shared.f32 [%0+1024], %1;" ::"r"(a), "f"(t_max) : "memory");
#else
    vx_emu_shared_u32(a) = rec; vx_emu_shared_u32(a + 512u) = desc; vx_emu_shared_u32(a + 1024u) = __float_as_uint(t_max);
#endif
}
__device__ __forceinline__ void stack_load(uint32_t a, uint32_t& rec, uint32_t& desc, float& t_max) {
#ifndef VX_HOST_EMULATION
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rec) : "r"(a) : "memory");
    asm volatile("ld.shared.u32 %0, [%1+512];" : "=r"(desc) : "r"(a) : "memory");
    asm volatile("ld.shared.f32 %0, [%1+1024];" : "=f"(t_max) : "r"(a) : "memory");
#else
    rec = vx_emu_shared_u32(a); desc = vx_emu_shared_u32(a + 512u); t_max = __uint_as_float(vx_emu_shared_u32(a + 1024u));
#endif
}
static_assert(VX_THREADS == 128, "stack_store/stack_load hard-code the 512-byte word stride of a 128-thread CTA");

// CSVO stack entries carry the node's header next to its depth, so a POP does not re-read it (the shader re-reads the header of
// `ptr` every iteration; the buffer is immutable while a frame is traced). depth is kept as a signed 16-bit value: it is a small
// count, a chunk's lod byte, or — for a ray that started inside a voxel and descends below the leaves — a small negative number.
__device__ __forceinline__ uint32_t csvo_pack_node(uint32_t depth, uint32_t hdr) { return (depth & 0xffffu) | (hdr << 16); }
__device__ __forceinline__ void csvo_unpack_node(Walk& w) {
    const uint32_t packed = w.desc;
    w.hdr = packed >> 16;
    w.desc = (uint32_t)(int32_t)(int16_t)(packed & 0xffffu);
}

// POP (svo.esvo.glsl:335-391 = svo.csvo.glsl:451-507) after an ADVANCE along the axes of step_mask that left the parent.
// Returns false when the ray left the octree (:365); otherwise the node changed: (rec, desc) = the shader's
// (ptr, parent_octant_idx) resp. (ptr, depth) come back from the stack.
template <int FMT>
__device__ __forceinline__ bool walk_pop(Walk& w, uint32_t stk, uint32_t stack_max_off, uint32_t step_mask) {
    uint32_t differing_bits = 0;                                              // :347-350
    if (step_mask & 1) differing_bits |= __float_as_uint(w.px) ^ __float_as_uint(w.px + w.se);
    if (step_mask & 2) differing_bits |= __float_as_uint(w.py) ^ __float_as_uint(w.py + w.se);
    if (step_mask & 4) differing_bits |= __float_as_uint(w.pz) ^ __float_as_uint(w.pz + w.se);
    const int scale = 31 - __clz(differing_bits);                             // :360 findMSB
    w.se = __int_as_float((scale + (127 - VX_MAX_SCALE)) << 23);              // :361 exp2(scale - 23)
    if (scale >= VX_MAX_SCALE) return false;                                  // :365
    w.soff = min((uint32_t)(VX_MAX_SCALE - 1 - scale) * VX_STACK_STRIDE, stack_max_off);   // :370-372
    stack_load(stk + w.soff, w.rec, w.desc, w.t_max);
    const uint32_t keep = 0xffffffffu << scale;                               // :377-382 floor(pos) at the new scale
    const uint32_t bx = __float_as_uint(w.px), by = __float_as_uint(w.py), bz = __float_as_uint(w.pz);
    w.px = __uint_as_float(bx & keep); w.py = __uint_as_float(by & keep); w.pz = __uint_as_float(bz & keep);
    const uint32_t idx = ((bx >> scale) & 1u) | (((by >> scale) & 1u) << 1) | (((bz >> scale) & 1u) << 2);   // :388
    w.ci = idx ^ (w.flags & 7u);
    w.h = 0.0f;                                                               // :390
    if (FMT == VX_FMT_CSVO) csvo_unpack_node(w);
    return true;
}

// ADVANCE / POP (svo.esvo.glsl:324-391 = svo.csvo.glsl:440-507) as the shader writes them. Returns false when the ray left the octree.
template <int FMT, bool LIMITED>
__device__ __forceinline__ bool walk_advance(Walk& w, uint32_t stk, uint32_t stack_max_off, float tcornx, float tcorny, float tcornz, float tc_max) {
    uint32_t step_mask = 0;                                                   // :324-327  ADVANCE
    if (tc_max >= tcornx) { step_mask ^= 1; w.px -= w.se; }
    if (tc_max >= tcorny) { step_mask ^= 2; w.py -= w.se; }
    if (tc_max >= tcornz) { step_mask ^= 4; w.pz -= w.se; }
    w.t_min = tc_max;                                                         // :330
    w.ci ^= step_mask;                                                        // :331
    // :153 "max_dst >= 0 && t_min > max_dst" of the NEXT iteration, tested here where t_min changes (it changes nowhere else: a PUSH
    // iteration does not pay for the test). Same outcome — a miss — and the same iteration count: the shader leaves before it
    // counts the next iteration.
    if (LIMITED && w.t_min > w.limit) return false;
    if (((w.ci ^ w.flags) & step_mask) != 0) return walk_pop<FMT>(w, stk, stack_max_off, step_mask);   // :335 (idx & step_mask)
    return true;
}

// Node decode of the current cell (svo.esvo.glsl:164-173 / svo.csvo.glsl:239-245): is there a child in slot ci, is it a leaf.
template <int FMT>
__device__ __forceinline__ void walk_decode(Walk& w, bool& is_child, bool& is_leaf) {
    if (FMT == VX_FMT_CSVO) {
        is_child = csvo_has_child(w.hdr, w.desc, w.ci);                       // svo.csvo.glsl:239-241
        is_leaf = w.desc < 2u;                                                // (only looked at when is_child)
        if (w.desc == 2u) w.preleaf = w.rec;                                  // :243-245
    } else {
        const uint32_t d = w.desc >> w.ci;                                    // bit 0: is_leaf, bit 8: is_child (:172-173)
        is_child = (d & 0x100u) != 0u;
        is_leaf = (d & 1u) != 0u;
    }
}

// The node part of a PUSH into child slot w.ci of the current node (svo.esvo.glsl:292 + the :168 read of the following iterations /
// svo.csvo.glsl:395-411): (rec, desc[, hdr, mat_ptr]) become the child's.
template <int FMT>
__device__ __forceinline__ void walk_descend(Walk& w, const Scene& s) {
    const uint32_t ci = w.ci;
    if (FMT == VX_FMT_CSVO) {
        bool crossed;
        uint32_t np = csvo_next_ptr(s, w.rec, w.desc, w.hdr, ci, crossed);
        uint32_t nd = w.desc - 1u;                                            // svo.csvo.glsl:401-402 (unsigned: wraps like the shader's uint)
        if (crossed) {                                                        // :404-411 entering a chunk record
            const uint32_t child_lod = csvo_read_byte(s, np);
            const uint32_t material_bytes = csvo_read_uint(s, np + 1u);
            np += 5u;
            w.mat_ptr = np;
            np += material_bytes;
            nd = child_lod;
        }
        w.rec = np; w.desc = nd;
        w.hdr = csvo_header(s, np, nd);
    } else {
#ifdef VX_V_HDR128   // A/B build only (north_star: "child-descriptor fetches are 128-bit vectorised loads"): the record's four header words in
                     // one 16-byte load (records are 16-byte aligned: DESIGN.md §2), then the word of child pair ci / 2 is selected
        const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(s.desc + w.rec));
        const uint32_t wb = __ldg(s.desc + (w.rec + 4u + ci));                // :292 get_octant_ptr
        const uint32_t wh = (ci & 4u) ? ((ci & 2u) ? h4.w : h4.z) : ((ci & 2u) ? h4.y : h4.x);
        w.desc = wh >> ((ci & 1u) << 4);
#else
        const uint32_t wh = __ldg(s.desc + (w.rec + (ci >> 1)));              // child masks of the new octant (the :168 read of later iterations)
        const uint32_t wb = __ldg(s.desc + (w.rec + 4u + ci));                // :292 get_octant_ptr
        w.desc = wh >> ((ci & 1u) << 4);                                      // bits above 15 are never looked at
#endif
        const uint32_t nr = (wb & 0x7fffffffu) + (((int32_t)wb < 0) ? (w.rec + 4u + ci) : 0u);   // relative (bit 31) or absolute
        w.rec = min(nr, s.max_rec);
    }
}

// One iteration of the loop at svo.esvo.glsl:152-392, minus the evaluation of a leaf candidate. Call only with
// w.state > 0. On return w.state is: the remaining budget (keep stepping while > 0), <= ST_LEAF (the current child is a leaf
// with t_min > 0, :185 — the caller evaluates it and either finishes the ray or calls walk_skip_leaf(), the ADVANCE/POP tail of
// THIS iteration, and goes on stepping), or ST_MISS.
// The record index `rec` is clamped once per PUSH to max_rec = capacity - 12 words, so every later access into that
// record (masks, child pointers, leaf values) is in bounds whatever the buffer holds.
// (Measured and dropped, profiles/r02_step_variants.md: stating what PUSH and ADVANCE have in common — three comparisons against
// plane times, three conditional moves of pos, three index bits — once on per-lane selected operands so that it runs for both
// kinds of lanes together. Bit-exact, but 12 % slower: the 32 rays of a warp tile are mostly in the SAME phase, so the extra
// selects were paid far more often than a serialized block was saved.)
template <int FMT, bool LIMITED, bool COUNT>
__device__ __forceinline__ void walk_step(Walk& w, const Scene& s, uint32_t stk, Counters& cnt) {
    --w.state;                                                                // :152 (:153 is tested where t_min changes: walk_init, walk_advance)
    if (COUNT) cnt.steps++;
    const float tcornx = __fmaf_rn(w.px, w.tcx, -w.tbx), tcorny = __fmaf_rn(w.py, w.tcy, -w.tby), tcornz = __fmaf_rn(w.pz, w.tcz, -w.tbz);   // :159
    const float tc_max = tmin2(tmin2(tcornx, tcorny), tcornz);                // :161
    bool is_child, is_leaf;
    walk_decode<FMT>(w, is_child, is_leaf);
    if (is_child && w.t_min <= w.t_max) {                                     // :178
        if (is_leaf) {
            if (w.t_min > 0.0f) { w.state = ST_LEAF - w.state; return; }      // :185
            if (w.t_min == 0.0f) w.flags |= VX_FLAG_INSIDE;                   // :180 inside_voxel
        }
        // :266-312 — also taken by a leaf at t_min == 0 (origin inside a voxel), see SURVEY Appendix B
        const float tv_max = tmin2(w.t_max, tc_max);                          // :278
        if (w.t_min <= tv_max) {                                              // :280  PUSH
            if (COUNT) cnt.pushes++;
            if (tc_max < w.h)                                                 // :284-288
                stack_store(stk + w.soff, w.rec, FMT == VX_FMT_CSVO ? csvo_pack_node(w.desc, w.hdr) : w.desc, w.t_max);
            w.h = tc_max;                                                     // :289
            const float half = w.se * 0.5f;                                   // :274
            const float tcx_ = __fmaf_rn(half, w.tcx, tcornx), tcy_ = __fmaf_rn(half, w.tcy, tcorny), tcz_ = __fmaf_rn(half, w.tcz, tcornz);   // :275
            walk_descend<FMT>(w, s);
            w.soff = min(w.soff + VX_STACK_STRIDE, s.stack_max_off);          // --scale (:295)
            w.se = half;                                                      // :297
            uint32_t idx = 0;                                                 // :301-304
            if (w.t_min < tcx_) { idx ^= 1; w.px += half; }
            if (w.t_min < tcy_) { idx ^= 2; w.py += half; }
            if (w.t_min < tcz_) { idx ^= 4; w.pz += half; }
            w.ci = idx ^ (w.flags & 7u);
            w.t_max = tv_max;                                                 // :307
            return;                                                           // :310
        }
    } else {
        w.flags &= ~VX_FLAG_ADJACENT;                                         // :315-316 (adjacent_leaf_count = 0)
    }
    if (!walk_advance<FMT, LIMITED>(w, stk, s.stack_max_off, tcornx, tcorny, tcornz, tc_max)) w.state = ST_MISS;
}

// ADVANCE/POP tail of the iteration that stopped at a rejected (translucent / repeated) leaf, svo.esvo.glsl:264-265 + :324.
// The rejected leaf used up no extra iteration: the budget stored in the leaf state is restored.
template <int FMT, bool LIMITED>
__device__ __forceinline__ void walk_skip_leaf(Walk& w, const Scene& s, uint32_t stk) {
    const int budget = ST_LEAF - w.state;
    const float tcornx = __fmaf_rn(w.px, w.tcx, -w.tbx), tcorny = __fmaf_rn(w.py, w.tcy, -w.tby), tcornz = __fmaf_rn(w.pz, w.tcz, -w.tbz);
    const float tc_max = tmin2(tmin2(tcornx, tcorny), tcornz);
    w.state = walk_advance<FMT, LIMITED>(w, stk, s.stack_max_off, tcornx, tcorny, tcornz, tc_max) ? budget : ST_MISS;
}

template <int FMT>
__device__ __forceinline__ uint32_t leaf_value(const Walk& w, const Scene& s) {   // svo.esvo.glsl:190-194 / svo.csvo.glsl:259
    if (FMT == VX_FMT_CSVO) return csvo_read_leaf(s, w.mat_ptr, w.preleaf, w.rec, w.ci);
    return __ldg(s.desc + (w.rec + 4u + w.ci));
}

// HIT block geometry, svo.esvo.glsl:197-224 + :233, for the leaf candidate walk_step stopped at.
__device__ __forceinline__ void leaf_geom(const Walk& w, float rox, float roy, float roz, float rdx, float rdy, float rdz, float inv_octree_scale, Leaf& g) {
    const uint32_t octant_mask = w.flags & 7u;
    const float se = w.se;
    const float tnx = __fmaf_rn(w.px + se, w.tcx, -w.tbx), tny = __fmaf_rn(w.py + se, w.tcy, -w.tby), tnz = __fmaf_rn(w.pz + se, w.tcz, -w.tbz);   // :197
    const float tc_min = tmax2(tmax2(tnx, tny), tnz);                         // :199
    float qx = w.px, qy = w.py, qz = w.pz;                                    // :202-205
    if (octant_mask & 1) qx = 3.0f - se - qx;
    if (octant_mask & 2) qy = 3.0f - se - qy;
    if (octant_mask & 4) qz = 3.0f - se - qz;
    const float inv_se = __int_as_float(0x7f000000 - __float_as_int(se));     // 1 / se, exact: se is a power of two in [2^-23, 1/2]
    if (tc_min == tnx) {                                                      // :210-224
        g.face_id = (__float_as_int(rdx) >> 31) & 1;
        g.u = ((roz + rdz * tnx) - qz) * inv_se; g.v = ((roy + rdy * tnx) - qy) * inv_se;
        if (rdx > 0) g.u = 1 - g.u;
    } else if (tc_min == tny) {
        g.face_id = 2 | ((__float_as_int(rdy) >> 31) & 1);
        g.u = ((rox + rdx * tny) - qx) * inv_se; g.v = ((roz + rdz * tny) - qz) * inv_se;
        if (rdy > 0) g.v = 1 - g.v;
    } else {
        g.face_id = 4 | ((__float_as_int(rdz) >> 31) & 1);
        g.u = ((rox + rdx * tnz) - qx) * inv_se; g.v = ((roy + rdy * tnz) - qy) * inv_se;
        if (rdz < 0) g.u = 1 - g.u;
    }
    g.dst = w.t_min * inv_octree_scale;                                       // :233 (exact: scale is a power of two)
    g.qx = qx; g.qy = qy; g.qz = qz; g.se = se;
}

// res.pos, svo.esvo.glsl:252-258
__device__ __forceinline__ void leaf_pos(float t_min, float rox, float roy, float roz, float rdx, float rdy, float rdz, const Leaf& g,
                                         float inv_octree_scale, float& x, float& y, float& z) {
    const float hx = gl_min(gl_max(rox + t_min * rdx, g.qx + VX_EPSILON), g.qx + g.se - VX_EPSILON);
    const float hy = gl_min(gl_max(roy + t_min * rdy, g.qy + VX_EPSILON), g.qy + g.se - VX_EPSILON);
    const float hz = gl_min(gl_max(roz + t_min * rdz, g.qz + VX_EPSILON), g.qz + g.se - VX_EPSILON);
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
    const float lx = ((m[0] * uvx + m[4] * uvy) + m[8] * -1.0f) + m[12];
    const float ly = ((m[1] * uvx + m[5] * uvy) + m[9] * -1.0f) + m[13];
    const float lz = ((m[2] * uvx + m[6] * uvy) + m[10] * -1.0f) + m[14];
    float vx_, vy_, vz_;
    if (m[3] == 0.0f && m[7] == 0.0f && m[11] == 0.0f && m[15] == 1.0f) {
        // affine view matrix (every look_to_rh inverse): both w components are exactly 1 and x / 1 == x, so the six
        // perspective divisions of world.glsl:123-129 are the identity. Warp-uniform branch.
        ox = m[12]; oy = m[13]; oz = m[14];
        vx_ = lx - ox; vy_ = ly - oy; vz_ = lz - oz;
    } else {
        const float rw = m[15];
        ox = m[12] / rw; oy = m[13] / rw; oz = m[14] / rw;
        const float lw = ((m[3] * uvx + m[7] * uvy) + m[11] * -1.0f) + m[15];
        vx_ = lx / lw - ox; vy_ = ly / lw - oy; vz_ = lz / lw - oz;
    }
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
        const float4 t = texture_lod<3>(s.tex, unorm, g.u, g.v, tex_normal_id, tex_lod, fetches);
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
