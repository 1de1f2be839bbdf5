// oracle.cpp — CPU restatement of voxel-rs's GLSL ray-cast path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for the CUDA path in voxel-rs_b200/csrc. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
// product (libvoxelrt.so and the host mirror) never links, imports or executes anything in oracle/.
//
// It follows the reference shaders statement by statement (paths relative to tim-oster/voxel-rs):
//   intersect_octree   assets/shaders/svo.esvo.glsl:50-393   -> vxo::intersect_octree
//   get_octant_ptr     assets/shaders/svo.esvo.glsl:9-16     -> vxo::get_octant_ptr
//   tables/structs     assets/shaders/svo.glsl:2-63          -> FACE_* tables, Material, OctreeResult
//   trace_ray          assets/shaders/world.glsl:27-90       -> vxo::trace_ray
//   get_sky_color      assets/shaders/world.glsl:92-108      -> vxo::get_sky_color
//   main (render)      assets/shaders/world.glsl:110-141     -> vxo_render
//   main (picker)      assets/shaders/picker.glsl:30-51      -> vxo_raycast
//   debug harness      assets/shaders/svo.test.glsl:44-76    -> vxo_debug_cast
//   sampler state      src/graphics/texture_array.rs:200-203 -> vxo::texture_lod
//   glReadPixels       src/graphics/framebuffer.rs:97-105    -> vxo_to_rgba8
//
// Pinning: tests/test_oracle_golden.py checks this oracle against the reference's own golden
// vectors (src/graphics/svo_shader_tests.rs:293-753 step traces and results, src/graphics/svo.rs:
// 402-449 picker rays, the CSVO goldens :756-1224, and BOTH images the reference commits, each under the reference's own
// metric and default 0.1 % threshold: assets/tests/graphics_svo_render_expected.png (close-up scene, NEAREST texels) and
// assets/tests/gamelogic_world_end_to_end_expected.png (the generated world at radius 15: LOD chunks, shadows and the mip-mapped
// trilinear texture path; diff 0.00045, 98.5 % of the pixels within 1 LSB). Third-party arithmetic that is NOT in the reference
// checkout (OpenGL driver: textureLod filtering arithmetic, glGenerateMipmap, GLSL built-ins) is restated from the OpenGL 4.5
// spec (section 8.14) with a 2x2 box-filter mip chain whose rounding is fitted to the second image (build_mips below).
//
// Numeric convention (shared with the CUDA kernels, see DESIGN.md "Numerics"): IEEE-754 binary32,
// round-to-nearest-even, no compiler-chosen contraction (build with -ffp-contract=off); a fused
// multiply-add is used in exactly the three places where the shader's own comment asks for one
// ("calculating the next interception distance is one FMA-operation per axis", svo.esvo.glsl:97-99):
// t_corner (:159), the leaf entry corner (:197) and t_center (:275) — written as explicit fmaf().
// IEEE division and sqrt, min/max as the GLSL spec writes them (min(x,y) = y<x ? y : x; max(x,y) = x<y ? y : x),
// exp2(integer) built from exponent bits, findMSB = 31 - clz.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace vxo {

// ---------------------------------------------------------------- helpers --

static inline float gl_min(float x, float y) { return (y < x) ? y : x; }
static inline float gl_max(float x, float y) { return (x < y) ? y : x; }
static inline float gl_clamp(float x, float lo, float hi) { return gl_min(gl_max(x, lo), hi); }
static inline int32_t f2i(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
static inline uint32_t f2u(float f) { uint32_t i; std::memcpy(&i, &f, 4); return i; }
static inline float i2f(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int find_msb(uint32_t v) { return v == 0 ? -1 : 31 - __builtin_clz(v); }
static inline float exp2i(int e) { return i2f((e + 127) << 23); }  // exact 2^e for -126 <= e <= 127

struct vec3 { float x, y, z; };
static inline vec3 v3(float x, float y, float z) { return vec3{x, y, z}; }
static inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline vec3 normalize(vec3 v) {
    float l = sqrtf(dot(v, v));
    return v3(v.x / l, v.y / l, v.z / l);
}
static inline float length(vec3 v) { return sqrtf(dot(v, v)); }

// svo.glsl:2-29
static const float FACE_NORMALS[6][3]    = {{-1,0,0},{1,0,0},{0,-1,0},{0,1,0},{0,0,-1},{0,0,1}};
static const float FACE_TANGENTS[6][3]   = {{0,0,1},{0,0,-1},{1,0,0},{1,0,0},{-1,0,0},{1,0,0}};
static const float FACE_BITANGENTS[6][3] = {{0,1,0},{0,1,0},{0,0,1},{0,0,1},{0,1,0},{0,1,0}};

// svo.glsl:48-59
struct Material {
    float specular_pow, specular_strength;
    int32_t tex_top, tex_side, tex_bottom;
    int32_t tex_top_normal, tex_side_normal, tex_bottom_normal;
};

// svo.glsl:31-40
struct OctreeResult {
    float t; uint32_t value; int32_t face_id;
    float pos[3]; float uv[2]; float color[4]; float lod; uint32_t inside_voxel;
};

// svo.test.glsl:23-33
struct DebugFrame {
    float t_min; uint32_t ptr, idx, parent_octant_idx; int32_t scale, is_child, is_leaf, crossed_boundary; uint32_t next_ptr;
};

struct Counters {
    uint64_t primary_rays, shadow_rays, steps, pushes, leaf_tests, tex_fetches;
};

// ---------------------------------------------------------------- textures --

// GL_TEXTURE_2D_ARRAY, RGBA8, texture_array.rs:191-236. Level l has (w>>l) x (h>>l) texels.
struct Texture {
    uint32_t w, h, layers, levels;
    std::vector<std::vector<uint8_t>> mips;  // [level][layer*wl*hl*4]
};

// glGenerateMipmap stand-in (texture_array.rs:258-260; the filter is implementation-defined in OpenGL): each texel of level
// l+1 is the mean of the 2x2 block below it, rounded as floor(mean + 1/4), i.e. (sum + 1) >> 2. That rounding is the one that
// reproduces the reference's own golden image of the mip-mapped far field (assets/tests/gamelogic_world_end_to_end_expected.png,
// tests/test_oracle_golden.py::test_world_end_to_end_png): mean |dRGB| 0.00045 with 79 % of the pixels identical and 98.5 %
// within 1 LSB, against 0.0053 with round-to-nearest ((sum + 2) >> 2), 0.0057 with truncation and 0.0028 with round-half-even.
static void build_mips(Texture& t) {
    for (uint32_t l = 1; l < t.levels; ++l) {
        uint32_t pw = t.w >> (l - 1), ph = t.h >> (l - 1);
        uint32_t cw = t.w >> l, ch = t.h >> l;
        if (cw == 0) cw = 1;
        if (ch == 0) ch = 1;
        t.mips[l].assign((size_t)t.layers * cw * ch * 4, 0);
        const std::vector<uint8_t>& src = t.mips[l - 1];
        for (uint32_t layer = 0; layer < t.layers; ++layer)
            for (uint32_t y = 0; y < ch; ++y)
                for (uint32_t x = 0; x < cw; ++x)
                    for (uint32_t c = 0; c < 4; ++c) {
                        uint32_t x0 = 2 * x, x1 = (2 * x + 1 < pw) ? 2 * x + 1 : pw - 1;
                        uint32_t y0 = 2 * y, y1 = (2 * y + 1 < ph) ? 2 * y + 1 : ph - 1;
                        size_t base = (size_t)layer * pw * ph;
                        uint32_t s = src[(base + (size_t)y0 * pw + x0) * 4 + c] + src[(base + (size_t)y0 * pw + x1) * 4 + c] +
                                     src[(base + (size_t)y1 * pw + x0) * 4 + c] + src[(base + (size_t)y1 * pw + x1) * 4 + c];
                        t.mips[l][(((size_t)layer * ch + y) * cw + x) * 4 + c] = (uint8_t)((s + 1) >> 2);
                    }
    }
}

static inline int ifloor_clamped(float x) {
    x = gl_min(gl_max(x, -16777216.0f), 16777216.0f);  // NaN passes through both and is mapped to 0 below
    if (!(x == x)) return 0;
    return (int)floorf(x);
}
static inline int imod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }
static inline int iclamp(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }

static inline void fetch_texel(const Texture& t, uint32_t level, int layer, int i, int j, float out[4], Counters* c) {
    uint32_t wl = t.w >> level, hl = t.h >> level;
    if (wl == 0) wl = 1;
    if (hl == 0) hl = 1;
    const uint8_t* p = &t.mips[level][(((size_t)layer * hl + (uint32_t)j) * wl + (uint32_t)i) * 4];
    for (int k = 0; k < 4; ++k) out[k] = (float)p[k] / 255.0f;
    if (c) c->tex_fetches++;
}

static inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }

static void sample_linear(const Texture& t, uint32_t level, int layer, float u, float v, float out[4], Counters* c) {
    int wl = (int)(t.w >> level), hl = (int)(t.h >> level);
    if (wl == 0) wl = 1;
    if (hl == 0) hl = 1;
    float uu = u * (float)wl - 0.5f, vv = v * (float)hl - 0.5f;
    int i0 = ifloor_clamped(uu), j0 = ifloor_clamped(vv);
    float a = uu - floorf(uu), b = vv - floorf(vv);
    if (!(a == a)) a = 0.0f;
    if (!(b == b)) b = 0.0f;
    int i0c = iclamp(i0, 0, wl - 1), i1c = iclamp(i0 + 1, 0, wl - 1);  // WRAP_S = CLAMP_TO_EDGE
    int j0w = imod(j0, hl), j1w = imod(j0 + 1, hl);                    // WRAP_T = REPEAT (GL default)
    float t00[4], t10[4], t01[4], t11[4];
    fetch_texel(t, level, layer, i0c, j0w, t00, c);
    fetch_texel(t, level, layer, i1c, j0w, t10, c);
    fetch_texel(t, level, layer, i0c, j1w, t01, c);
    fetch_texel(t, level, layer, i1c, j1w, t11, c);
    for (int k = 0; k < 4; ++k) out[k] = mixf(mixf(t00[k], t10[k], a), mixf(t01[k], t11[k], a), b);
}

// textureLod(sampler2DArray, vec3(uv, layer), lod) with MIN=LINEAR_MIPMAP_LINEAR, MAG=NEAREST
// (texture_array.rs:200-203). OpenGL 4.5 section 8.14: with MAG=NEAREST the min/mag switch-over is
// c = 0, so lod <= 0 magnifies (NEAREST on level 0) and lod > 0 minifies (trilinear).
static void texture_lod(const Texture& t, float u, float v, int tex_id, float lod, float out[4], Counters* c) {
    int layer = iclamp(tex_id, 0, (int)t.layers - 1);  // layer = clamp(round(float(id)), 0, d-1)
    if (!(lod > 0.0f)) {
        int i = iclamp(ifloor_clamped(u * (float)t.w), 0, (int)t.w - 1);
        int j = imod(ifloor_clamped(v * (float)t.h), (int)t.h);
        fetch_texel(t, 0, layer, i, j, out, c);
        return;
    }
    float maxl = (float)(t.levels - 1);
    float l = gl_min(lod, maxl);
    float fl = floorf(l);
    uint32_t d1 = (uint32_t)fl;
    uint32_t d2 = (d1 + 1 < t.levels) ? d1 + 1 : t.levels - 1;
    float f = l - fl;
    float c1[4];
    sample_linear(t, d1, layer, u, v, c1, c);
    if (d2 == d1 || f == 0.0f) {
        for (int k = 0; k < 4; ++k) out[k] = c1[k];
        return;
    }
    float c2[4];
    sample_linear(t, d2, layer, u, v, c2, c);
    for (int k = 0; k < 4; ++k) out[k] = mixf(c1[k], c2[k], f);
}

// ------------------------------------------------------------------- scene --

struct Scene {
    const uint8_t* world;        // byte 0 = f32 octree_scale, byte 4.. = descriptors[] (svo.esvo.glsl:3-6)
    uint64_t world_len;
    const Material* materials;
    uint32_t n_materials;
    const Texture* tex;
    int format = 0;              // 0 = ESVO (svo.esvo.glsl), 1 = CSVO (svo.csvo.glsl): the shader's `#define SVO_TYPE`
    // NOT part of the reference: the product's world-box clipping (traverse.cuh Clip), restated so that iteration counters of
    // the two can be compared. 0 = off (the shader as written), 1 = box {clo, chi} in [1,2) space, 2 = the world is empty
    int clip_mode = 0;
    float clo[3] = {0, 0, 0}, chi[3] = {0, 0, 0};
    float octree_scale() const { float f; std::memcpy(&f, world, 4); return f; }
    uint32_t desc(uint32_t i) const { uint32_t v; std::memcpy(&v, world + 4 + (uint64_t)i * 4, 4); return v; }
    // CSVO: `uint root_ptr` at byte 4, `uint descriptors[]` from byte 8 (svo.csvo.glsl:1-5). A word that does not lie
    // completely inside the buffer reads as 0 (robust-buffer-access policy shared with the CUDA kernels: out-of-spec
    // descents — a ray origin inside a voxel — walk through bytes that are not nodes).
    uint32_t csvo_root_ptr() const { uint32_t v; std::memcpy(&v, world + 4, 4); return v; }
    uint32_t csvo_word(uint32_t i) const {
        const uint64_t off = 8 + (uint64_t)i * 4;
        if (off + 4 > world_len) return 0;
        uint32_t v; std::memcpy(&v, world + off, 4); return v;
    }
};

// svo.esvo.glsl:9-16
static inline uint32_t get_octant_ptr(const Scene& s, uint32_t ptr, uint32_t idx) {
    uint32_t next_ptr = s.desc(ptr + 4 + idx);
    if ((next_ptr & (1u << 31)) != 0) next_ptr = ptr + 4 + idx + (next_ptr & 0x7fffffffu);
    return next_ptr;
}

#define MAX_STEPS 1000
#define MAX_SCALE 23
static const float EPSILON = 0.00000011920929f;  // svo.esvo.glsl:24

struct Trace { DebugFrame* frames; uint32_t cap; int32_t stack_ptr; };

static void intersect_octree_csvo(const Scene& s, vec3 ro, vec3 rd, float max_dst, bool cast_translucent, OctreeResult& res, Counters* cnt,
                                  Trace* trace);
static void intersect_octree_esvo(const Scene& s, vec3 ro, vec3 rd, float max_dst, bool cast_translucent, OctreeResult& res, Counters* cnt,
                                  Trace* trace);
// svo.glsl:65-74: the shader includes one of the two traversals by `#define SVO_TYPE`
static void intersect_octree(const Scene& s, vec3 ro, vec3 rd, float max_dst, bool cast_translucent, OctreeResult& res, Counters* cnt,
                             Trace* trace) {
    if (s.format == 1) intersect_octree_csvo(s, ro, rd, max_dst, cast_translucent, res, cnt, trace);
    else intersect_octree_esvo(s, ro, rd, max_dst, cast_translucent, res, cnt, trace);
}

// World-box clipping of the product (NOT in the shader): the ray parameter after which the ray has left the box that holds every
// voxel. Same operations, in the same order, as walk_init in voxel-rs_b200/csrc/traverse.cuh. +inf when clipping is off.
static inline float clip_limit(const Scene& s, vec3 ro, vec3 rd, vec3 t_coef) {
    if (s.clip_mode == 2) return -1.0f;
    if (s.clip_mode != 1) return i2f(0x7f800000);
    const float ix = rd.x > 0 ? -t_coef.x : t_coef.x, iy = rd.y > 0 ? -t_coef.y : t_coef.y, iz = rd.z > 0 ? -t_coef.z : t_coef.z;
    const float ex = fmaxf((s.clo[0] - ro.x) * ix, (s.chi[0] - ro.x) * ix);
    const float ey = fmaxf((s.clo[1] - ro.y) * iy, (s.chi[1] - ro.y) * iy);
    const float ez = fmaxf((s.clo[2] - ro.z) * iz, (s.chi[2] - ro.z) * iz);
    const float te = fminf(fminf(ex, ey), ez);
    return te * 1.0009765625f + 3.814697265625e-06f;
}

// svo.esvo.glsl:50-393. `trace` (optional) reproduces OCTREE_RAYTRACE_DEBUG_FN of svo.test.glsl.
static void intersect_octree_esvo(const Scene& s, vec3 ro, vec3 rd, float max_dst, bool cast_translucent,
                                  OctreeResult& res, Counters* cnt, Trace* trace) {
    const float octree_scale = s.octree_scale();
    uint32_t ptr_stack[MAX_SCALE + 1];
    uint32_t parent_octant_idx_stack[MAX_SCALE + 1];
    float t_max_stack[MAX_SCALE + 1];
    for (int i = 0; i <= MAX_SCALE; ++i) { ptr_stack[i] = 0; parent_octant_idx_stack[i] = 0; t_max_stack[i] = 0; }

    ro.x *= octree_scale; ro.y *= octree_scale; ro.z *= octree_scale;   // :52
    max_dst *= octree_scale;                                            // :53

    res.t = -1; res.value = 0; res.face_id = 0;                         // :56-62
    res.pos[0] = res.pos[1] = res.pos[2] = 0; res.uv[0] = res.uv[1] = 0;
    res.color[0] = res.color[1] = res.color[2] = res.color[3] = 0; res.lod = 0; res.inside_voxel = 0;

    ro.x += 1; ro.y += 1; ro.z += 1;                                    // :66

    uint32_t ptr = 0, parent_octant_idx = 0;                            // :68-69
    int scale = MAX_SCALE - 1;                                          // :74
    float scale_exp2 = 0.5f;                                            // :75
    uint32_t last_leaf_value = 0xffffffffu;                             // :80
    int adjacent_leaf_count = 0;                                        // :81

    const int32_t sign_mask = (int32_t)0x80000000u;                     // :85-89
    const int32_t eps_bits = f2i(EPSILON) & ~sign_mask;
    if (fabsf(rd.x) < EPSILON) rd.x = i2f(eps_bits | (f2i(rd.x) & sign_mask));
    if (fabsf(rd.y) < EPSILON) rd.y = i2f(eps_bits | (f2i(rd.y) & sign_mask));
    if (fabsf(rd.z) < EPSILON) rd.z = i2f(eps_bits | (f2i(rd.z) & sign_mask));

    vec3 t_coef = v3(1.0f / -fabsf(rd.x), 1.0f / -fabsf(rd.y), 1.0f / -fabsf(rd.z));   // :105
    vec3 t_bias = v3(t_coef.x * ro.x, t_coef.y * ro.y, t_coef.z * ro.z);               // :106
    const float clip_t = clip_limit(s, ro, rd, t_coef);                               // (product extension, off by default)

    int octant_mask = 0;                                                // :121-124
    if (rd.x > 0) { octant_mask ^= 1; t_bias.x = 3.0f * t_coef.x - t_bias.x; }
    if (rd.y > 0) { octant_mask ^= 2; t_bias.y = 3.0f * t_coef.y - t_bias.y; }
    if (rd.z > 0) { octant_mask ^= 4; t_bias.z = 3.0f * t_coef.z - t_bias.z; }

    float t_min = gl_max(gl_max(2.0f * t_coef.x - t_bias.x, 2.0f * t_coef.y - t_bias.y), 2.0f * t_coef.z - t_bias.z);  // :129
    t_min = gl_max(0.0f, t_min);                                        // :130
    float t_max = gl_min(gl_min(t_coef.x - t_bias.x, t_coef.y - t_bias.y), t_coef.z - t_bias.z);  // :133
    float h = t_max;                                                    // :134

    int idx = 0;                                                        // :139
    vec3 pos = v3(1.0f, 1.0f, 1.0f);                                    // :142
    if (t_min < 1.5f * t_coef.x - t_bias.x) { idx ^= 1; pos.x = 1.5f; } // :147-149
    if (t_min < 1.5f * t_coef.y - t_bias.y) { idx ^= 2; pos.y = 1.5f; }
    if (t_min < 1.5f * t_coef.z - t_bias.z) { idx ^= 4; pos.z = 1.5f; }

    for (int i = 0; i < MAX_STEPS; ++i) {                               // :152
        if (max_dst >= 0 && t_min > max_dst) return;                    // :153-156
        if (t_min > clip_t) return;                                     // (product extension: the ray has left the occupied box)
        if (cnt) cnt->steps++;

        vec3 t_corner = v3(fmaf(pos.x, t_coef.x, -t_bias.x), fmaf(pos.y, t_coef.y, -t_bias.y), fmaf(pos.z, t_coef.z, -t_bias.z));  // :159 (FMA, :97-99)
        float tc_max = gl_min(gl_min(t_corner.x, t_corner.y), t_corner.z);   // :161

        uint32_t octant_idx = (uint32_t)(idx ^ octant_mask);            // :164
        uint32_t bit = 1u << octant_idx;                                // :165

        uint32_t descriptor = s.desc(ptr + (parent_octant_idx / 2));    // :168
        if ((parent_octant_idx % 2) != 0) descriptor >>= 16;            // :169-171
        bool is_child = (descriptor & (bit << 8)) != 0;                 // :172
        bool is_leaf = (descriptor & bit) != 0;                         // :173

        if (trace) {                                                    // :175, svo.test.glsl:48-59
            trace->stack_ptr += 1;
            if ((uint32_t)trace->stack_ptr < trace->cap) {
                DebugFrame& f = trace->frames[trace->stack_ptr];
                f.t_min = t_min / octree_scale; f.ptr = ptr; f.idx = octant_idx; f.parent_octant_idx = parent_octant_idx;
                f.scale = scale; f.is_child = is_child; f.is_leaf = is_leaf; f.crossed_boundary = 0; f.next_ptr = 0;
            }
        }

        if (is_child && t_min <= t_max) {                               // :178
            if (is_leaf && t_min == 0) res.inside_voxel = 1;            // :180-182

            if (is_leaf && t_min > 0) {                                 // :185  phase: HIT
                if (cnt) cnt->leaf_tests++;
                uint32_t next_ptr = get_octant_ptr(s, ptr, parent_octant_idx);   // :190
                next_ptr = next_ptr + 4 + octant_idx;                   // :191
                uint32_t value = s.desc(next_ptr);                      // :194

                vec3 tcn = v3(fmaf(pos.x + scale_exp2, t_coef.x, -t_bias.x), fmaf(pos.y + scale_exp2, t_coef.y, -t_bias.y),
                              fmaf(pos.z + scale_exp2, t_coef.z, -t_bias.z));    // :197 (FMA)
                float tc_min = gl_max(gl_max(tcn.x, tcn.y), tcn.z);     // :199

                vec3 p = pos;                                           // :202-205
                if ((octant_mask & 1) != 0) p.x = 3.0f - scale_exp2 - p.x;
                if ((octant_mask & 2) != 0) p.y = 3.0f - scale_exp2 - p.y;
                if ((octant_mask & 4) != 0) p.z = 3.0f - scale_exp2 - p.z;

                int face_id; float uvx, uvy;                            // :210-224
                if (tc_min == tcn.x) {
                    face_id = (f2i(rd.x) >> 31) & 1;
                    uvx = ((ro.z + rd.z * tcn.x) - p.z) / scale_exp2; uvy = ((ro.y + rd.y * tcn.x) - p.y) / scale_exp2;
                    if (rd.x > 0) uvx = 1 - uvx;
                } else if (tc_min == tcn.y) {
                    face_id = 2 | ((f2i(rd.y) >> 31) & 1);
                    uvx = ((ro.x + rd.x * tcn.y) - p.x) / scale_exp2; uvy = ((ro.z + rd.z * tcn.y) - p.z) / scale_exp2;
                    if (rd.y > 0) uvy = 1 - uvy;
                } else {
                    face_id = 4 | ((f2i(rd.z) >> 31) & 1);
                    uvx = ((ro.x + rd.x * tcn.z) - p.x) / scale_exp2; uvy = ((ro.y + rd.y * tcn.z) - p.y) / scale_exp2;
                    if (rd.z < 0) uvx = 1 - uvx;
                }

                // :227-230 (value indexes the material SSBO; out-of-range reads are clamped here)
                const Material& mat = s.materials[value < s.n_materials ? value : s.n_materials - 1];
                int tex_id = mat.tex_side;
                if (face_id == 3) tex_id = mat.tex_top;
                else if (face_id == 2) tex_id = mat.tex_bottom;

                float dst = t_min / octree_scale;                       // :233
                float sm = gl_clamp((dst - 15.0f) / (25.0f - 15.0f), 0.0f, 1.0f);   // smoothstep(15,25,dst) :235
                sm = (sm * sm) * (3.0f - 2.0f * sm);
                float tex_lod = (sm * (dst - 15.0f)) * 0.05f;

                float tex_color[4];
                texture_lod(*s.tex, uvx, uvy, tex_id, tex_lod, tex_color, cnt);     // :237

                bool first_of_kind = adjacent_leaf_count == 0 || value != last_leaf_value;   // :241
                if ((tex_color[3] > 0 || !cast_translucent) && first_of_kind) {              // :242
                    res.t = dst; res.face_id = face_id; res.uv[0] = uvx; res.uv[1] = uvy; res.value = value;
                    for (int k = 0; k < 4; ++k) res.color[k] = tex_color[k];
                    res.lod = tex_lod;
                    res.pos[0] = gl_min(gl_max(ro.x + t_min * rd.x, p.x + EPSILON), p.x + scale_exp2 - EPSILON);   // :252-254
                    res.pos[1] = gl_min(gl_max(ro.y + t_min * rd.y, p.y + EPSILON), p.y + scale_exp2 - EPSILON);
                    res.pos[2] = gl_min(gl_max(ro.z + t_min * rd.z, p.z + EPSILON), p.z + scale_exp2 - EPSILON);
                    for (int k = 0; k < 3; ++k) { res.pos[k] -= 1; res.pos[k] /= octree_scale; }   // :257-258
                    return;
                }
                ++adjacent_leaf_count;                                  // :264-265
                last_leaf_value = value;
            } else {
                float half_scale = scale_exp2 * 0.5f;                   // :274
                vec3 t_center = v3(fmaf(half_scale, t_coef.x, t_corner.x), fmaf(half_scale, t_coef.y, t_corner.y),
                                   fmaf(half_scale, t_coef.z, t_corner.z));   // :275 (FMA)
                float tv_max = gl_min(t_max, tc_max);                   // :278
                if (t_min <= tv_max) {                                  // :280  phase: PUSH
                    if (cnt) cnt->pushes++;
                    if (tc_max < h) {                                   // :284-288
                        ptr_stack[scale] = ptr; parent_octant_idx_stack[scale] = parent_octant_idx; t_max_stack[scale] = t_max;
                    }
                    h = tc_max;                                         // :289
                    ptr = get_octant_ptr(s, ptr, parent_octant_idx);    // :292
                    --scale; parent_octant_idx = octant_idx; scale_exp2 = half_scale;   // :295-297
                    idx = 0;                                            // :301-304
                    if (t_min < t_center.x) { idx ^= 1; pos.x += scale_exp2; }
                    if (t_min < t_center.y) { idx ^= 2; pos.y += scale_exp2; }
                    if (t_min < t_center.z) { idx ^= 4; pos.z += scale_exp2; }
                    t_max = tv_max;                                     // :307
                    continue;                                           // :310
                }
            }
        } else {
            adjacent_leaf_count = 0;                                    // :315-316
            last_leaf_value = 0xffffffffu;
        }

        int step_mask = 0;                                              // :324-327  phase: ADVANCE
        if (tc_max >= t_corner.x) { step_mask ^= 1; pos.x -= scale_exp2; }
        if (tc_max >= t_corner.y) { step_mask ^= 2; pos.y -= scale_exp2; }
        if (tc_max >= t_corner.z) { step_mask ^= 4; pos.z -= scale_exp2; }
        t_min = tc_max;                                                 // :330
        idx ^= step_mask;                                               // :331

        if ((idx & step_mask) != 0) {                                   // :335  phase: POP
            uint32_t differing_bits = 0;                                // :347-350
            if ((step_mask & 1) != 0) differing_bits |= f2u(pos.x) ^ f2u(pos.x + scale_exp2);
            if ((step_mask & 2) != 0) differing_bits |= f2u(pos.y) ^ f2u(pos.y + scale_exp2);
            if ((step_mask & 4) != 0) differing_bits |= f2u(pos.z) ^ f2u(pos.z + scale_exp2);
            scale = find_msb(differing_bits);                           // :360
            scale_exp2 = exp2i(scale - MAX_SCALE);                      // :361
            if (scale >= MAX_SCALE) return;                             // :365-367
            ptr = ptr_stack[scale];                                     // :370-372
            parent_octant_idx = parent_octant_idx_stack[scale];
            t_max = t_max_stack[scale];
            int shx = f2i(pos.x) >> scale, shy = f2i(pos.y) >> scale, shz = f2i(pos.z) >> scale;   // :377-382
            pos.x = i2f(shx << scale); pos.y = i2f(shy << scale); pos.z = i2f(shz << scale);
            idx = (shx & 1) | ((shy & 1) << 1) | ((shz & 1) << 2);      // :388
            h = 0;                                                      // :390
        }
    }
}


// ------------------------------------------------------------------------------------------------ CSVO (svo.csvo.glsl) --

// bitfieldInsert(0u, 0xffffffffu, 0, bits): the low `bits` bits set. GLSL leaves bits < 0 / > 32 undefined; here 0 / all.
static inline uint32_t low_bits(int bits) { return bits <= 0 ? 0u : (bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u)); }

// svo.csvo.glsl:25-35
static inline uint32_t csvo_read_uint(const Scene& s, uint32_t ptr) {
    const uint32_t index = ptr / 4, mod = ptr % 4;
    const uint32_t lshift = (4 - mod) * 8;
    const uint32_t mask = low_bits((int)lshift);
    const uint32_t v0 = (s.csvo_word(index) >> (mod * 8)) & mask;
    const uint32_t v1 = lshift >= 32 ? 0u : ((s.csvo_word(index + 1) << lshift) & ~mask);   // `x << 32` is undefined in GLSL; & ~mask is 0 then
    return v0 | v1;
}
static inline uint32_t csvo_read_ushort(const Scene& s, uint32_t ptr) { return csvo_read_uint(s, ptr) & 0xffffu; }   // :39-42
static inline uint32_t csvo_read_byte(const Scene& s, uint32_t ptr) { return (s.csvo_word(ptr / 4) >> ((ptr % 4) * 8)) & 0xffu; }   // :45-49

static const uint32_t INVALID_PTR = 0xffffffffu;

// svo.csvo.glsl:53-133
static uint32_t csvo_read_next_ptr(const Scene& s, uint32_t ptr, uint32_t depth, uint32_t idx, bool& crossed_boundary) {
    crossed_boundary = false;
    if (depth > 3) {                                                    // internal nodes
        const uint32_t header_mask = csvo_read_ushort(s, ptr);
        const uint32_t child_mask = (header_mask >> (idx * 2)) & 3u;
        if (child_mask == 0) return INVALID_PTR;
        const uint32_t offset_mask = (1u << (idx * 2)) - 1u;
        const uint32_t preceding_mask = header_mask & offset_mask;
        uint32_t offset = 0, ptr_bytes = 0;
        for (int k = 0; k < 8; ++k) {
            offset += (1u << ((preceding_mask >> (k * 2)) & 3u)) >> 1;
            ptr_bytes += (1u << ((header_mask >> (k * 2)) & 3u)) >> 1;
        }
        uint32_t ptr_offset = csvo_read_uint(s, ptr + 2 + offset);
        ptr_offset &= low_bits((int)(1u << (child_mask - 1)) * 8);      // keep the pointer's own 1 / 2 / 4 bytes
        if ((ptr_offset & (1u << 31)) != 0) {                           // absolute pointer (32-bit pointers only)
            crossed_boundary = true;
            return ptr_offset ^ (1u << 31);
        }
        return ptr + 2 + ptr_bytes + ptr_offset;
    }
    const uint32_t header_mask = csvo_read_byte(s, ptr);
    const uint32_t child_mask = (header_mask >> idx) & 1u;
    if (child_mask == 0) return INVALID_PTR;
    const uint32_t offset = (uint32_t)__builtin_popcount(header_mask & ((1u << idx) - 1u));
    if (depth == 3) {                                                   // pre-leaf nodes
        const uint32_t ptr_bytes = (uint32_t)__builtin_popcount(header_mask);
        const uint32_t ptr_offset = csvo_read_byte(s, ptr + 1 + offset);
        return ptr + 1 + ptr_bytes + ptr_offset;
    }
    return ptr + 1 + 2 + offset;                                        // leaf nodes: 1-byte mask + 2-byte material offset
}

// svo.csvo.glsl:136-150
static uint32_t csvo_read_leaf(const Scene& s, uint32_t material_section_ptr, uint32_t pre_leaf_ptr, uint32_t ptr, uint32_t idx) {
    const uint32_t material_section_offset = csvo_read_ushort(s, pre_leaf_ptr + 1);
    const int leaf_index = (int)(ptr - (pre_leaf_ptr + 3));
    const int bit_mark = leaf_index * 8 + (int)idx;
    const uint32_t v0 = csvo_read_uint(s, pre_leaf_ptr + 3) & low_bits(bit_mark < 32 ? bit_mark : 32);
    const uint32_t v1 = csvo_read_uint(s, pre_leaf_ptr + 3 + 4) & low_bits(bit_mark - 32 > 0 ? bit_mark - 32 : 0);
    const uint32_t preceding_leaves = (uint32_t)(__builtin_popcount(v0) + __builtin_popcount(v1));
    return csvo_read_uint(s, material_section_ptr + material_section_offset * 4 + preceding_leaves * 4);
}

// svo.csvo.glsl:171-509. Same ray marching as the ESVO variant; node decode, the (ptr, depth) state, the per-chunk material
// section and the absolute-pointer crossing differ. The debug frame carries `depth` in parent_octant_idx (svo.csvo.glsl:246).
// Stack slots: the shader indexes three 24-entry arrays with `scale`; an out-of-spec descent (origin inside a voxel keeps
// PUSHing through bytes that are not nodes) can leave that range, which GLSL leaves undefined. Policy shared with the CUDA
// kernel: slot = min(22 - scale, levels - 1) with levels = depth + 3.
static void intersect_octree_csvo(const Scene& s, vec3 ro, vec3 rd, float max_dst, bool cast_translucent, OctreeResult& res, Counters* cnt,
                                  Trace* trace) {
    const float octree_scale = s.octree_scale();
    uint32_t ptr_stack[MAX_SCALE + 1], depth_stack[MAX_SCALE + 1];
    float t_max_stack[MAX_SCALE + 1];
    for (int i = 0; i <= MAX_SCALE; ++i) { ptr_stack[i] = 0; depth_stack[i] = 0; t_max_stack[i] = 0; }

    ro.x *= octree_scale; ro.y *= octree_scale; ro.z *= octree_scale;   // :173
    max_dst *= octree_scale;                                            // :174
    res.t = -1; res.value = 0; res.face_id = 0;                         // :177-183
    res.pos[0] = res.pos[1] = res.pos[2] = 0; res.uv[0] = res.uv[1] = 0;
    res.color[0] = res.color[1] = res.color[2] = res.color[3] = 0; res.lod = 0; res.inside_voxel = 0;
    ro.x += 1; ro.y += 1; ro.z += 1;                                    // :187

    uint32_t ptr = s.csvo_root_ptr();                                   // :190
    int scale = MAX_SCALE - 1;                                          // :195
    float scale_exp2 = 0.5f;
    uint32_t last_leaf_value = 0xffffffffu;                             // :201-202
    int adjacent_leaf_count = 0;

    const int32_t sign_mask = (int32_t)0x80000000u;                     // :206-210
    const int32_t eps_bits = f2i(EPSILON) & ~sign_mask;
    if (fabsf(rd.x) < EPSILON) rd.x = i2f(eps_bits | (f2i(rd.x) & sign_mask));
    if (fabsf(rd.y) < EPSILON) rd.y = i2f(eps_bits | (f2i(rd.y) & sign_mask));
    if (fabsf(rd.z) < EPSILON) rd.z = i2f(eps_bits | (f2i(rd.z) & sign_mask));

    vec3 t_coef = v3(1.0f / -fabsf(rd.x), 1.0f / -fabsf(rd.y), 1.0f / -fabsf(rd.z));   // :226
    vec3 t_bias = v3(t_coef.x * ro.x, t_coef.y * ro.y, t_coef.z * ro.z);
    const float clip_t = clip_limit(s, ro, rd, t_coef);                               // (product extension, off by default)
    int octant_mask = 0;                                                // :242-245
    if (rd.x > 0) { octant_mask ^= 1; t_bias.x = 3.0f * t_coef.x - t_bias.x; }
    if (rd.y > 0) { octant_mask ^= 2; t_bias.y = 3.0f * t_coef.y - t_bias.y; }
    if (rd.z > 0) { octant_mask ^= 4; t_bias.z = 3.0f * t_coef.z - t_bias.z; }
    float t_min = gl_max(gl_max(2.0f * t_coef.x - t_bias.x, 2.0f * t_coef.y - t_bias.y), 2.0f * t_coef.z - t_bias.z);
    t_min = gl_max(0.0f, t_min);
    float t_max = gl_min(gl_min(t_coef.x - t_bias.x, t_coef.y - t_bias.y), t_coef.z - t_bias.z);
    float h = t_max;
    int idx = 0;
    vec3 pos = v3(1.0f, 1.0f, 1.0f);
    if (t_min < 1.5f * t_coef.x - t_bias.x) { idx ^= 1; pos.x = 1.5f; }
    if (t_min < 1.5f * t_coef.y - t_bias.y) { idx ^= 2; pos.y = 1.5f; }
    if (t_min < 1.5f * t_coef.z - t_bias.z) { idx ^= 4; pos.z = 1.5f; }

    uint32_t depth = 127 - ((f2u(octree_scale) >> 23) & 0xff);          // max depth from the exponent of the scale
    const int levels = (int)depth + 3 < MAX_SCALE + 1 ? (int)depth + 3 : MAX_SCALE + 1;
    auto slot = [levels](int sc) { int l = MAX_SCALE - 1 - sc; return l < 0 ? 0 : (l < levels ? l : levels - 1); };
    uint32_t material_section_ptr = INVALID_PTR;
    uint32_t pre_leaf_pointer = INVALID_PTR;

    for (int i = 0; i < MAX_STEPS; ++i) {
        if (max_dst >= 0 && t_min > max_dst) return;
        if (t_min > clip_t) return;                                     // (product extension)
        if (cnt) cnt->steps++;

        vec3 t_corner = v3(fmaf(pos.x, t_coef.x, -t_bias.x), fmaf(pos.y, t_coef.y, -t_bias.y), fmaf(pos.z, t_coef.z, -t_bias.z));
        float tc_max = gl_min(gl_min(t_corner.x, t_corner.y), t_corner.z);
        uint32_t octant_idx = (uint32_t)(idx ^ octant_mask);

        bool crossed_boundary = false;
        uint32_t next_ptr = csvo_read_next_ptr(s, ptr, depth, octant_idx, crossed_boundary);
        bool is_child = next_ptr != INVALID_PTR;
        bool is_leaf = is_child && depth < 2;
        if (depth == 2) pre_leaf_pointer = ptr;

        if (trace) {
            trace->stack_ptr += 1;
            if ((uint32_t)trace->stack_ptr < trace->cap) {
                DebugFrame& f = trace->frames[trace->stack_ptr];
                f.t_min = t_min / octree_scale; f.ptr = ptr; f.idx = octant_idx; f.parent_octant_idx = depth;
                f.scale = scale; f.is_child = is_child; f.is_leaf = is_leaf; f.crossed_boundary = crossed_boundary; f.next_ptr = next_ptr;
            }
        }

        if (is_child && t_min <= t_max) {
            if (is_leaf && t_min == 0) res.inside_voxel = 1;
            if (is_leaf && t_min > 0) {                                 // phase: HIT
                if (cnt) cnt->leaf_tests++;
                uint32_t value = csvo_read_leaf(s, material_section_ptr, pre_leaf_pointer, ptr, octant_idx);
                vec3 tcn = v3(fmaf(pos.x + scale_exp2, t_coef.x, -t_bias.x), fmaf(pos.y + scale_exp2, t_coef.y, -t_bias.y),
                              fmaf(pos.z + scale_exp2, t_coef.z, -t_bias.z));
                float tc_min = gl_max(gl_max(tcn.x, tcn.y), tcn.z);
                vec3 p = pos;
                if ((octant_mask & 1) != 0) p.x = 3.0f - scale_exp2 - p.x;
                if ((octant_mask & 2) != 0) p.y = 3.0f - scale_exp2 - p.y;
                if ((octant_mask & 4) != 0) p.z = 3.0f - scale_exp2 - p.z;
                int face_id; float uvx, uvy;
                if (tc_min == tcn.x) {
                    face_id = (f2i(rd.x) >> 31) & 1;
                    uvx = ((ro.z + rd.z * tcn.x) - p.z) / scale_exp2; uvy = ((ro.y + rd.y * tcn.x) - p.y) / scale_exp2;
                    if (rd.x > 0) uvx = 1 - uvx;
                } else if (tc_min == tcn.y) {
                    face_id = 2 | ((f2i(rd.y) >> 31) & 1);
                    uvx = ((ro.x + rd.x * tcn.y) - p.x) / scale_exp2; uvy = ((ro.z + rd.z * tcn.y) - p.z) / scale_exp2;
                    if (rd.y > 0) uvy = 1 - uvy;
                } else {
                    face_id = 4 | ((f2i(rd.z) >> 31) & 1);
                    uvx = ((ro.x + rd.x * tcn.z) - p.x) / scale_exp2; uvy = ((ro.y + rd.y * tcn.z) - p.y) / scale_exp2;
                    if (rd.z < 0) uvx = 1 - uvx;
                }
                const Material& mat = s.materials[value < s.n_materials ? value : s.n_materials - 1];
                int tex_id = mat.tex_side;
                if (face_id == 3) tex_id = mat.tex_top;
                else if (face_id == 2) tex_id = mat.tex_bottom;
                float dst = t_min / octree_scale;
                float sm = gl_clamp((dst - 15.0f) / (25.0f - 15.0f), 0.0f, 1.0f);
                sm = (sm * sm) * (3.0f - 2.0f * sm);
                float tex_lod = (sm * (dst - 15.0f)) * 0.05f;
                float tex_color[4];
                texture_lod(*s.tex, uvx, uvy, tex_id, tex_lod, tex_color, cnt);
                bool first_of_kind = adjacent_leaf_count == 0 || value != last_leaf_value;
                if ((tex_color[3] > 0 || !cast_translucent) && first_of_kind) {
                    res.t = dst; res.face_id = face_id; res.uv[0] = uvx; res.uv[1] = uvy; res.value = value;
                    for (int k = 0; k < 4; ++k) res.color[k] = tex_color[k];
                    res.lod = tex_lod;
                    res.pos[0] = gl_min(gl_max(ro.x + t_min * rd.x, p.x + EPSILON), p.x + scale_exp2 - EPSILON);
                    res.pos[1] = gl_min(gl_max(ro.y + t_min * rd.y, p.y + EPSILON), p.y + scale_exp2 - EPSILON);
                    res.pos[2] = gl_min(gl_max(ro.z + t_min * rd.z, p.z + EPSILON), p.z + scale_exp2 - EPSILON);
                    for (int k = 0; k < 3; ++k) { res.pos[k] -= 1; res.pos[k] /= octree_scale; }
                    return;
                }
                ++adjacent_leaf_count;
                last_leaf_value = value;
            } else {
                float half_scale = scale_exp2 * 0.5f;
                vec3 t_center = v3(fmaf(half_scale, t_coef.x, t_corner.x), fmaf(half_scale, t_coef.y, t_corner.y),
                                   fmaf(half_scale, t_coef.z, t_corner.z));
                float tv_max = gl_min(t_max, tc_max);
                if (t_min <= tv_max) {                                  // phase: PUSH
                    if (cnt) cnt->pushes++;
                    if (tc_max < h) { ptr_stack[slot(scale)] = ptr; depth_stack[slot(scale)] = depth; t_max_stack[slot(scale)] = t_max; }
                    h = tc_max;
                    --depth;
                    ptr = next_ptr;
                    if (crossed_boundary) {                             // entering a chunk record: [lod][material bytes][materials][nodes]
                        uint32_t child_lod = csvo_read_byte(s, ptr);
                        uint32_t material_bytes = csvo_read_uint(s, ptr + 1);
                        ptr += 5;
                        material_section_ptr = ptr;
                        ptr += material_bytes;
                        depth = child_lod;
                    }
                    --scale;
                    scale_exp2 = half_scale;
                    idx = 0;
                    if (t_min < t_center.x) { idx ^= 1; pos.x += scale_exp2; }
                    if (t_min < t_center.y) { idx ^= 2; pos.y += scale_exp2; }
                    if (t_min < t_center.z) { idx ^= 4; pos.z += scale_exp2; }
                    t_max = tv_max;
                    continue;
                }
            }
        } else {
            adjacent_leaf_count = 0;
            last_leaf_value = 0xffffffffu;
        }

        int step_mask = 0;                                              // phase: ADVANCE
        if (tc_max >= t_corner.x) { step_mask ^= 1; pos.x -= scale_exp2; }
        if (tc_max >= t_corner.y) { step_mask ^= 2; pos.y -= scale_exp2; }
        if (tc_max >= t_corner.z) { step_mask ^= 4; pos.z -= scale_exp2; }
        t_min = tc_max;
        idx ^= step_mask;
        if ((idx & step_mask) != 0) {                                   // phase: POP
            uint32_t differing_bits = 0;
            if ((step_mask & 1) != 0) differing_bits |= f2u(pos.x) ^ f2u(pos.x + scale_exp2);
            if ((step_mask & 2) != 0) differing_bits |= f2u(pos.y) ^ f2u(pos.y + scale_exp2);
            if ((step_mask & 4) != 0) differing_bits |= f2u(pos.z) ^ f2u(pos.z + scale_exp2);
            scale = find_msb(differing_bits);
            scale_exp2 = exp2i(scale - MAX_SCALE);
            if (scale >= MAX_SCALE) return;
            ptr = ptr_stack[slot(scale)];
            depth = depth_stack[slot(scale)];
            t_max = t_max_stack[slot(scale)];
            int shx = f2i(pos.x) >> scale, shy = f2i(pos.y) >> scale, shz = f2i(pos.z) >> scale;
            pos.x = i2f(shx << scale); pos.y = i2f(shy << scale); pos.z = i2f(shz << scale);
            idx = (shx & 1) | ((shy & 1) << 1) | ((shz & 1) << 2);
            h = 0;
        }
    }
}

struct RenderParams {
    float view[16];
    float fov_y_rad, aspect_ratio, ambient_intensity;
    float light_dir[3], cam_pos[3], highlight_pos[3];
    uint32_t render_shadows;
    float shadow_distance;
};

// world.glsl:27-90
static void trace_ray(const Scene& s, const RenderParams& u, vec3 ro, vec3 rd, bool& hit, float out[4], Counters* cnt) {
    OctreeResult res;
    intersect_octree(s, ro, rd, -1.0f, true, res, cnt, nullptr);        // :29
    hit = res.t != -1.0f;                                               // :31
    if (res.t < 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }   // :33-36
    if (floorf(res.pos[0]) == floorf(u.highlight_pos[0]) && floorf(res.pos[1]) == floorf(u.highlight_pos[1]) &&
        floorf(res.pos[2]) == floorf(u.highlight_pos[2])) {             // :37
        const float thickness = 1.0f / 16.0f;
        float lx = fabsf(res.uv[0] - 0.5f) * 2.0f, ly = fabsf(res.uv[1] - 0.5f) * 2.0f;
        float lmax = gl_max(lx, ly);
        if (lmax > 1.0f - thickness) { out[0] = out[1] = out[2] = out[3] = 1; return; }   // :42-44
    }
    const Material& mat = s.materials[res.value < s.n_materials ? res.value : s.n_materials - 1];   // :48
    int tex_normal_id = mat.tex_side_normal;                            // :49-51
    if (res.face_id == 3) tex_normal_id = mat.tex_top_normal;
    else if (res.face_id == 2) tex_normal_id = mat.tex_bottom_normal;

    vec3 normal = v3(FACE_NORMALS[res.face_id][0], FACE_NORMALS[res.face_id][1], FACE_NORMALS[res.face_id][2]);   // :54-56
    vec3 tangent = v3(FACE_TANGENTS[res.face_id][0], FACE_TANGENTS[res.face_id][1], FACE_TANGENTS[res.face_id][2]);
    vec3 bitangent = v3(FACE_BITANGENTS[res.face_id][0], FACE_BITANGENTS[res.face_id][1], FACE_BITANGENTS[res.face_id][2]);

    if (tex_normal_id != -1) {                                          // :59
        float tx[4];
        texture_lod(*s.tex, res.uv[0], res.uv[1], tex_normal_id, res.lod, tx, cnt);
        vec3 tex = v3(tx[0], tx[2], tx[1]);                             // .xzy :60
        tex = normalize(v3(tex.x * 2 - 1, tex.y * 2 - 1, tex.z * 2 - 1));   // :63
        normal = v3((tex.x * tangent.x + tex.y * normal.x) + tex.z * bitangent.x,   // :66
                    (tex.x * tangent.y + tex.y * normal.y) + tex.z * bitangent.y,
                    (tex.x * tangent.z + tex.y * normal.z) + tex.z * bitangent.z);
    }

    vec3 nl = v3(-u.light_dir[0], -u.light_dir[1], -u.light_dir[2]);
    float diffuse = gl_max(dot(normal, nl), 0.0f);                      // :70

    vec3 view_dir = normalize(v3(res.pos[0] - u.cam_pos[0], res.pos[1] - u.cam_pos[1], res.pos[2] - u.cam_pos[2]));   // :73
    float dni = dot(normal, nl);                                        // reflect(I,N) = I - 2*dot(N,I)*N  :74
    vec3 reflect_dir = v3(nl.x - (2.0f * dni) * normal.x, nl.y - (2.0f * dni) * normal.y, nl.z - (2.0f * dni) * normal.z);
    float specular = powf(gl_max(dot(view_dir, reflect_dir), 0.0f), mat.specular_pow) * mat.specular_strength;   // :75

    float shadow = 1;                                                   // :79-84
    if (u.render_shadows && res.t < u.shadow_distance) {
        OctreeResult sres;
        if (cnt) cnt->shadow_rays++;
        intersect_octree(s, v3(res.pos[0] + normal.x * 0.001f, res.pos[1] + normal.y * 0.001f, res.pos[2] + normal.z * 0.001f),
                         nl, -1.0f, true, sres, cnt, nullptr);
        shadow = sres.t < 0 ? 1.0f : 0.0f;
    }
    float light = gl_clamp(u.ambient_intensity + (diffuse + specular) * shadow, 0.0f, 1.0f);   // :87
    out[0] = res.color[0] * light; out[1] = res.color[1] * light; out[2] = res.color[2] * light; out[3] = res.color[3];   // :88-89
}

// world.glsl:92-108
static void get_sky_color(vec3 rd, float out[3]) {
    const float SKY[3] = {135.0f / 255.0f, 206.0f / 255.0f, 235.0f / 255.0f};
    float HORIZON[3];
    for (int k = 0; k < 3; ++k) HORIZON[k] = mixf(1.0f, SKY[k], 0.3f);
    vec3 p = normalize(v3(rd.x, 0.0f, rd.z));                           // :97
    // :98. GLSL leaves acos undefined for |x| > 1, and rounding pushes the argument to 1+ulp on the horizon row;
    // the reference's expected PNG (graphics_svo_render_expected.png, row 244) pins the outcome to acos(1) = 0.
    float a = acosf(gl_min(dot(rd, p) / fabsf(length(rd)) * fabsf(length(p)), 1.0f));
    float grad = a / 1.570796f;                                         // :101
    grad = 1 - powf(1 - grad, 3.0f);                                    // :104
    for (int k = 0; k < 3; ++k) out[k] = mixf(HORIZON[k], SKY[k], grad);   // :107
}

// world.glsl:110-141 for one invocation. tan_half_fov = tan(u_fovy * 0.5), hoisted by the caller.
static void render_pixel(const Scene& s, const RenderParams& u, float tan_half_fov, uint32_t gx, uint32_t gy, uint32_t w, uint32_t h,
                         float out[4], Counters* cnt) {
    float uvx = (float)gx / (float)w, uvy = (float)gy / (float)h;       // :112
    uvx = uvx * 2.0f - 1.0f; uvy = uvy * 2.0f - 1.0f;                   // :113
    uvx *= u.aspect_ratio;                                              // :114
    uvx *= tan_half_fov; uvy *= tan_half_fov;                           // :115
    const float* m = u.view;                                            // column-major
    // ro_view = u_view * vec4(0,0,0,1); ro = xyz / w  (:121-122)
    float rw = m[15];
    vec3 ro = v3(m[12] / rw, m[13] / rw, m[14] / rw);
    // look_at_view = u_view * vec4(uv, -1, 1)  (:125-126); sum order: ((c0*x + c1*y) + c2*z) + c3*w
    float lx = ((m[0] * uvx + m[4] * uvy) + m[8] * -1.0f) + m[12];
    float ly = ((m[1] * uvx + m[5] * uvy) + m[9] * -1.0f) + m[13];
    float lz = ((m[2] * uvx + m[6] * uvy) + m[10] * -1.0f) + m[14];
    float lw = ((m[3] * uvx + m[7] * uvy) + m[11] * -1.0f) + m[15];
    vec3 look_at = v3(lx / lw, ly / lw, lz / lw);
    vec3 rd = normalize(v3(look_at.x - ro.x, look_at.y - ro.y, look_at.z - ro.z));   // :129
    bool hit = false;
    if (cnt) cnt->primary_rays++;
    trace_ray(s, u, ro, rd, hit, out, cnt);                             // :132
    if (!hit) {                                                         // :135-138
        float sky[3];
        get_sky_color(rd, sky);
        out[0] = sky[0]; out[1] = sky[1]; out[2] = sky[2]; out[3] = 1.0f;
    }
}

}  // namespace vxo

// =========================================================== extern "C" API ==

extern "C" {

struct VxoTexture { vxo::Texture t; };

// TextureArrayBuilder::build (texture_array.rs:83-153): rgba8 = level-0 images already v-flipped.
VxoTexture* vxo_texture_create(const uint8_t* rgba8, uint32_t w, uint32_t h, uint32_t layers, uint32_t mip_levels) {
    VxoTexture* x = new VxoTexture();
    uint32_t m = w < h ? w : h, il = 0;
    while ((m >> (il + 1)) != 0) ++il;                                  // ilog2(min(w,h))  texture_array.rs:105
    uint32_t levels = mip_levels < il ? mip_levels : il;
    if (levels < 1) levels = 1;
    x->t.w = w; x->t.h = h; x->t.layers = layers; x->t.levels = levels;
    x->t.mips.resize(levels);
    x->t.mips[0].assign(rgba8, rgba8 + (size_t)w * h * layers * 4);
    vxo::build_mips(x->t);
    return x;
}
void vxo_texture_destroy(VxoTexture* t) { delete t; }
uint32_t vxo_texture_levels(const VxoTexture* t) { return t->t.levels; }
// copies level `level` (all layers) into out; returns byte count
uint64_t vxo_texture_level(const VxoTexture* t, uint32_t level, uint8_t* out) {
    const std::vector<uint8_t>& v = t->t.mips[level];
    if (out) std::memcpy(out, v.data(), v.size());
    return v.size();
}

// Occupied box of the world at the granularity of octree level L = min(depth, 6): an independent (recursive, CPU) statement of what
// svo_bounds_kernel computes on the device. b = {min xyz, max xyz (exclusive)} in voxel units; min > max = nothing found.
static void bounds_mark(uint32_t b[6], uint32_t depth, uint32_t level, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t cell = 1u << (depth - level);
    const uint32_t c[3] = {x * cell, y * cell, z * cell};
    for (int k = 0; k < 3; ++k) { if (c[k] < b[k]) b[k] = c[k]; if (c[k] + cell > b[3 + k]) b[3 + k] = c[k] + cell; }
}
static void bounds_walk_esvo(const vxo::Scene& s, uint32_t ptr, uint32_t pidx, uint32_t level, uint32_t x, uint32_t y, uint32_t z, uint32_t depth,
                             uint32_t L, uint32_t b[6]) {
    uint32_t d = s.desc(ptr + pidx / 2);
    if (pidx % 2) d >>= 16;
    const uint32_t child_mask = (d >> 8) & 0xffu, leaf_mask = d & 0xffu;
    for (uint32_t i = 0; i < 8; ++i) {
        if (!((child_mask >> i) & 1u)) continue;
        const uint32_t cx = x * 2 + (i & 1u), cy = y * 2 + ((i >> 1) & 1u), cz = z * 2 + ((i >> 2) & 1u);
        if (((leaf_mask >> i) & 1u) || level + 1 == L) { bounds_mark(b, depth, level + 1, cx, cy, cz); continue; }
        bounds_walk_esvo(s, vxo::get_octant_ptr(s, ptr, pidx), i, level + 1, cx, cy, cz, depth, L, b);
    }
}
static void bounds_walk_csvo(const vxo::Scene& s, uint32_t ptr, uint32_t node_depth, uint32_t level, uint32_t x, uint32_t y, uint32_t z,
                             uint32_t depth, uint32_t L, uint32_t b[6]) {
    for (uint32_t i = 0; i < 8; ++i) {
        bool crossed = false;
        uint32_t next = vxo::csvo_read_next_ptr(s, ptr, node_depth, i, crossed);
        if (next == vxo::INVALID_PTR) continue;
        const uint32_t cx = x * 2 + (i & 1u), cy = y * 2 + ((i >> 1) & 1u), cz = z * 2 + ((i >> 2) & 1u);
        if (node_depth < 2 || level + 1 == L) { bounds_mark(b, depth, level + 1, cx, cy, cz); continue; }
        uint32_t nd = node_depth - 1;
        if (crossed) {                                                  // svo.csvo.glsl:404-411
            nd = vxo::csvo_read_byte(s, next);
            next += 5 + vxo::csvo_read_uint(s, next + 1);
        }
        bounds_walk_csvo(s, next, nd, level + 1, cx, cy, cz, depth, L, b);
    }
}
static int g_clip = 0;
void vxo_set_clip(int on) { g_clip = on; }

static vxo::Scene make_scene(const uint8_t* world, uint64_t world_len, const void* materials, uint32_t n_materials, const VxoTexture* tex,
                             int svo_format) {
    vxo::Scene s;
    s.world = world; s.world_len = world_len;
    s.materials = (const vxo::Material*)materials; s.n_materials = n_materials; s.tex = &tex->t;
    s.format = svo_format;
    if (g_clip) {
        const float scale = s.octree_scale();
        uint32_t bits; std::memcpy(&bits, &scale, 4);
        const uint32_t depth = 127u - ((bits >> 23) & 0xffu);
        uint32_t b[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0};
        if (depth >= 1 && depth <= 23) {
            const uint32_t L = depth < 6 ? depth : 6;
            if (svo_format == 1) bounds_walk_csvo(s, s.csvo_root_ptr(), depth, 0, 0, 0, 0, depth, L, b);
            else bounds_walk_esvo(s, 0, 0, 0, 0, 0, 0, depth, L, b);
            s.clip_mode = (b[0] >= b[3] || b[1] >= b[4] || b[2] >= b[5]) ? 2 : 1;
            for (int k = 0; k < 3; ++k) { s.clo[k] = (float)b[k] * scale + 1.0f; s.chi[k] = (float)b[3 + k] * scale + 1.0f; }
        }
    }
    return s;
}

// svo.test.glsl main()
void vxo_debug_cast(const uint8_t* world, uint64_t world_len, const void* materials, uint32_t n_materials, const VxoTexture* tex, int svo_format,
                    const float pos[3], const float dir[3], float max_dst, uint32_t cast_translucent,
                    vxo::OctreeResult* result, vxo::DebugFrame* frames, uint32_t frames_cap, uint32_t* n_frames) {
    vxo::Scene s = make_scene(world, world_len, materials, n_materials, tex, svo_format);
    s.clip_mode = 0;   // svo.test.glsl: every iteration of the shader, whatever vxo_set_clip says
    vxo::Trace tr{frames, frames_cap, -1};
    vxo::intersect_octree(s, vxo::v3(pos[0], pos[1], pos[2]), vxo::v3(dir[0], dir[1], dir[2]), max_dst, cast_translucent != 0,
                          *result, nullptr, &tr);
    if (n_frames) *n_frames = (uint32_t)(tr.stack_ptr + 1);
}

struct VxoTask { float max_dst, _p0[3], pos[3], _p1, dir[3], _p2; };
struct VxoResult { float dst; uint32_t inside_voxel; float _p0[2], pos[3], _p1, normal[3], _p2; };

// picker.glsl main() for tasks [0,n). threads<=0 -> all cores.
void vxo_raycast(const uint8_t* world, uint64_t world_len, const void* materials, uint32_t n_materials, const VxoTexture* tex, int svo_format,
                 const VxoTask* tasks, uint64_t n, VxoResult* results, vxo::Counters* counters, int threads) {
    vxo::Scene s = make_scene(world, world_len, materials, n_materials, tex, svo_format);
    vxo::Counters total{};
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        vxo::Counters c{};
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const VxoTask& t = tasks[i];
            vxo::OctreeResult res;
            c.primary_rays++;
            vxo::intersect_octree(s, vxo::v3(t.pos[0], t.pos[1], t.pos[2]), vxo::v3(t.dir[0], t.dir[1], t.dir[2]), t.max_dst, false,
                                  res, &c, nullptr);                    // picker.glsl:37
            VxoResult& r = results[i];
            std::memset(&r, 0, sizeof(r));
            if (res.t > 0) {                                            // picker.glsl:40-44
                r.dst = res.t; r.inside_voxel = res.inside_voxel;
                for (int k = 0; k < 3; ++k) { r.pos[k] = res.pos[k]; r.normal[k] = vxo::FACE_NORMALS[res.face_id][k]; }
            } else {                                                    // picker.glsl:45-50
                r.dst = -1; r.inside_voxel = 0;
            }
        }
#pragma omp critical
        {
            total.primary_rays += c.primary_rays; total.shadow_rays += c.shadow_rays; total.steps += c.steps;
            total.pushes += c.pushes; total.leaf_tests += c.leaf_tests; total.tex_fetches += c.tex_fetches;
        }
    }
    if (counters) *counters = total;
}

// world.glsl main() over rows [y0,y1) of a w x h image; out is the FULL image (w*h*4 floats, row 0 = bottom).
void vxo_render(const uint8_t* world, uint64_t world_len, const void* materials, uint32_t n_materials, const VxoTexture* tex, int svo_format,
                const vxo::RenderParams* params, uint32_t w, uint32_t h, uint32_t y0, uint32_t y1, float* out,
                vxo::Counters* counters, int threads) {
    vxo::Scene s = make_scene(world, world_len, materials, n_materials, tex, svo_format);
    const float tan_half_fov = tanf(params->fov_y_rad * 0.5f);
    vxo::Counters total{};
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        vxo::Counters c{};
#pragma omp for schedule(dynamic, 4)
        for (int64_t y = y0; y < (int64_t)y1; ++y)
            for (uint32_t x = 0; x < w; ++x)
                vxo::render_pixel(s, *params, tan_half_fov, x, (uint32_t)y, w, h, out + ((size_t)y * w + x) * 4, &c);
#pragma omp critical
        {
            total.primary_rays += c.primary_rays; total.shadow_rays += c.shadow_rays; total.steps += c.steps;
            total.pushes += c.pushes; total.leaf_tests += c.leaf_tests; total.tex_fetches += c.tex_fetches;
        }
    }
    if (counters) *counters = total;
}

// Per-pixel primary-hit record for parity diffing of hit voxel / material / face / distance.
struct VxoHit { float t; uint32_t value; int32_t face_id; float pos[3]; uint32_t near_boundary; };
void vxo_primary_hits(const uint8_t* world, uint64_t world_len, const void* materials, uint32_t n_materials, const VxoTexture* tex, int svo_format,
                      const vxo::RenderParams* params, uint32_t w, uint32_t h, VxoHit* out, int threads) {
    vxo::Scene s = make_scene(world, world_len, materials, n_materials, tex, svo_format);
    const float tan_half_fov = tanf(params->fov_y_rad * 0.5f);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < (int64_t)h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const float* m = params->view;
            float ux = (float)x / (float)w, uy = (float)y / (float)h;
            ux = ux * 2.0f - 1.0f; uy = uy * 2.0f - 1.0f; ux *= params->aspect_ratio; ux *= tan_half_fov; uy *= tan_half_fov;
            float rw = m[15];
            vxo::vec3 ro = vxo::v3(m[12] / rw, m[13] / rw, m[14] / rw);
            float lx = ((m[0] * ux + m[4] * uy) + m[8] * -1.0f) + m[12];
            float ly = ((m[1] * ux + m[5] * uy) + m[9] * -1.0f) + m[13];
            float lz = ((m[2] * ux + m[6] * uy) + m[10] * -1.0f) + m[14];
            float lw = ((m[3] * ux + m[7] * uy) + m[11] * -1.0f) + m[15];
            vxo::vec3 rd = vxo::normalize(vxo::v3(lx / lw - ro.x, ly / lw - ro.y, lz / lw - ro.z));
            vxo::OctreeResult res;
            vxo::intersect_octree(s, ro, rd, -1.0f, true, res, nullptr, nullptr);
            VxoHit& o = out[(size_t)y * w + x];
            o.t = res.t; o.value = res.value; o.face_id = res.face_id;
            for (int k = 0; k < 3; ++k) o.pos[k] = res.pos[k];
            // within-epsilon-of-a-voxel-boundary flag (north_star): uv within 1e-4 of a face edge
            o.near_boundary = (res.t >= 0 && (res.uv[0] < 1e-4f || res.uv[0] > 1.0f - 1e-4f || res.uv[1] < 1e-4f || res.uv[1] > 1.0f - 1e-4f)) ? 1u : 0u;
        }
}

// Analysis aid (tests/analysis/warp_efficiency.py): loop iterations (svo.esvo.glsl:152) of every pixel's primary ray and of its
// shadow ray (0 = none cast), so that warp-level scheduling policies can be compared offline.
void vxo_render_steps(const uint8_t* world, uint64_t world_len, const void* materials, uint32_t n_materials, const VxoTexture* tex, int svo_format,
                      const vxo::RenderParams* params, uint32_t w, uint32_t h, uint32_t* primary_steps, uint32_t* shadow_steps, int threads) {
    vxo::Scene s = make_scene(world, world_len, materials, n_materials, tex, svo_format);
    const float tan_half_fov = tanf(params->fov_y_rad * 0.5f);
    vxo::RenderParams no_shadows = *params;
    no_shadows.render_shadows = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < (int64_t)h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            float px[4];
            vxo::Counters a{}, b{};
            vxo::render_pixel(s, no_shadows, tan_half_fov, x, (uint32_t)y, w, h, px, &a);
            vxo::render_pixel(s, *params, tan_half_fov, x, (uint32_t)y, w, h, px, &b);
            primary_steps[(size_t)y * w + x] = (uint32_t)a.steps;
            shadow_steps[(size_t)y * w + x] = (uint32_t)(b.steps - a.steps);
        }
}

// glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE) of an RGBA32F attachment (framebuffer.rs:97-105)
void vxo_to_rgba8(const float* rgba32f, uint64_t n_pixels, uint8_t* out) {
    for (uint64_t i = 0; i < n_pixels * 4; ++i) {
        float c = rgba32f[i];
        if (!(c == c)) c = 0.0f;
        c = vxo::gl_clamp(c, 0.0f, 1.0f);
        out[i] = (uint8_t)(int)(c * 255.0f + 0.5f);
    }
}

int vxo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
