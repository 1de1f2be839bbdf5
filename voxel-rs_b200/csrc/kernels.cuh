// kernels.cuh — the __global__ entry points of libvoxelrt (sm_100a).
//
// A frame (world.glsl main()) is a WAVEFRONT of three kernels on one stream:
//   trace_primary_kernel   primary rays: ray generation + SVO traversal only -> one 32-byte hit record per pixel
//   shade_kernel           one thread per pixel, straight-line: texture colour, highlight, normal map, lighting, sky;
//                          writes final pixels, and for lit pixels appends a shadow-ray record to a compacted list
//   trace_shadow_kernel    shadow rays of that list: the same traversal loop -> final pixel
// and the picker (picker.glsl main()) is the same traversal loop over a task array (trace_picker_kernel).
// Why not one fused kernel: the fused version was measured instruction-cache bound (profiles/r01_v0_*: the largest
// stall reason was no_instruction; 55-70 KB of SASS against a 32 KB L1.5 / ~6 KB L0 I-cache, with warps of one SM
// at unrelated places of it). Split this way the traversal kernels are a few KB of SASS that stay in the L0 cache,
// and the big straight-line shading code is walked by all warps of an SM together.
//
// The traversal kernels are persistent: grid = SMs x resident CTAs, every warp loops
//   refill   idle lanes take the next rays of the warp's current run (the 32 pixels of an 8x4 warp tile / 32 shadow-list
//            entries / 128 picker tasks); when the run is used up lane 0 claims the next one with one atomicAdd and
//            broadcasts it by shuffle;
//   walk     all lanes step their ray in lock-step; a warp vote after every step leaves the loop once fewer than
//            `refill_threshold` lanes are still walking;
//   events   finished lanes write their result and become idle.
// All of them are compiled per node format (template <int FMT>: VX_FMT_ESVO / VX_FMT_CSVO).
#pragma once
#include "../../include/voxelrt.h"
#include "traverse.cuh"

namespace vx {

struct RenderArgs {
    Scene scene;
    RenderUniforms u;
    float4* frame;                // RGBA32F, row 0 = bottom (world.glsl:140); may be a peer GPU's memory (write-only here)
    uint32_t* frame8;             // non-null: finished pixels are stored as RGBA8 here INSTEAD (glReadPixels rounding, framebuffer.rs:97-105);
                                  // may be a peer GPU's memory — a quarter of the NVLink bytes of the RGBA32F gather
    float4* hit0;                 // per pixel slot: {dst, value, u, v}
    float4* hit1;                 // per pixel slot: {pos.x, pos.y, pos.z, flags}  flags: bits 0-2 face, bit 3 hit
    float4* sh0;                  // shadow list: {origin.xyz, diffuse+specular}
    float4* sh1;                  // shadow list: res.color
    uint32_t* sh_pix;             // shadow list: pixel index (y * width + x)
    unsigned int* shadow_count;   // entries in the shadow list
    Counters* counters;
    unsigned int* work_counter;   // persistent kernels: next unclaimed run
    uint32_t macro_x, macro_y;    // frame size in 32x16-pixel macro blocks
    uint32_t macro_x_magic;       // floor(2^32 / macro_x) + 1: n / macro_x == __umulhi(n, magic) for n * macro_x < 2^32 (set_macro_grid)
    uint32_t macro0, n_macros;    // this launch covers macro blocks [macro0, macro0 + n_macros) (a band of macro rows, or the frame)
    uint32_t first_owned;         // first macro block >= macro0 owned by this shard
    uint32_t n_owned;             // macro blocks of [macro0, macro0 + n_macros) owned by this shard
    uint32_t shard_rank, shard_size;
    uint32_t shard_rows;          // 0: macro block m belongs to shard m % size (interleaved blocks: the finest mix of cheap and expensive
                                  // image regions, for frames that stay on a GPU); 1: macro ROW r belongs to shard r % size (every shard's
                                  // pixels are whole 16-pixel-high stripes of the frame: one strided DMA moves them to a host frame)
    uint32_t refill_threshold;    // leave the walk loop when fewer lanes than this are still walking
    uint32_t shadow_refill;       // the same for trace_shadow_kernel
    uint32_t fetch_tiles;         // warp tiles claimed per work-counter atomicAdd (1..4: fewer same-address atomics on big frames)
    const uint32_t* work_list;    // A/B (vx_set_option 13, whole unsharded frames only): non-null = the k-th work unit is macro block
                                  // work_list[k] (a Z-order curve over the frame's macro blocks) instead of the k-th in row-major order
    unsigned int* strip_done;     // non-null: OVERLAPPED wavefront. strip_done[strip] counts the pixels of that 32x4 strip whose hit record is
                                  // written; shade_kernel runs concurrently with trace_primary_kernel (own stream) and a CTA waits for its
                                  // strip to be complete instead of for the whole kernel: it fills the SMs the tracing kernel's tail frees
    unsigned int* sync_errors;    // counts waits of the overlapped wavefront that gave up
    unsigned int* shade_counter;  // overlapped wavefront: shade_kernel is PERSISTENT too — a grid small enough to be resident as a whole (next
                                  // to the tracing kernel's CTAs), its CTAs claim strips in order from this counter
    uint32_t shade_blocks;        // strips to shade in this launch (owned macro blocks x 4)
    uint32_t tma_writeback;       // shade_kernel: stage the strip's pixels in shared memory and write them back with bulk async
                                  // copies (TMA engine, cp.async.bulk -> SASS UBLKCP), one 512-byte row per copy
    uint32_t lifo;                // bit 0: LIFO hand-over of the wavefront buffers (vx_set_option 14): hit records and shadow-list entries are
                                  // stored with the default L2 policy, the consuming kernel walks them LAST WRITTEN FIRST (what the
                                  // producer wrote last is what the 126 MB L2 still holds) and discards every line it has read
                                  // (discard.global.L2: a dirty line that is dead needs no write-back; bit 1: shade_kernel discards the
                                  // hit records, bit 2: trace_shadow_kernel discards the shadow list) — see wave_discard()
};

// Work units. The frame is cut into macro blocks of 32x16 pixels (row-major over the frame; a shard owns every
// shard_size-th block). A macro block is 4 strips of 32x4 pixels, a strip is 4 warp tiles of 8x4 pixels laid side by
// side, and pixel p of a strip is lane p%32 of tile p/32: consecutive work indices, consecutive pixels of a warp and
// consecutive tiles of a strip are all spatial neighbours (coherent rays, shared nodes in L1).
// Pixel slot (index into hit0/hit1) = strip * 128 + p.
// Frame size in macro blocks + the multiplier that replaces the kernels' divisions by macro_x (every shade CTA and every work fetch of
// the tracing kernels turns a macro-block index into a position: an integer division is ~20 instructions, the multiply is one).
inline void set_macro_grid(RenderArgs& a, uint32_t width, uint32_t height) {
    a.macro_x = (width + 31) / 32; a.macro_y = (height + 15) / 16;
    a.macro_x_magic = a.macro_x > 1 ? (uint32_t)(0x100000000ull / a.macro_x) + 1u : 0u;   // exact while macro blocks x macro_x < 2^32 (a 64K x 64K frame)
}
__host__ __device__ __forceinline__ uint32_t div_macro_x(const RenderArgs& a, uint32_t n) {
#if defined(__CUDA_ARCH__) || defined(VX_HOST_EMULATION)
    return a.macro_x_magic ? __umulhi(n, a.macro_x_magic) : n;
#else
    return n / a.macro_x;
#endif
}
__host__ __device__ __forceinline__ bool macro_owned(const RenderArgs& a, uint32_t macro) {
    if (a.shard_size <= 1) return true;
    return (a.shard_rows ? div_macro_x(a, macro) % a.shard_size : macro % a.shard_size) == a.shard_rank;
}
// k-th macro block this shard owns inside the band of the launch (first_owned: its first owned macro block, resp. macro ROW)
__device__ __forceinline__ uint32_t owned_macro(const RenderArgs& a, uint32_t k) {
    if (!a.shard_rows) return a.first_owned + k * a.shard_size;
    const uint32_t j = div_macro_x(a, k);
    return (a.first_owned + j * a.shard_size) * a.macro_x + (k - j * a.macro_x);
}
// Band [row0, row1) of macro rows: what of it this shard owns. Host side of the launch (voxelrt.cu, and tests/emu's stand-in for it).
inline void shard_band(RenderArgs& a, uint32_t row0, uint32_t row1) {
    a.macro0 = row0 * a.macro_x;
    a.n_macros = (row1 - row0) * a.macro_x;
    const uint32_t size = a.shard_size, rank = a.shard_rank;
    if (a.shard_rows) {
        a.first_owned = row0 + ((rank + size - (row0 % size)) % size);
        a.n_owned = a.first_owned < row1 ? ((row1 - a.first_owned + size - 1) / size) * a.macro_x : 0;
    } else {
        a.first_owned = a.macro0 + ((rank + size - (a.macro0 % size)) % size);
        const uint32_t band_end = a.macro0 + a.n_macros;
        a.n_owned = a.first_owned < band_end ? (band_end - a.first_owned + size - 1) / size : 0;
    }
}
__device__ __forceinline__ bool strip_origin(const RenderArgs& a, uint32_t strip, uint32_t& x0, uint32_t& y0) {
    const uint32_t macro = strip >> 2;
    if (!macro_owned(a, macro)) return false;
    const uint32_t row = div_macro_x(a, macro);
    x0 = (macro - row * a.macro_x) * 32;
    y0 = row * 16 + (strip & 3u) * 4;
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

// texels the shader's textureLod reads for this lod (the counters follow the shader, which samples at every leaf test)
__device__ __forceinline__ uint32_t logical_texels(const TexInfo* ti, float lod) {
    if (!(lod > 0.0f)) return 1;
    const uint32_t levels = __ldg(&ti->levels);
    const float l = gl_min(lod, (float)(levels - 1));
    const float fl = floorf(l);
    const uint32_t d1 = (uint32_t)fl, d2 = (d1 + 1 < levels) ? d1 + 1 : levels - 1;
    return (d2 == d1 || l - fl == 0.0f) ? 4 : 8;
}
__device__ __forceinline__ float lod_of_dst(float dst) {   // svo.esvo.glsl:235
    float sm = gl_clamp((dst - 15.0f) / (25.0f - 15.0f), 0.0f, 1.0f);
    sm = (sm * sm) * (3.0f - 2.0f * sm);
    return (sm * (dst - 15.0f)) * 0.05f;
}

// Translucency rule of a render ray at a leaf whose material is not known to be opaque (svo.esvo.glsl:227-242, 264-265):
// samples the texel and applies "alpha > 0 and first of its kind". Out of line and rare (glass, leaves, water).
__device__ __noinline__ bool translucent_leaf_accepts(const TexInfo* tex, const Material* materials, uint32_t n_materials, const float* unorm,
                                                      uint32_t value, int face_id, float u, float v, float dst, bool first_of_kind) {
    const Material* m = materials + (value < n_materials ? value : n_materials - 1);
    int tex_id = __ldg(&m->tex_side);
    if (face_id == 3) tex_id = __ldg(&m->tex_top);
    else if (face_id == 2) tex_id = __ldg(&m->tex_bottom);
    uint32_t nf = 0;
    const float4 c = texture_lod<4>(tex, unorm, u, v, tex_id, lod_of_dst(dst), &nf);
    return c.w > 0.0f && first_of_kind;
}

// Render-ray leaf candidate (cast_translucent = true). Returns true when the leaf is the hit; fills g (and value).
template <int FMT, bool COUNT>
__device__ __forceinline__ bool render_leaf(Walk& w, const Scene& s, const Smem& sm, float inv_scale, uint32_t& last_leaf, Leaf& g, Counters& cnt) {
    g.value = leaf_value<FMT>(w, s);
    if (COUNT) { cnt.leaf_tests++; cnt.tex_fetches += logical_texels(s.tex, lod_of_dst(w.t_min * inv_scale)); }
    const bool opaque = g.value < 64u && ((s.opaque_materials >> g.value) & 1ull);
    const float* c = sm.cold;
    leaf_geom(w, c[0], c[VX_THREADS], c[2 * VX_THREADS], c[3 * VX_THREADS], c[4 * VX_THREADS], c[5 * VX_THREADS], inv_scale, g);
    if (opaque) return true;
    const bool first_of_kind = !(w.flags & VX_FLAG_ADJACENT) || g.value != last_leaf;   // :241
    if (translucent_leaf_accepts(s.tex, s.materials, s.n_materials, sm.unorm, g.value, g.face_id, g.u, g.v, g.dst, first_of_kind)) return true;
    last_leaf = g.value; w.flags |= VX_FLAG_ADJACENT;                         // :264-265
    return false;
}

// The walk loop of a warp: every lane with a live ray steps it; the warp leaves the loop when fewer than `thresh` lanes are
// still walking. thresh == 1 (run every ray of the warp to its end, then refill all 32 lanes at once) needs no population
// count and gets its own copy of the loop.
// TWO = two steps per warp vote in the threshold-1 loop. Measured (profiles/r02_step_variants.md, 4K frame): trace_primary_kernel
// 0.771 -> 0.752 ms, trace_shadow_kernel 0.498 -> 0.544 ms — the lanes of a coherent primary tile stay in the same phase and save the vote,
// its WARPSYNC and the loop branch every other iteration; shadow rays start in 32 different voxels and the second, nested step runs
// once per phase group of the first. So the primary kernel (ESVO) instantiates it and the others do not.
template <int FMT, bool LIMITED, bool COUNT, int TWO = 0>   // TWO: 0 one step per vote, 1 two (the second nested in the first), 2 two (one after the other)
__device__ __forceinline__ void walk_warp(Walk& w, const Scene& s, uint32_t stk, Counters& cnt, int thresh) {
    if (thresh <= 1) {
        do {
            if (w.state > 0) {
                walk_step<FMT, LIMITED, COUNT>(w, s, stk, cnt);
                if (TWO == 1 && w.state > 0) walk_step<FMT, LIMITED, COUNT>(w, s, stk, cnt);
            }
            if (TWO == 2 && w.state > 0) walk_step<FMT, LIMITED, COUNT>(w, s, stk, cnt);
        } while (__any_sync(0xffffffffu, w.state > 0));
    } else {
        do {
            if (w.state > 0) walk_step<FMT, LIMITED, COUNT>(w, s, stk, cnt);
#ifndef VX_V_THRESH_ONE   // two steps per population count in the refilling loop (the picker's): 2.558 -> 2.499 ms per 16 Mi rays; a lane may
                          // take one step past the moment the warp would have left the loop — the same steps of the same ray, only earlier
            if (w.state > 0) walk_step<FMT, LIMITED, COUNT>(w, s, stk, cnt);
#endif
        } while (__popc(__ballot_sync(0xffffffffu, w.state > 0)) >= thresh);
    }
}

// octree_scale = the f32 at byte 0 of the world buffer: one word (ESVO) / two words (CSVO: scale, root_ptr) before descriptors[]
template <int FMT>
__device__ __forceinline__ float load_octree_scale(const Scene& s) { return __uint_as_float(__ldg(s.desc - (FMT == VX_FMT_CSVO ? 2 : 1))); }

// round(clamp(c, 0, 1) * 255) per channel: what glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE) returns for the RGBA32F attachment
__device__ __forceinline__ uint32_t pack_rgba8(float4 c) {
    const float v[4] = {c.x, c.y, c.z, c.w};
    uint32_t p = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float f = v[k];
        if (!(f == f)) f = 0.0f;
        f = gl_clamp(f, 0.0f, 1.0f);
        p |= (uint32_t)(int)(f * 255.0f + 0.5f) << (8 * k);
    }
    return p;
}
// A finished pixel leaves the kernel: 16-byte streaming store into the RGBA32F frame, or 4 bytes into the RGBA8 frame.
__device__ __forceinline__ void store_pixel(const RenderArgs& a, uint32_t pix, float4 c) {
    if (a.frame8) __stcs(a.frame8 + pix, pack_rgba8(c));
    else __stcs(a.frame + pix, c);
}

// ---- LIFO hand-over of the wavefront buffers ---------------------------------------------------------------------------------------
// The hit records (32 B per pixel) and the shadow list (36 B per entry) are written once and read once; as streaming (evict-first)
// stores behind a frame-sized kernel they all went to HBM and came back (profiles/r02_v5_frame_wavefront.md: 0.78 GB of DRAM traffic
// per 4K frame against 0.17 GB compulsory). With RenderArgs::lifo the consumer starts at the producer's END, where the lines are
// still dirty in L2, and tells the L2 that a line it has consumed is dead: no fill from HBM, no write-back to HBM for that part.
// MEASURED AND OFF BY DEFAULT (profiles/r02_lifo.md): 265 MB of records with normal priority evict the 33 MB SVO from the 126 MB L2;
// trace_shadow_kernel loses more (0.469 -> 0.498 ms) than shade_kernel could ever gain on a path that is issue-bound, not DRAM-bound.
// A warp's 32 records are 512 contiguous, 512-byte-aligned bytes of each array = four 128-byte lines: lanes 0-3 drop them.
__device__ __forceinline__ void wave_discard(const void* line) {
#ifndef VX_HOST_EMULATION
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(line) : "memory");
#else
    std::memset(const_cast<void*>(line), 0xCD, 128);   // the emulator poisons what the product declares dead: a later read would show
#endif
}
// A value that is 0, but only known once `loaded` has arrived in EVERY lane of the warp (the shuffle reads the warp's register):
// added to the address of a discard it keeps the discard behind the loads of the lines it drops.
__device__ __forceinline__ uint32_t after_warp_loads(uint32_t loaded) {
    return loaded - __shfl_sync(0xffffffffu, loaded, (int)(threadIdx.x & 31u));
}

// ---- overlapped wavefront: strip completion flags ------------------------------------------------------------------------------
// Producer side (trace_primary_kernel, after its lanes wrote their hit records): the lanes of the warp that finished a pixel in
// this round are counted per strip — they normally all belong to one strip, two when a refill straddled tiles. __syncwarp orders
// the other lanes' stores before the leader's release-reduction, which publishes them device-wide.
__device__ __forceinline__ void strips_signal(unsigned int* strip_done, bool finished, uint32_t strip) {
    unsigned fin = __ballot_sync(0xffffffffu, finished);
    if (!fin) return;
#ifndef VX_HOST_EMULATION
    __syncwarp();
#endif
    const uint32_t lane = threadIdx.x & 31;
    while (fin) {
        const int leader = __ffs(fin) - 1;
        const uint32_t s = __shfl_sync(0xffffffffu, strip, leader);
        const unsigned same = __ballot_sync(0xffffffffu, finished && strip == s);
        if ((int)lane == leader) {
#ifndef VX_HOST_EMULATION
            // release-reduction: orders this warp's record stores before the count without the full fence of __threadfence()
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(strip_done + s), "r"((unsigned)__popc(same)) : "memory");
#else
            strip_done[s] += (unsigned)__popc(same);
#endif
        }
        fin &= ~same;
    }
}
// Consumer side (shade_kernel, one thread of the CTA): wait until `expected` pixels of the strip are in. Gives up after ~1 s and
// counts it: a lost producer must show up as an error, never as a hung GPU.
__device__ __forceinline__ void strip_wait(const unsigned int* strip_done, uint32_t strip, uint32_t expected, unsigned int* sync_errors) {
    unsigned int v = 0, spins = 0;
    for (;;) {
#ifndef VX_HOST_EMULATION
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(strip_done + strip) : "memory");
#else
        v = strip_done[strip];
#endif
        if (v >= expected) break;
        __nanosleep(spins < 64 ? 100 : 1000);
        if (++spins > (1u << 20)) { atomicAdd(sync_errors, 1u); break; }
    }
}

// ---- primary rays ------------------------------------------------------------------------------------------------------
#ifndef VX_PRIMARY_CLIP
#define VX_PRIMARY_CLIP true   // world-box clipping of primary rays (A/B: tools/ab_kernels.py builds a variant with false)
#endif
#ifndef VX_PRIMARY_UNROLL2
#define VX_PRIMARY_UNROLL2 1      // steps per vote in the primary kernel's walk loop: walk_warp's TWO (A/B builds: 0, 2)
#endif
#ifndef VX_CSVO_UNROLL2
#define VX_CSVO_UNROLL2 true      // ... for the CSVO instantiation too (measured: 1.291 -> 1.255 ms although two copies of its 241-instruction loop
#endif                            // do not fit the L0 I-cache)
#ifndef VX_SHADOW_UNROLL2
#define VX_SHADOW_UNROLL2 0       // the shadow kernel's (measured: 1 is 9 % slower there)
#endif
template <int FMT, bool COUNT, int MINB>
__global__ void __launch_bounds__(VX_THREADS, MINB) trace_primary_kernel(RenderArgs a) {
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene.unorm, a.scene.stack_levels, smem_raw);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    const float octree_scale = load_octree_scale<FMT>(a.scene);
    const float inv_scale = 1.0f / octree_scale;
    const Clip clip = load_clip(a.scene, octree_scale);
    float* cold = sm.cold;   // 0-2 origin, 3-5 direction (both in [1,2) space / epsilon-clamped)
    Counters cnt = {0, 0, 0, 0, 0, 0};

    // warp-uniform work state: the unit of work fetch is one warp tile (8x4 pixels = 32 rays, a quarter of a strip), so that a
    // shard of a frame (1/8 of 4K = 32 k tiles over 4.7 k resident warps) still load-balances
    // Only the macro blocks this shard owns are enumerated (work index w -> owned macro block w / 16): a shard of 1/8 of
    // the frame performs 1/8 of the work-counter atomics (enumerating all tiles and skipping 7 of 8 made the single
    // counter the bottleneck of the sharded kernel, profiles/r01_scaling_before_owned_tiles.jsonl).
    uint32_t tile = 0, strip_x0 = 0, strip_y0 = 0, next_px = 32, tile_px0 = 0;
    uint32_t work = 0, work_end = 0;   // claimed run of work indices [work, work_end): a.fetch_tiles per atomicAdd
    const uint32_t n_tiles = a.n_owned * 16u;
    bool more_work = true;
    uint32_t slot = 0, last_leaf = 0xffffffffu;
    Walk w;
    w.state = ST_IDLE;

    for (;;) {
        // ---------------------------------------------------------------- refill
        unsigned want = __ballot_sync(0xffffffffu, w.state == ST_IDLE);
        while (want && more_work) {
            if (next_px >= 32) {
                if (work >= work_end) {
                    if (lane == 0) work = atomicAdd(a.work_counter, a.fetch_tiles);
                    work = __shfl_sync(0xffffffffu, work, 0);
                    work_end = min(work + a.fetch_tiles, n_tiles);
                    if (work >= n_tiles) { more_work = false; break; }
                }
                tile = work++;
                tile = (a.work_list ? __ldg(a.work_list + (tile >> 4)) : owned_macro(a, tile >> 4)) * 16u + (tile & 15u);
                if (!strip_origin(a, tile >> 2, strip_x0, strip_y0)) continue;
                tile_px0 = (tile & 3u) * 32u;
                next_px = 0;
            }
            const uint32_t n_take = min((uint32_t)__popc(want), 32u - next_px);
            const uint32_t my_rank = __popc(want & lanemask_lt);
            if (((want >> lane) & 1u) && my_rank < n_take) {
                uint32_t gx, gy;
                strip_pixel(strip_x0, strip_y0, tile_px0 + next_px + my_rank, gx, gy);
                if (gx < a.u.width && gy < a.u.height) {
                    float ox, oy, oz, dx, dy, dz, rox, roy, roz, rdx, rdy, rdz;
                    primary_ray(a.u, gx, gy, ox, oy, oz, dx, dy, dz);
                    walk_init<FMT>(w, a.scene, VX_PRIMARY_CLIP ? clip : Clip{0, 0, 0, 0, 0, 0, 0}, octree_scale, ox, oy, oz, dx, dy, dz, -1.0f, rox, roy, roz, rdx, rdy, rdz);
                    cold[0] = rox; cold[VX_THREADS] = roy; cold[2 * VX_THREADS] = roz;
                    cold[3 * VX_THREADS] = rdx; cold[4 * VX_THREADS] = rdy; cold[5 * VX_THREADS] = rdz;
                    slot = (tile >> 2) * 128u + tile_px0 + next_px + my_rank;
                    last_leaf = 0xffffffffu;
                    if (COUNT) cnt.primary_rays++;
                }
            }
            next_px += n_take;
            want = __ballot_sync(0xffffffffu, w.state == ST_IDLE);
        }
        const unsigned busy = __ballot_sync(0xffffffffu, w.state != ST_IDLE);
        if (!busy) break;

        // ---------------------------------------------------------------- walk
        walk_warp<FMT, VX_PRIMARY_CLIP, COUNT, (FMT == VX_FMT_ESVO || VX_CSVO_UNROLL2) ? VX_PRIMARY_UNROLL2 : 0>(w, a.scene, sm.stack, cnt, min((int)a.refill_threshold, __popc(busy)));

        // ---------------------------------------------------------------- events
        bool finished = false;
        if (state_at_leaf(w.state)) {
            Leaf g;
            if (render_leaf<FMT, COUNT>(w, a.scene, sm, inv_scale, last_leaf, g, cnt)) {
                float px, py, pz;
                leaf_pos(w.t_min, cold[0], cold[VX_THREADS], cold[2 * VX_THREADS], cold[3 * VX_THREADS], cold[4 * VX_THREADS], cold[5 * VX_THREADS], g,
                         inv_scale, px, py, pz);
                const float4 r0 = make_float4(g.dst, __uint_as_float(g.value), g.u, g.v);
                const float4 r1 = make_float4(px, py, pz, __uint_as_float(8u | (uint32_t)g.face_id));
                if (a.lifo) { a.hit0[slot] = r0; a.hit1[slot] = r1; }      // stays in L2 for shade_kernel (LIFO hand-over)
                else { __stcs(a.hit0 + slot, r0); __stcs(a.hit1 + slot, r1); }
                w.state = ST_IDLE;
                finished = true;
            } else {
                walk_skip_leaf<FMT, VX_PRIMARY_CLIP>(w, a.scene, sm.stack);   // translucent / repeated leaf: finish this iteration at ADVANCE
            }
        } else if (state_missed(w.state)) {
            if (a.lifo) a.hit1[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            else __stcs(a.hit1 + slot, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            w.state = ST_IDLE;
            finished = true;
        }
        if (a.strip_done) strips_signal(a.strip_done, finished, slot >> 7);
    }
    if (COUNT) flush_counters(a.counters, cnt);
}

// ---- shading ---------------------------------------------------------------------------------------------------------------
#ifndef VX_SHADE_STRIPS
#define VX_SHADE_STRIPS 4u   // strips one CTA of the static shade grid shades, one after the other (A/B builds: 1, 2, 8)
#endif
// grid = owned strips, block = 128 threads = the 128 pixels of one 32x4 strip (4 warp tiles of 8x4).
template <bool COUNT, bool PERSIST>
__global__ void __launch_bounds__(VX_THREADS, PERSIST ? 5 : 10) shade_kernel(RenderArgs a) {   // static grid: 10 CTAs / SM (<= 51 registers)
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene.unorm, 0, smem_raw, false);
    __shared__ unsigned int s_warp_count[VX_THREADS / 32];
    __shared__ unsigned int s_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned int s_next;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    // static grid: CTA b shades the VX_SHADE_STRIPS strips b * VX_SHADE_STRIPS ... (= the 4 strips of one macro block: the CTA's prologue —
    // the unorm table, the launch arguments — is paid once per 512 pixels instead of once per 128; it was 15 % of the kernel's
    // instructions). Overlapped wavefront: the CTAs of a small resident grid claim strips from shade_counter.
    uint32_t static_i = 0;
    for (uint32_t blk = 0;;) {
    if (PERSIST) {
        __syncthreads();   // everybody is done with the shared arrays of the previous strip (and with s_next)
        if (threadIdx.x == 0) s_next = atomicAdd(a.shade_counter, 1u);
        __syncthreads();
        blk = s_next;
        if (blk >= a.shade_blocks) break;
    } else {
        if (static_i) __syncthreads();   // the shared arrays of the previous strip are free again
        blk = blockIdx.x * VX_SHADE_STRIPS + static_i;
        if (static_i++ >= VX_SHADE_STRIPS || blk >= a.shade_blocks) break;
        if (a.lifo) blk = a.shade_blocks - 1u - blk;   // LIFO: the strips traced last first
    }
    const uint32_t strip = ((a.work_list ? __ldg(a.work_list + (blk >> 2)) : owned_macro(a, blk >> 2)) << 2) | (blk & 3u);
    uint32_t x0 = 0, y0 = 0, gx = 0, gy = 0;
    const bool have = (strip >> 2) < a.macro0 + a.n_macros && strip_origin(a, strip, x0, y0);
    if (have) strip_pixel(x0, y0, threadIdx.x, gx, gy);
    const bool live = have && gx < a.u.width && gy < a.u.height;
    if (PERSIST) {        // overlapped wavefront: this strip's hit records may still be on their way
        if (threadIdx.x == 0 && have)
            strip_wait(a.strip_done, strip, min(32u, a.u.width - x0) * min(4u, a.u.height - y0), a.sync_errors);
        __syncthreads();
    }
    bool want_shadow = false;
    float4 s0 = make_float4(0, 0, 0, 0), s1 = make_float4(0, 0, 0, 0);
    const uint32_t pix = gy * a.u.width + gx;
    // TMA write-back of the strip (framebuffer tile = 4 rows x 32 pixels x 16 B): only whole strips of a local RGBA32F frame;
    // a pixel that still waits for its shadow ray gets res.color here and is overwritten by trace_shadow_kernel afterwards
    __shared__ __align__(128) float4 s_tile[VX_THREADS];
    const bool tile_store = a.tma_writeback && !a.frame8 && have && x0 + 32u <= a.u.width && y0 + 4u <= a.u.height;   // CTA-uniform
    float4* const tile_slot = s_tile + ((threadIdx.x >> 3) & 3u) * 32u + (threadIdx.x >> 5) * 8u + (threadIdx.x & 7u);
    const uint32_t slot = strip * 128u + threadIdx.x;
    float4 h0 = make_float4(0, 0, 0, 0), h1 = make_float4(0, 0, 0, 0);
    if (live) {
        if (a.lifo) { h1 = __ldcg(a.hit1 + slot); h0 = __ldcg(a.hit0 + slot); }
        else { h1 = __ldcs(a.hit1 + slot); h0 = __ldcs(a.hit0 + slot); }   // issued together (a miss leaves its hit0 slot unwritten: loaded, never used)
    }
    if ((a.lifo & 2u) && have) {   // CTA-uniform. This warp's records are read: their eight lines are dead (lanes 0-3 hit0, 4-7 hit1)
        const uint32_t zero = after_warp_loads(__float_as_uint(h0.x) ^ __float_as_uint(h1.w));
        if (lane < 8u) wave_discard(((lane & 4u) ? a.hit1 : a.hit0) + (slot - lane) + (lane & 3u) * 8u + zero);
    }
    if (live) {
        const uint32_t flags = __float_as_uint(h1.w);
        if (flags & 8u) {
            Leaf g;
            g.dst = h0.x; g.value = __float_as_uint(h0.y); g.u = h0.z; g.v = h0.w; g.face_id = (int)(flags & 7u);
            int tex_id; float tex_lod;
            leaf_texture(a.scene, g, tex_id, tex_lod);
            uint32_t nf = 0;
            const float4 c = texture_lod<4>(a.scene.tex, sm.unorm, g.u, g.v, tex_id, tex_lod, &nf);   // svo.esvo.glsl:237 (counted by the trace kernel)
            Shade sh;
            sh.r = c.x; sh.g = c.y; sh.b = c.z; sh.a = c.w;
            nf = 0;
            shade_hit(a.scene, sm.unorm, a.u, g, tex_lod, h1.x, h1.y, h1.z, sh, &nf);
            if (COUNT) cnt.tex_fetches += nf;
            float4 outc;
            if (sh.done) {
                outc = make_float4(sh.r, sh.g, sh.b, sh.a);
            } else if (sh.want_shadow) {
                want_shadow = true;
                s0 = make_float4(sh.sox, sh.soy, sh.soz, sh.lit);
                s1 = make_float4(sh.r, sh.g, sh.b, sh.a);
                outc = s1;
            } else {
                outc = shade_finish(a.u, sh.r, sh.g, sh.b, sh.a, sh.lit, 1.0f);
            }
            if (tile_store) *tile_slot = outc;
            else if (!want_shadow) store_pixel(a, pix, outc);
        } else {
            float ox, oy, oz, dx, dy, dz;
            primary_ray(a.u, gx, gy, ox, oy, oz, dx, dy, dz);
            const float4 outc = sky_color(dx, dy, dz);
            if (tile_store) *tile_slot = outc;
            else store_pixel(a, pix, outc);
        }
    }
#ifndef VX_HOST_EMULATION
    if (tile_store) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the async (TMA) proxy
#endif
    // compact the shadow rays of this strip into the global list: one atomicAdd per CTA, strip order kept inside it
    const unsigned m = __ballot_sync(0xffffffffu, want_shadow);
    if (lane == 0) s_warp_count[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int k = 0; k < VX_THREADS / 32; ++k) tot += s_warp_count[k];
        s_base = tot ? atomicAdd(a.shadow_count, tot) : 0u;
    }
    if (tile_store && threadIdx.x >= 32 && threadIdx.x < 36) {   // one bulk copy per row of the tile (lanes 0-3 of warp 1; warp 0 does the atomic)
        const uint32_t row = threadIdx.x - 32u;
        const float4* dst = a.frame + (size_t)(y0 + row) * a.u.width + x0;
#ifndef VX_HOST_EMULATION
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(s_tile + row * 32u);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(dst), "r"(src) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#else
        for (uint32_t k = 0; k < 32u; ++k) const_cast<float4*>(dst)[k] = s_tile[row * 32u + k];   // what the bulk copy moves
#endif
    }
    __syncthreads();
    if (want_shadow) {
        unsigned off = s_base + __popc(m & ((1u << lane) - 1u));
        for (uint32_t k = 0; k < warp; ++k) off += s_warp_count[k];
        a.sh0[off] = s0; a.sh1[off] = s1; a.sh_pix[off] = pix;
    }
    }
    if (COUNT) flush_counters(a.counters, cnt);
}

// ---- shadow rays -------------------------------------------------------------------------------------------------------------
template <int FMT, bool COUNT, int MINB>
__global__ void __launch_bounds__(VX_THREADS, MINB) trace_shadow_kernel(RenderArgs a) {
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene.unorm, a.scene.stack_levels, smem_raw);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    const float octree_scale = load_octree_scale<FMT>(a.scene);
    const float inv_scale = 1.0f / octree_scale;
    const Clip clip = load_clip(a.scene, octree_scale);
    const uint32_t n = *a.shadow_count;
    float* cold = sm.cold;
    Counters cnt = {0, 0, 0, 0, 0, 0};

    uint32_t run_base = 0, next = 32, run_len = 32;   // warp-uniform; runs of 32 list entries (one warp round) keep small shards balanced
    bool more_work = true;
    uint32_t entry = 0, last_leaf = 0xffffffffu;
    float lit = 0.0f;
    Walk w;
    w.state = ST_IDLE;
    // LIFO hand-over (RenderArgs::lifo): runs are claimed from the END of the list (what shade_kernel appended last is still in L2),
    // and (bit 2 of lifo) a run whose 32 rays are all finished is discarded: 4 lines of sh0, 4 of sh1, 1 of sh_pix (lanes 0-8). Only
    // with shadow_refill == 1, where a warp claims a new run when — and only when — every ray of the previous one is done.
    const bool lifo = a.lifo != 0u, lifo_discard = (a.lifo & 4u) && a.shadow_refill <= 1u;
    const uint32_t n_runs = (n + 31u) / 32u;
    uint32_t dead_run = 0xffffffffu;

    for (;;) {
        unsigned want = __ballot_sync(0xffffffffu, w.state == ST_IDLE);
        while (want && more_work) {
            if (next >= run_len) {
                if (lifo_discard && dead_run != 0xffffffffu && want == 0xffffffffu) {
                    if (lane < 4u) wave_discard(a.sh0 + dead_run + lane * 8u);
                    else if (lane < 8u) wave_discard(a.sh1 + dead_run + (lane - 4u) * 8u);
                    else if (lane == 8u) wave_discard(a.sh_pix + dead_run);
                    dead_run = 0xffffffffu;
                }
                if (lane == 0) run_base = atomicAdd(a.work_counter, 32u);
                run_base = __shfl_sync(0xffffffffu, run_base, 0);
                if (run_base >= n) { more_work = false; break; }
                if (lifo) run_base = (n_runs - 1u - run_base / 32u) * 32u;
                run_len = min(32u, n - run_base);
                next = 0;
                if (run_len == 32u) dead_run = run_base;   // (a ragged last run shares its lines with nothing, but is not worth a special case)
            }
            const uint32_t n_take = min((uint32_t)__popc(want), run_len - next);
            const uint32_t my_rank = __popc(want & lanemask_lt);
            if (((want >> lane) & 1u) && my_rank < n_take) {
                entry = run_base + next + my_rank;
                const float4 s0 = __ldcs(a.sh0 + entry);
                float rox, roy, roz, rdx, rdy, rdz;
                walk_init<FMT>(w, a.scene, clip, octree_scale, s0.x, s0.y, s0.z, -a.u.lx, -a.u.ly, -a.u.lz, -1.0f, rox, roy, roz, rdx, rdy, rdz);   // world.glsl:82
                cold[0] = rox; cold[VX_THREADS] = roy; cold[2 * VX_THREADS] = roz;
                cold[3 * VX_THREADS] = rdx; cold[4 * VX_THREADS] = rdy; cold[5 * VX_THREADS] = rdz;
                lit = s0.w;
                last_leaf = 0xffffffffu;
                if (COUNT) cnt.shadow_rays++;
            }
            next += n_take;
            want = __ballot_sync(0xffffffffu, w.state == ST_IDLE);
        }
        const unsigned busy = __ballot_sync(0xffffffffu, w.state != ST_IDLE);
        if (!busy) break;
        walk_warp<FMT, true, COUNT, VX_SHADOW_UNROLL2>(w, a.scene, sm.stack, cnt, min((int)a.shadow_refill, __popc(busy)));

        if (w.state <= 0 && w.state != ST_IDLE) {
            bool done = true;
            float shadow = 1.0f;                                     // world.glsl:83: res.t < 0 -> 1
            if (state_at_leaf(w.state)) {
                Leaf g;
                // fully opaque materials block the sun whatever the texel is; others go through the translucency rule
                const uint32_t value = leaf_value<FMT>(w, a.scene);
                if (value < 64u && ((a.scene.opaque_materials >> value) & 1ull)) {
                    if (COUNT) { cnt.leaf_tests++; cnt.tex_fetches += logical_texels(a.scene.tex, lod_of_dst(w.t_min * inv_scale)); }
                    shadow = 0.0f;
                } else if (render_leaf<FMT, COUNT>(w, a.scene, sm, inv_scale, last_leaf, g, cnt)) {
                    shadow = 0.0f;
                } else {
                    walk_skip_leaf<FMT, true>(w, a.scene, sm.stack);
                    done = false;
                }
            }
            if (done) {
                const float4 c = __ldcs(a.sh1 + entry);
                const uint32_t pix = __ldcs(a.sh_pix + entry);
                store_pixel(a, pix, shade_finish(a.u, c.x, c.y, c.z, c.w, lit, shadow));
                w.state = ST_IDLE;
            }
        }
    }
    if (COUNT) flush_counters(a.counters, cnt);
}

// ---- picker kernel: picker.glsl main() -------------------------------------------------------------------------------
// Same scheme: warps claim runs of 128 consecutive tasks, lanes refill from the run as their ray ends, so a warp is not
// held hostage by its longest ray (16 M random rays differ in length by 100x).
struct RaycastArgs {
    Scene scene;
    const float4* tasks;      // VxPickerTask = 3 x float4
    float4* results;          // VxPickerResult = 3 x float4
    unsigned long long n;
    Counters* counters;
    unsigned long long* work_counter;
    uint32_t refill_threshold;
    const uint32_t* order;    // non-null: the k-th ray traced is task order[k] (ray binning, see bin_count_kernel); results stay in task order
};

template <int FMT, bool COUNT>
__global__ void __launch_bounds__(VX_THREADS) trace_picker_kernel(RaycastArgs a) {
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene.unorm, a.scene.stack_levels, smem_raw, true, false);   // no texture is ever sampled here: no unorm table
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    const float octree_scale = load_octree_scale<FMT>(a.scene);
    const float inv_scale = 1.0f / octree_scale;
    const Clip clip = load_clip(a.scene, octree_scale);
    float* cold = sm.cold;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    unsigned long long run_base = 0;
    uint32_t next = 128, run_len = 128;
    bool more_work = true;
    unsigned long long my_task = 0;
    Walk w;
    w.state = ST_IDLE;

    for (;;) {
        unsigned want = __ballot_sync(0xffffffffu, w.state == ST_IDLE);
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
                if (a.order) my_task = __ldcs(a.order + my_task);
                const float4 t0 = __ldcs(a.tasks + 3 * my_task), t1 = __ldcs(a.tasks + 3 * my_task + 1), t2 = __ldcs(a.tasks + 3 * my_task + 2);
                float rox, roy, roz, rdx, rdy, rdz;
                walk_init<FMT>(w, a.scene, clip, octree_scale, t1.x, t1.y, t1.z, t2.x, t2.y, t2.z, t0.x, rox, roy, roz, rdx, rdy, rdz);
                cold[0] = rox; cold[VX_THREADS] = roy; cold[2 * VX_THREADS] = roz;
                cold[3 * VX_THREADS] = rdx; cold[4 * VX_THREADS] = rdy; cold[5 * VX_THREADS] = rdz;
                if (COUNT) cnt.primary_rays++;
            }
            next += n_take;
            want = __ballot_sync(0xffffffffu, w.state == ST_IDLE);
        }
        const unsigned busy = __ballot_sync(0xffffffffu, w.state != ST_IDLE);
        if (!busy) break;
        walk_warp<FMT, VX_PRIMARY_CLIP, COUNT>(w, a.scene, sm.stack, cnt, min((int)a.refill_threshold, __popc(busy)));
        if (w.state <= 0 && w.state != ST_IDLE) {
            // cast_translucent = false: the first leaf is the hit whatever its texel is (svo.esvo.glsl:241-242); the picker
            // never reads value or colour (picker.glsl:40-44), so neither the leaf word nor the texture is fetched.
            float4 o0 = make_float4(-1.0f, 0.0f, 0.0f, 0.0f), o1 = make_float4(0, 0, 0, 0), o2 = make_float4(0, 0, 0, 0);
            if (state_at_leaf(w.state)) {
                if (COUNT) cnt.leaf_tests++;
                Leaf g;
                leaf_geom(w, cold[0], cold[VX_THREADS], cold[2 * VX_THREADS], cold[3 * VX_THREADS], cold[4 * VX_THREADS], cold[5 * VX_THREADS], inv_scale, g);
                if (g.dst > 0.0f) {                                        // picker.glsl:40 (res.t > 0)
                    float px, py, pz;
                    leaf_pos(w.t_min, cold[0], cold[VX_THREADS], cold[2 * VX_THREADS], cold[3 * VX_THREADS], cold[4 * VX_THREADS], cold[5 * VX_THREADS],
                             g, inv_scale, px, py, pz);
                    o0.x = g.dst; o0.y = __uint_as_float((w.flags >> 8) & 1u);
                    o1 = make_float4(px, py, pz, 0.0f);
                    const int axis = g.face_id >> 1;
                    const float sgn = (g.face_id & 1) ? 1.0f : -1.0f;      // FACE_NORMALS, svo.glsl:2-9
                    o2 = make_float4(axis == 0 ? sgn : 0.0f, axis == 1 ? sgn : 0.0f, axis == 2 ? sgn : 0.0f, 0.0f);
                }
            }
            __stcs(a.results + 3 * my_task, o0); __stcs(a.results + 3 * my_task + 1, o1); __stcs(a.results + 3 * my_task + 2, o2);
            w.state = ST_IDLE;
        }
    }
    if (COUNT) flush_counters(a.counters, cnt);
}

// ---- ray binning for incoherent picker batches ("optional Morton/direction-octant binning for the 16 M incoherent config") ----
// A batch of random rays starts anywhere in the world; the 32 rays of a warp then descend through 32 unrelated chains of nodes
// (every node a cache miss) and, refilled at different times, are never in the same phase of the walk. The pre-pass orders the
// rays along a Z-order curve of their ORIGIN cell (optionally direction octant below it): rays traced together start in the same
// corner of the octree, so the root-to-origin descent — more than half of a picker ray's iterations — reads the same nodes and
// runs in lock-step. It is a counting sort in one atomic pass:
//   bin_count_kernel    key = Morton code of the origin cell [| octant]; rank inside the bin = atomicAdd(hist[key], 1)
//   bin_scan_*          exclusive prefix sum over the histogram, in place (reduce / scan of the block sums / apply)
//   bin_scatter_kernel  order[offset[key] + rank] = task index
// The order inside a bin depends on the atomics' timing; nothing observable does: the picker kernel writes result i for task i.
__device__ __forceinline__ uint32_t spread3(uint32_t v) {   // bits 0-9 of v -> bits 0, 3, 6, ... (Morton interleave of one axis)
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t bin_key_bits(uint32_t bits_axis, uint32_t with_octant) { return 3u * bits_axis + (with_octant ? 3u : 0u); }
__global__ void __launch_bounds__(256) bin_count_kernel(const float4* tasks, unsigned long long n, const uint32_t* scale_word, uint32_t bits_axis,
                                                        uint32_t with_octant, uint32_t* hist, uint2* keyrank) {
    const float cells = (float)(1u << bits_axis);
    const float to_cell = __uint_as_float(__ldg(scale_word)) * cells;          // octree_scale maps SVO voxel space to [0, 1): x cells
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 pos = __ldg(tasks + 3 * i + 1);
        // fmaxf / fminf return the non-NaN operand: a NaN origin lands in cell 0, an origin outside the octree in the nearest border cell
        const uint32_t cx = (uint32_t)fminf(fmaxf(pos.x * to_cell, 0.0f), cells - 1.0f);
        const uint32_t cy = (uint32_t)fminf(fmaxf(pos.y * to_cell, 0.0f), cells - 1.0f);
        const uint32_t cz = (uint32_t)fminf(fmaxf(pos.z * to_cell, 0.0f), cells - 1.0f);
        uint32_t key = spread3(cx) | (spread3(cy) << 1) | (spread3(cz) << 2);
        if (with_octant) {
            const float4 dir = __ldg(tasks + 3 * i + 2);
            key = (key << 3) | (dir.x > 0.0f ? 1u : 0u) | (dir.y > 0.0f ? 2u : 0u) | (dir.z > 0.0f ? 4u : 0u);   // octant_mask, svo.esvo.glsl:112-126
        }
        const uint32_t rank = atomicAdd(hist + key, 1u);
        keyrank[i] = make_uint2(key, rank);
    }
}
__global__ void __launch_bounds__(256) bin_scatter_kernel(const uint2* keyrank, unsigned long long n, const uint32_t* offsets, uint32_t* order) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint2 kr = keyrank[i];
        order[__ldg(offsets + kr.x) + kr.y] = (uint32_t)i;
    }
}
// Exclusive prefix sum of `data[0, n)` in place, n a multiple of VX_SCAN_TILE: block b owns elements [b * TILE, (b + 1) * TILE), thread t of
// it the 16 consecutive ones from t * 16.
#define VX_SCAN_TILE 4096u
__device__ __forceinline__ uint32_t scan_block_exclusive(uint32_t mine, uint32_t* total) {   // exclusive scan of one value per thread over 256 threads
    __shared__ uint32_t s_scan[2][256];
    uint32_t cur = 0;
    s_scan[0][threadIdx.x] = mine;
    __syncthreads();
    for (uint32_t d = 1; d < 256u; d <<= 1) {
        const uint32_t v = s_scan[cur][threadIdx.x] + (threadIdx.x >= d ? s_scan[cur][threadIdx.x - d] : 0u);
        s_scan[cur ^ 1u][threadIdx.x] = v;
        cur ^= 1u;
        __syncthreads();
    }
    const uint32_t incl = s_scan[cur][threadIdx.x];
    if (total) *total = s_scan[cur][255];
    __syncthreads();   // the arrays may be reused by the caller's next call
    return incl - mine;
}
__global__ void __launch_bounds__(256) bin_scan_reduce_kernel(const uint32_t* data, uint32_t* block_sums) {
    const uint4* p = reinterpret_cast<const uint4*>(data + (size_t)blockIdx.x * VX_SCAN_TILE + threadIdx.x * 16u);
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const uint4 v = p[k]; sum += v.x + v.y + v.z + v.w; }
    uint32_t total;
    scan_block_exclusive(sum, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) bin_scan_sums_kernel(uint32_t* block_sums, uint32_t n_blocks) {   // one CTA: exclusive scan of the block sums
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_blocks; base += 256u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = scan_block_exclusive(v, &total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(256) bin_scan_apply_kernel(uint32_t* data, const uint32_t* block_sums) {
    uint4* p = reinterpret_cast<uint4*>(data + (size_t)blockIdx.x * VX_SCAN_TILE + threadIdx.x * 16u);
    uint4 v[4];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = p[k]; sum += v[k].x + v[k].y + v[k].z + v[k].w; }
    uint32_t run = block_sums[blockIdx.x] + scan_block_exclusive(sum, nullptr);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint4 o;
        o.x = run; run += v[k].x; o.y = run; run += v[k].y; o.z = run; run += v[k].z; o.w = run; run += v[k].w;
        p[k] = o;
    }
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

// The step function does not carry the shader's (ptr, parent_octant_idx); the debug kernel shadows them (plus their
// stacks) next to it to emit reference-format frames. Launched <<<1, VX_THREADS>>>; thread 0 casts.
template <int FMT>
__global__ void __launch_bounds__(VX_THREADS) debug_cast_kernel(DebugArgs a) {
    extern __shared__ uint32_t smem_raw[];
    const Smem sm = make_smem(a.scene.unorm, a.scene.stack_levels, smem_raw);
    if (threadIdx.x != 0) return;
    const Scene& s = a.scene;
    const float octree_scale = load_octree_scale<FMT>(s);
    const float inv_scale = 1.0f / octree_scale;
    Walk w;
    Counters cnt = {0, 0, 0, 0, 0, 0};
    float rox, roy, roz, rdx, rdy, rdz;
    Clip no_clip; no_clip.mode = 0;   // the debug cast reproduces every iteration of the shader
    walk_init<FMT>(w, s, no_clip, octree_scale, a.pos[0], a.pos[1], a.pos[2], a.dir[0], a.dir[1], a.dir[2], a.max_dst, rox, roy, roz, rdx, rdy, rdz);
    uint32_t ptr = 0, pidx = 0;
    uint32_t ptr_stack[VX_MAX_SCALE + 1], pidx_stack[VX_MAX_SCALE + 1];
    for (int i = 0; i <= VX_MAX_SCALE; ++i) { ptr_stack[i] = 0; pidx_stack[i] = 0; }
    uint32_t n = 0, last_leaf = 0xffffffffu;
    bool hit = false;
    Leaf g; float tex_lod = 0.0f; float4 color = make_float4(0, 0, 0, 0);
    for (;;) {
        const uint32_t oi = w.ci;
        const int scale_before = walk_scale(w);
        const uint32_t rec_before = w.rec;
        // the frame the shader emits at :175 for this iteration (if it gets past :152-156)
        if (w.state > 0 && !(w.t_min > w.limit)) {
            if (n < a.frames_cap) {
                VxDebugFrame& f = a.frames[n];
                f.t_min = w.t_min * inv_scale; f.idx = oi; f.scale = scale_before;
                if (FMT == VX_FMT_CSVO) {   // svo.csvo.glsl:246: (ptr, depth) ride in (ptr, parent_octant_idx); next_ptr is INVALID_PTR for no child
                    const bool child = csvo_has_child(w.hdr, w.desc, oi);
                    bool crossed = false;
                    const uint32_t np = child ? csvo_next_ptr(s, w.rec, w.desc, w.hdr, oi, crossed) : 0xffffffffu;
                    f.ptr = w.rec; f.parent_octant_idx = w.desc;
                    f.is_child = np != 0xffffffffu; f.is_leaf = f.is_child && w.desc < 2u;
                    f.crossed_boundary = crossed; f.next_ptr = np;
                } else {
                    f.ptr = ptr; f.parent_octant_idx = pidx;
                    f.is_child = ((w.desc >> oi) & 0x100u) != 0; f.is_leaf = ((w.desc >> oi) & 1u) != 0;
                    f.crossed_boundary = 0; f.next_ptr = 0;
                }
            }
            ++n;
        }
        const float h_before = w.h;
        const float tcx = __fmaf_rn(w.px, w.tcx, -w.tbx), tcy = __fmaf_rn(w.py, w.tcy, -w.tby), tcz = __fmaf_rn(w.pz, w.tcz, -w.tbz);
        const float tc_max = tmin2(tmin2(tcx, tcy), tcz);
        if (w.state <= 0) break;                      // MAX_STEPS used up (:152)
        walk_step<FMT, true, false>(w, s, sm.stack, cnt);
        if (state_at_leaf(w.state)) {
            g.value = leaf_value<FMT>(w, s);
            leaf_geom(w, rox, roy, roz, rdx, rdy, rdz, inv_scale, g);
            int tex_id;
            leaf_texture(s, g, tex_id, tex_lod);
            uint32_t nf = 0;
            color = texture_lod<4>(s.tex, sm.unorm, g.u, g.v, tex_id, tex_lod, &nf);   // the shader samples in both modes (:237)
            const bool first_of_kind = !(w.flags & VX_FLAG_ADJACENT) || g.value != last_leaf;
            if ((color.w > 0.0f || !a.cast_translucent) && first_of_kind) { hit = true; break; }
            last_leaf = g.value; w.flags |= VX_FLAG_ADJACENT;
            walk_skip_leaf<FMT, true>(w, s, sm.stack);
        }
        if (w.state == ST_MISS) break;
        const int scale_after = walk_scale(w);
        if (scale_after == scale_before - 1) {        // PUSH happened
            if (tc_max < h_before) { ptr_stack[scale_before] = ptr; pidx_stack[scale_before] = pidx; }
            ptr = rec_before; pidx = oi;
        } else if (scale_after > scale_before) {      // POP happened
            ptr = ptr_stack[scale_after]; pidx = pidx_stack[scale_after];
        }
    }
    VxOctreeResult& o = *a.result;
    o.t = -1.0f; o.value = 0; o.face_id = 0; o.pos[0] = o.pos[1] = o.pos[2] = 0; o.uv[0] = o.uv[1] = 0;
    o.color[0] = o.color[1] = o.color[2] = o.color[3] = 0; o.lod = 0; o.inside_voxel = (w.flags >> 8) & 1u;
    if (hit) {
        o.t = g.dst; o.value = g.value; o.face_id = g.face_id;
        leaf_pos(w.t_min, rox, roy, roz, rdx, rdy, rdz, g, inv_scale, o.pos[0], o.pos[1], o.pos[2]);
        o.uv[0] = g.u; o.uv[1] = g.v; o.lod = tex_lod;
        o.color[0] = color.x; o.color[1] = color.y; o.color[2] = color.z; o.color[3] = color.w;
    }
    *a.n_frames = n;
}

// ---- cross-GPU frame flags (multi-GPU tile gather over peer memory) ----------------------------------------------------------
// A flag is a 32-bit frame counter in GPU 0's memory. signal: everything this GPU wrote before (its pixels in GPU 0's frame)
// is made visible system-wide, then the flag is raised. wait: spin (acquire, system scope) until every flag of the range
// reached `value`; gives up after ~2 s and records it in flags[63] so that a lost peer shows up as an error, not as a hang.
__global__ void flag_signal_kernel(unsigned int* flag, unsigned int value) {
    __threadfence_system();
#ifndef VX_HOST_EMULATION
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
#else
    *flag = value;
#endif
}
__global__ void flag_wait_kernel(unsigned int* flags, unsigned int first, unsigned int count, unsigned int value, unsigned int* error_word) {
    const unsigned int i = threadIdx.x;
    if (i < count) {
        unsigned int v = 0;
        unsigned int spins = 0;
        for (;;) {
#ifndef VX_HOST_EMULATION
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + first + i) : "memory");
#else
            v = flags[first + i];
#endif
            if ((int)(v - value) >= 0) break;
            __nanosleep(200);
            if (++spins > (1u << 23)) { atomicAdd(error_word, 1u); break; }
        }
    }
    __syncthreads();
    __threadfence_system();
}

// ---- occupied box of the SVO (Clip) -----------------------------------------------------------------------------------------
// One thread per cell of the octree level L = min(depth, 6) (chunk granularity for the default world). A thread follows its cell's
// path from the root and, if the path does not end in an empty child, grows the box by the cell it stopped at: its level-L cell,
// or the coarser cell of a leaf met on the way. Most threads stop after one or two node reads. bounds = {min xyz = 0xffffffff,
// max xyz = 0} before the launch. Runs on the upload stream after every change of the world buffer.
__global__ void bounds_init_kernel(uint32_t* bounds) {
    if (threadIdx.x < 8) bounds[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
}
template <int FMT>
__global__ void __launch_bounds__(256) svo_bounds_kernel(Scene s, uint32_t* bounds) {
    const float octree_scale = load_octree_scale<FMT>(s);
    const uint32_t depth = 127u - ((__float_as_uint(octree_scale) >> 23) & 0xffu);
    if (depth == 0u || depth > 23u) return;
    const uint32_t L = depth < 6u ? depth : 6u;
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (1u << (3u * L))) return;
    const uint32_t mask = (1u << L) - 1u;
    const uint32_t cx = id & mask, cy = (id >> L) & mask, cz = (id >> (2u * L)) & mask;
    Walk w;
    w.mat_ptr = 0xffffffffu; w.preleaf = 0xffffffffu; w.hdr = 0;
    if (FMT == VX_FMT_CSVO) {
        w.rec = __ldg(s.desc - 1);
        w.desc = depth;
        w.hdr = csvo_header(s, w.rec, w.desc);
    } else {
        const uint32_t w0 = ld_desc(s, 0), w4 = ld_desc(s, 4);
        w.desc = w0 & 0xffffu;
        const uint32_t root = (w4 & 0x80000000u) ? (4u + (w4 & 0x7fffffffu)) : w4;
        w.rec = root < s.max_rec ? root : s.max_rec;
    }
    for (uint32_t k = 1; k <= L; ++k) {
        const uint32_t sh = L - k;
        w.ci = ((cx >> sh) & 1u) | (((cy >> sh) & 1u) << 1) | (((cz >> sh) & 1u) << 2);
        bool is_child, is_leaf;
        walk_decode<FMT>(w, is_child, is_leaf);
        if (!is_child) return;
        if (is_leaf || k == L) {
            // only ONE thread per stopping cell reports it: the one whose finer coordinates are all zero
            if (((cx | cy | cz) & ((1u << sh) - 1u)) != 0u) return;
            const uint32_t cell = 1u << (depth - k);
            const uint32_t x = (cx >> sh) * cell, y = (cy >> sh) * cell, z = (cz >> sh) * cell;
            atomicMin(bounds + 0, x); atomicMin(bounds + 1, y); atomicMin(bounds + 2, z);
            atomicMax(bounds + 3, x + cell); atomicMax(bounds + 4, y + cell); atomicMax(bounds + 5, z + cell);
            return;
        }
        walk_descend<FMT>(w, s);
    }
}

// ---- read-bandwidth probe (the roofline denominators: measured, not quoted) ------------------------------------------------------
// Every thread streams 16-byte words of a buffer, `passes` times; a buffer that fits the L2 measures L2 read bandwidth after the
// first pass, a buffer several times the L2 measures HBM. The sum goes to `sink` so that the loads cannot be dropped.
__global__ void __launch_bounds__(256) read_probe_kernel(const uint4* buf, unsigned long long n16, uint32_t passes, uint32_t* sink) {
    uint32_t acc = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (uint32_t p = 0; p < passes; ++p)
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i);   // cache at L2 only: an L1 hit would measure the wrong level
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}

// ---- small utility kernels ---------------------------------------------------------------------------------------------

// unorm[b] = b / 255.0f with an IEEE division (what unpacking an RGBA8 texel means), once per context
__global__ void unorm_kernel(float* table) { table[threadIdx.x] = (float)threadIdx.x / 255.0f; }

// glGenerateMipmap stand-in (texture_array.rs:258-260): level l+1 texel = floor(mean of the 2x2 block below + 1/4), the rounding that
// reproduces the reference's golden image of the mip-mapped far field
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
        out |= ((sum + 1) >> 2) << (8 * c);
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
    out[i] = pack_rgba8(frame[i]);
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

// Applies a packed dirty set — n VxRange headers, then the payload [head bytes of the world buffer][range 0 bytes][range 1 bytes]... —
// to the world buffer of a replica. The payload is byte-packed and ranges may sit at any byte offset (CSVO ranges are not
// word-aligned): a thread moves one 4-byte group of the payload, as one word when source and destination are both aligned and
// the group lies inside one range, byte by byte otherwise. A range that does not fit the buffer (offset + length + head >
// capacity) is skipped and counted in *error_word: a corrupt header must not become an out-of-bounds device write.
__global__ void scatter_ranges_kernel(uint8_t* world, const uint8_t* packed, uint32_t n_ranges, unsigned long long payload_bytes, uint32_t head_bytes,
                                      unsigned long long capacity, unsigned int* error_word) {
    const VxRange* hdr = reinterpret_cast<const VxRange*>(packed);
    const uint8_t* payload = packed + (size_t)n_ranges * sizeof(VxRange);
    const unsigned long long groups = (payload_bytes + 3) / 4;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const bool src_aligned = (reinterpret_cast<unsigned long long>(payload) & 3ull) == 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += stride) {
        const unsigned long long b0 = i * 4, b1 = (b0 + 4 < payload_bytes) ? b0 + 4 : payload_bytes;
        if (b1 <= head_bytes) {                                               // the head (scale + preamble / root pointer): word-aligned
            for (unsigned long long b = b0; b < b1; ++b) world[b] = payload[b];
            continue;
        }
        unsigned long long off = head_bytes;                                  // payload offset of range k
        uint32_t k = 0;
        for (unsigned long long b = b0; b < b1; ++b) {
            if (b < head_bytes) { world[b] = payload[b]; continue; }
            while (k < n_ranges && b >= off + hdr[k].length) { off += hdr[k].length; ++k; }
            if (k >= n_ranges) break;                                         // payload_bytes longer than the headers say: ignore the rest
            const unsigned long long ro = hdr[k].offset, rl = hdr[k].length;
            if (ro + rl + head_bytes > capacity || ro + rl < ro) {            // would leave the buffer: skip the range, count it once
                if (b == off) atomicAdd(error_word, 1u);
                continue;
            }
            const unsigned long long dst = head_bytes + ro + (b - off);
            if (b == b0 && b0 + 4 == b1 && src_aligned && (dst & 3ull) == 0 && b0 + 4 <= off + rl) {
                *reinterpret_cast<uint32_t*>(world + dst) = *reinterpret_cast<const uint32_t*>(payload + b0);
                break;
            }
            world[dst] = payload[b];
        }
    }
}

}  // namespace vx
