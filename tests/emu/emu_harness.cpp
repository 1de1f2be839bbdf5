// emu_harness.cpp — runs the product's CUDA kernels (voxel-rs_b200/csrc/kernels.cuh, compiled for the host through the stand-in
// cuda_runtime.h of this directory) on CPU fibers. TEST INFRASTRUCTURE: tests/test_kernels_emulated.py compares what comes out with
// the oracle. The host-side sequence below restates launch_wavefront / prepare_render / vx_set_textures / launch_raycast of
// csrc/voxelrt.cu (which cannot be compiled without nvcc: it uses <<< >>> and the CUDA runtime).
#include <cuda_runtime.h>   // the stand-in

#include <cstdio>
#include <functional>
#include <vector>

#include "../../voxel-rs_b200/csrc/kernels.cuh"
#include "../../voxel-rs_b200/csrc/chunks.cuh"

namespace emu {
Lane* cur = nullptr;
ucontext_t scheduler;
char* smem_base = nullptr;
unsigned long long collectives = 0, switches = 0;
static std::function<void()> kernel_body;
static std::vector<char> stacks;
static const size_t STACK_BYTES = 256 * 1024;

static void fiber_main() {
    kernel_body();
    Lane* me = cur;
    me->done = true;
    // a lane that leaves while others wait for it would hang them: complete what it was the last one missing from
    Warp* w = me->warp;
    if (--w->live && w->arrived == w->live) { std::memcpy(w->out[w->gen & 1u], w->in, sizeof(w->in)); w->arrived = 0; ++w->gen; }
    Cta* c = me->cta;
    if (--c->live && c->arrived == c->live) { c->arrived = 0; ++c->gen; }
    // returning switches to uc_link = the scheduler
}

// One kernel launch: CTAs run one after the other, the `threads` CUDA threads of a CTA as cooperating fibers.
static void launch(unsigned grid, unsigned threads, const std::function<void()>& body) {
    kernel_body = body;
    if (stacks.size() < (size_t)threads * STACK_BYTES) stacks.resize((size_t)threads * STACK_BYTES);
    const unsigned n_warps = (threads + 31) / 32;
    std::vector<Warp> warps(n_warps);
    std::vector<Lane> lanes(threads);
    for (unsigned b = 0; b < grid; ++b) {
        Cta cta;
        cta.live = threads;
        for (unsigned w = 0; w < n_warps; ++w) { warps[w] = Warp(); warps[w].live = std::min(32u, threads - 32u * w); }
        for (unsigned t = 0; t < threads; ++t) {
            Lane& l = lanes[t];
            l = Lane();
            l.tid.x = t; l.bid.x = b; l.bdim.x = threads; l.gdim.x = grid;
            l.lane = t & 31u; l.warp = &warps[t >> 5]; l.cta = &cta;
            getcontext(&l.ctx);
            l.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * STACK_BYTES;
            l.ctx.uc_stack.ss_size = STACK_BYTES;
            l.ctx.uc_link = &scheduler;
            makecontext(&l.ctx, fiber_main, 0);
        }
        for (;;) {
            bool alive = false, progressed = false;
            for (Lane& l : lanes) {
                if (l.done) continue;
                alive = true;
                if (l.wait_word && *l.wait_word == l.wait_value) continue;
                l.wait_word = nullptr;
                cur = &l;
                swapcontext(&scheduler, &l.ctx);
                progressed = true;
            }
            if (!alive) break;
            if (!progressed) { std::fprintf(stderr, "emu: deadlock in CTA %u (a collective not reached by every live lane)\n", b); std::abort(); }
        }
    }
    cur = nullptr;
}
}  // namespace emu

namespace vx {
thread_local uint32_t smem_raw[64 * 1024];   // the CTA's dynamic shared memory (extern __shared__ uint32_t smem_raw[] in the kernels)
thread_local __attribute__((aligned(16))) unsigned char chunk_smem_raw[sizeof(ChunkSmem) + 64];   // ... of serialize_chunks_kernel
}

using namespace vx;

namespace {

struct Device {   // what VxCtx holds on the GPU
    std::vector<uint32_t> world;      // 8 bytes of slack in front are not needed here; guard words behind the capacity are
    size_t capacity = 0;
    uint32_t fmt = 0, depth = 0;
    std::vector<Material> materials;
    std::vector<uint32_t> texels;
    TexInfo texinfo{};
    float unorm[256];
    unsigned long long opaque_materials = 0;
    uint32_t bounds[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0, 0, 0};   // svo_bounds_kernel (refresh_bounds in voxelrt.cu)
};

void refresh_bounds(Device& d);
void upload(Device& d, const uint8_t* world, uint64_t world_bytes, int fmt, uint32_t depth, const VxMaterial* materials, uint32_t n_materials,
            const uint8_t* rgba8, uint32_t tw, uint32_t th, uint32_t layers, uint32_t mip_levels) {
    d.fmt = (uint32_t)fmt; d.depth = depth;
    d.capacity = ((size_t)world_bytes + 255) / 4 * 4;
    d.world.assign(d.capacity / 4 + 16, 0u);                                  // + the zero guard words of vx_create (cap + 64 bytes)
    std::memcpy(d.world.data(), world, world_bytes);
    static_assert(sizeof(Material) == sizeof(VxMaterial), "material layout");
    d.materials.resize(n_materials);
    std::memcpy(d.materials.data(), materials, n_materials * sizeof(Material));
    emu::smem_base = reinterpret_cast<char*>(vx::smem_raw);
    emu::launch(1, 256, [&] { unorm_kernel(d.unorm); });
    // vx_set_textures: level count, offsets, mip chain and opacity bits (voxelrt.cu)
    uint32_t m = tw < th ? tw : th, il = 0;
    while ((m >> (il + 1)) != 0) ++il;
    uint32_t levels = mip_levels < il ? mip_levels : il;
    if (levels < 1) levels = 1;
    if (levels > 16) levels = 16;
    size_t total = 0;
    uint32_t off[16] = {};
    for (uint32_t l = 0; l < levels; ++l) {
        uint32_t wl = tw >> l, hl = th >> l;
        wl = wl ? wl : 1; hl = hl ? hl : 1;
        off[l] = (uint32_t)total;
        total += (size_t)wl * hl * layers;
    }
    d.texels.assign(total, 0u);
    std::memcpy(d.texels.data(), rgba8, (size_t)tw * th * layers * 4);
    for (uint32_t l = 1; l < levels; ++l) {
        uint32_t pw = tw >> (l - 1), ph = th >> (l - 1), cw = tw >> l, ch = th >> l;
        pw = pw ? pw : 1; ph = ph ? ph : 1; cw = cw ? cw : 1; ch = ch ? ch : 1;
        const uint32_t n = cw * ch * layers;
        emu::launch((n + 255) / 256, 256, [&] { mip_kernel(d.texels.data() + off[l - 1], d.texels.data() + off[l], pw, ph, cw, ch, layers); });
    }
    d.texinfo = TexInfo{};
    d.texinfo.texels = d.texels.data(); d.texinfo.w = tw; d.texinfo.h = th; d.texinfo.layers = layers; d.texinfo.levels = levels;
    for (int i = 0; i < 16; ++i) d.texinfo.off[i] = off[i];
    d.texinfo.opaque_layers = layers >= 64 ? ~0ull : ((1ull << layers) - 1ull);
    for (uint32_t l = 0; l < levels; ++l) {
        uint32_t wl = tw >> l, hl = th >> l;
        wl = wl ? wl : 1; hl = hl ? hl : 1;
        const uint32_t n = wl * hl * layers;
        emu::launch((n + 255) / 256, 256, [&] { opaque_kernel(d.texels.data() + off[l], wl * hl, layers, &d.texinfo.opaque_layers); });
    }
    d.opaque_materials = 0;   // update_opaque_materials
    for (size_t i = 0; i < d.materials.size() && i < 64; ++i) {
        const int ids[3] = {d.materials[i].tex_top, d.materials[i].tex_side, d.materials[i].tex_bottom};
        bool ok = true;
        for (int id : ids) {
            const int layer = id < 0 ? 0 : (id >= (int)layers ? (int)layers - 1 : id);
            ok = ok && layer < 64 && ((d.texinfo.opaque_layers >> layer) & 1ull);
        }
        if (ok) d.opaque_materials |= 1ull << i;
    }
    refresh_bounds(d);
}

Scene make_scene(const Device& d);
void refresh_bounds(Device& d) {      // voxelrt.cu refresh_bounds
    const uint32_t L = d.depth < 6 ? d.depth : 6;
    const unsigned blocks = ((1u << (3 * L)) + 255) / 256;
    Scene s = make_scene(d);
    emu::launch(blocks ? blocks : 1, 256, [&] {
        if (d.fmt == VX_FMT_CSVO) svo_bounds_kernel<VX_FMT_CSVO>(s, d.bounds); else svo_bounds_kernel<VX_FMT_ESVO>(s, d.bounds);
    });
}

Scene make_scene(const Device& d) {   // voxelrt.cu make_scene
    Scene s{};
    const size_t desc_off = d.fmt == VX_FMT_CSVO ? 8 : 4;
    s.desc = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(d.world.data()) + desc_off);
    s.desc_words = (uint32_t)((d.capacity - desc_off) / 4);
    s.format = d.fmt;
    s.max_rec = s.desc_words - 12;
    s.opaque_materials = d.opaque_materials;
    s.materials = d.materials.data(); s.n_materials = (uint32_t)d.materials.size();
    s.tex = &d.texinfo;
    s.unorm = d.unorm;
    const uint32_t levels = d.depth + (d.fmt == VX_FMT_CSVO ? 3 : 1);
    s.stack_levels = levels < 2 ? 2 : (levels > VX_MAX_SCALE ? VX_MAX_SCALE : levels);
    s.stack_max_off = (s.stack_levels - 1u) * VX_STACK_STRIDE;
    s.bounds = d.bounds;
    return s;
}

}  // namespace

// The library is built with -fvisibility=hidden -Wl,-Bsymbolic: its kernels carry the same C++ names as the host-side launch stubs
// that nvcc puts into libvoxelrt.so (loaded RTLD_GLOBAL by the package), and must not be interposed by them.
#define EMU_API __attribute__((visibility("default")))

// The ray-binning pre-pass exactly as launch_raycast issues it (voxelrt.cu): histogram + ranks, in-place exclusive scan, scatter.
static void bin_order(const float4* tasks, uint64_t n, const uint32_t* scale_word, uint32_t bin, std::vector<uint32_t>& order, std::vector<uint32_t>& keys) {
    const uint32_t bits_axis = bin & 15u, with_octant = (bin >> 4) & 1u;
    const uint32_t key_bits = bin_key_bits(bits_axis, with_octant);
    const size_t bins = (size_t)1 << (key_bits < 12 ? 12 : key_bits);
    std::vector<uint32_t> hist(bins, 0u), sums(bins / VX_SCAN_TILE, 0u);
    std::vector<uint2> keyrank(n);
    order.assign(n, 0xffffffffu);
    const unsigned g = (unsigned)std::min<uint64_t>((n + 255) / 256, 3);
    emu::launch(g ? g : 1, 256, [&] { bin_count_kernel(tasks, n, scale_word, bits_axis, with_octant, hist.data(), keyrank.data()); });
    const unsigned n_tiles = (unsigned)(bins / VX_SCAN_TILE);
    emu::launch(n_tiles, 256, [&] { bin_scan_reduce_kernel(hist.data(), sums.data()); });
    emu::launch(1, 256, [&] { bin_scan_sums_kernel(sums.data(), n_tiles); });
    emu::launch(n_tiles, 256, [&] { bin_scan_apply_kernel(hist.data(), sums.data()); });
    emu::launch(g ? g : 1, 256, [&] { bin_scatter_kernel(keyrank.data(), n, hist.data(), order.data()); });
    keys.resize(n);
    for (uint64_t i = 0; i < n; ++i) keys[i] = keyrank[i].x;
}


extern "C" {

// One frame through trace_primary_kernel -> shade_kernel -> trace_shadow_kernel exactly as vx_render issues them.
// options: [0] refill threshold, [1] shadow refill (0 = same), [2] CTAs of the persistent kernels, [3] count, [4] RGBA8 output,
// [5] bit 0 TMA-style tile write-back, bit 1 LIFO hand-over, [6] shard rank, [7] shard size, [8] bands (vx_render_read_rgba8's banded wavefront; 0/1 = whole frame)
EMU_API int emu_render(const uint8_t* world, uint64_t world_bytes, int fmt, uint32_t depth, const VxMaterial* materials, uint32_t n_materials,
               const uint8_t* tex_rgba8, uint32_t tw, uint32_t th, uint32_t layers, uint32_t mip_levels, const VxRenderParams* p, uint32_t width,
               uint32_t height, const uint32_t options[10], float* frame_out, uint32_t* frame8_out, uint64_t counters_out[6]) {
    Device d;
    upload(d, world, world_bytes, fmt, depth, materials, n_materials, tex_rgba8, tw, th, layers, mip_levels);
    RenderArgs a{};
    a.scene = make_scene(d);
    std::memcpy(a.u.view, p->view, sizeof(a.u.view));
    a.u.tan_half_fov = tanf(p->fov_y_rad * 0.5f);
    a.u.aspect = p->aspect_ratio; a.u.ambient = p->ambient_intensity;
    a.u.lx = p->light_dir[0]; a.u.ly = p->light_dir[1]; a.u.lz = p->light_dir[2];
    a.u.cx = p->cam_pos[0]; a.u.cy = p->cam_pos[1]; a.u.cz = p->cam_pos[2];
    a.u.hx = p->highlight_pos[0]; a.u.hy = p->highlight_pos[1]; a.u.hz = p->highlight_pos[2];
    a.u.render_shadows = p->render_shadows; a.u.shadow_distance = p->shadow_distance;
    a.u.width = width; a.u.height = height;
    set_macro_grid(a, width, height);
    const size_t slots = (size_t)a.macro_x * a.macro_y * 512;
    std::vector<float4> hit0(slots), hit1(slots), sh0(slots), sh1(slots), frame((size_t)width * height, float4{-1, -1, -1, -1});
    std::vector<uint32_t> sh_pix(slots), frame8((size_t)width * height, 0xdeadbeefu);
    Counters counters{};
    a.frame = frame.data();
    a.frame8 = options[4] ? frame8.data() : nullptr;
    a.hit0 = hit0.data(); a.hit1 = hit1.data(); a.sh0 = sh0.data(); a.sh1 = sh1.data(); a.sh_pix = sh_pix.data();
    a.counters = &counters;
    a.shard_rank = options[6]; a.shard_size = options[7] ? (options[7] & 0x7fffffffu) : 1; a.shard_rows = (options[7] >> 31) & 1u;
    a.refill_threshold = options[0] ? options[0] : 1;
    a.shadow_refill = options[1] ? options[1] : a.refill_threshold;
    a.tma_writeback = options[5] & 1u;
    a.lifo = ((options[5] >> 1) & 1u) ? 7u : 0u;   // LIFO hand-over of the wavefront buffers (discarded lines are poisoned here)
    // bands of macro-block rows, top of the image first, sizes shrinking by 0.6 (vx_render_read_rgba8); one band = vx_render
    uint32_t bands = options[8] ? options[8] : 1;
    if (bands > 16) bands = 16;
    if (bands > a.macro_y) bands = a.macro_y;
    const double ratio = 0.6;
    double wsum = 0.0, wk = 1.0;
    for (uint32_t k = 0; k < bands; ++k) { wsum += wk; wk *= ratio; }
    uint32_t edge[17];
    {
        double acc = 0.0; wk = 1.0;
        edge[0] = a.macro_y;
        for (uint32_t k = 0; k < bands; ++k) {
            acc += wk; wk *= ratio;
            uint32_t e = a.macro_y - (uint32_t)((double)a.macro_y * acc / wsum + 0.5);
            if (k + 1 == bands) e = 0;
            if (e > edge[k]) e = edge[k];
            edge[k + 1] = e;
        }
    }
    unsigned int work_all[16 * 8] = {};
    // options[9]: the overlapped wavefront's strip-completion flags (kernels run one after the other here, so every wait is
    // already satisfied — what is checked is that producers count exactly what consumers expect)
    std::vector<unsigned int> strip_done(slots / 128, 0u);
    unsigned int sync_errors = 0;
    if (options[9]) { a.strip_done = strip_done.data(); a.sync_errors = &sync_errors; }
    const bool count = options[3] != 0, csvo = fmt == VX_FMT_CSVO;
    const unsigned grid = options[2] ? options[2] : 3;
    for (uint32_t b = 0; b < bands; ++b) {
        const uint32_t row0 = edge[b + 1], row1 = edge[b];
        if (row1 == row0) continue;
        // launch_wavefront(c, a, shadows, band b, row0, row1)
        unsigned int* work = work_all + b * 8;
        shard_band(a, row0, row1);
        const uint32_t owned = a.n_owned;
        a.shadow_count = work + 4;
        a.fetch_tiles = 1;
        if (!owned) continue;
        a.work_counter = work;
        emu::launch(grid, VX_THREADS, [&] {
            if (csvo) { if (count) trace_primary_kernel<VX_FMT_CSVO, true, 8>(a); else trace_primary_kernel<VX_FMT_CSVO, false, 8>(a); }
            else { if (count) trace_primary_kernel<VX_FMT_ESVO, true, 8>(a); else trace_primary_kernel<VX_FMT_ESVO, false, 8>(a); }
        });
        a.shade_blocks = owned * 4;
        a.shade_counter = options[9] ? work + 6 : nullptr;   // overlapped: persistent shade grid (here: 3 CTAs, one after the other)
        if (options[9]) emu::launch(owned * 4 < 3 ? owned * 4 : 3, VX_THREADS, [&] { if (count) shade_kernel<true, true>(a); else shade_kernel<false, true>(a); });
        else emu::launch((owned * 4 + VX_SHADE_STRIPS - 1) / VX_SHADE_STRIPS, VX_THREADS, [&] { if (count) shade_kernel<true, false>(a); else shade_kernel<false, false>(a); });
        if (p->render_shadows) {
            a.work_counter = work + 2;
            emu::launch(grid, VX_THREADS, [&] {
                if (csvo) { if (count) trace_shadow_kernel<VX_FMT_CSVO, true, 8>(a); else trace_shadow_kernel<VX_FMT_CSVO, false, 8>(a); }
                else { if (count) trace_shadow_kernel<VX_FMT_ESVO, true, 8>(a); else trace_shadow_kernel<VX_FMT_ESVO, false, 8>(a); }
            });
        }
    }
    if (frame_out) std::memcpy(frame_out, frame.data(), frame.size() * sizeof(float4));
    if (frame8_out) std::memcpy(frame8_out, frame8.data(), frame8.size() * 4);
    if (counters_out) std::memcpy(counters_out, &counters, sizeof(Counters));
    return sync_errors ? -7 : 0;
}

// vx_raycast: one trace_picker_kernel launch over n tasks. options: [0] refill threshold, [2] CTAs, [3] count
EMU_API int emu_raycast(const uint8_t* world, uint64_t world_bytes, int fmt, uint32_t depth, const VxMaterial* materials, uint32_t n_materials,
                const uint8_t* tex_rgba8, uint32_t tw, uint32_t th, uint32_t layers, uint32_t mip_levels, const VxPickerTask* tasks, uint64_t n,
                VxPickerResult* results, const uint32_t options[8], uint64_t counters_out[6]) {
    Device d;
    upload(d, world, world_bytes, fmt, depth, materials, n_materials, tex_rgba8, tw, th, layers, mip_levels);
    RaycastArgs a{};
    a.scene = make_scene(d);
    a.tasks = reinterpret_cast<const float4*>(tasks);
    a.results = reinterpret_cast<float4*>(results);
    a.n = n;
    Counters counters{};
    unsigned long long work = 0;
    a.counters = &counters;
    a.work_counter = &work;
    a.refill_threshold = options[0] ? options[0] : 24;
    const bool count = options[3] != 0, csvo = fmt == VX_FMT_CSVO;
    std::vector<uint32_t> order;
    if (options[4]) {   // ray binning (vx_set_option 15), the launches of launch_raycast
        std::vector<uint32_t> keys;
        bin_order(a.tasks, n, a.scene.desc - (csvo ? 2 : 1), options[4], order, keys);
        a.order = order.data();
    }
    emu::launch(options[2] ? options[2] : 3, VX_THREADS, [&] {
        if (csvo) { if (count) trace_picker_kernel<VX_FMT_CSVO, true>(a); else trace_picker_kernel<VX_FMT_CSVO, false>(a); }
        else { if (count) trace_picker_kernel<VX_FMT_ESVO, true>(a); else trace_picker_kernel<VX_FMT_ESVO, false>(a); }
    });
    if (counters_out) std::memcpy(counters_out, &counters, sizeof(Counters));
    return 0;
}

// The binning pre-pass alone: order[] (a permutation of the task indices) and every task's key. scale = the world's octree_scale.
EMU_API int emu_bin_order(const VxPickerTask* tasks, uint64_t n, float scale, uint32_t bin, uint32_t* order_out, uint32_t* keys_out) {
    uint32_t scale_word;
    std::memcpy(&scale_word, &scale, 4);
    std::vector<uint32_t> order, keys;
    bin_order(reinterpret_cast<const float4*>(tasks), n, &scale_word, bin, order, keys);
    std::memcpy(order_out, order.data(), n * 4);
    std::memcpy(keys_out, keys.data(), n * 4);
    return 0;
}

// vx_serialize_chunks_esvo: one CTA of VX_CHUNK_THREADS threads per chunk, bump-allocated output (voxelrt.cu)
EMU_API int emu_serialize_chunks(const uint32_t* blocks, uint32_t n_chunks, const uint8_t* lods, VxChunkInfo* infos_out, void* records_out,
                                 uint64_t records_capacity, uint64_t* total_bytes) {
    static_assert(sizeof(ChunkOut) == sizeof(VxChunkInfo), "VxChunkInfo layout");
    std::vector<uint32_t> out((size_t)n_chunks * 4681 * 12);
    std::vector<ChunkOut> infos(n_chunks);
    unsigned long long bump[2] = {0, 0};
    emu::launch(n_chunks, VX_CHUNK_THREADS, [&] {
        serialize_chunks_kernel(blocks, lods, n_chunks, out.data(), (unsigned long long)out.size(), bump, infos.data(), reinterpret_cast<unsigned int*>(bump + 1));
    });
    std::memcpy(infos_out, infos.data(), n_chunks * sizeof(ChunkOut));
    if ((unsigned int)bump[1]) return VX_E_CAPACITY;
    const uint64_t total = bump[0] * 4ull;
    if (total_bytes) *total_bytes = total;
    if (records_out) {
        if (total > records_capacity) return VX_E_CAPACITY;
        std::memcpy(records_out, out.data(), total);
    }
    return 0;
}

// scatter_ranges_kernel (vx_svo_commit's staged path / vx_svo_commit_packed_device): packed = [n VxRange | head | range bytes...]
EMU_API int emu_scatter_ranges(uint8_t* world, const uint8_t* packed, uint32_t n_ranges, uint64_t payload_bytes, uint32_t head_bytes, uint32_t blocks,
                               uint64_t capacity, unsigned int* errors) {
    emu::launch(blocks ? blocks : 1, 256, [&] { scatter_ranges_kernel(world, packed, n_ranges, payload_bytes, head_bytes, capacity, errors); });
    return 0;
}

// vx_pack_shard / vx_unpack_shard and Framebuffer::read_pixels' conversion
EMU_API int emu_shard_copy(float* frame, float* packed, uint32_t width, uint32_t height, uint32_t rank, uint32_t size, int pack) {
    const uint32_t macro_x = (width + 31) / 32, n_macros = macro_x * ((height + 15) / 16);
    const uint32_t owned = n_macros > rank ? (n_macros - rank + size - 1) / size : 0;
    if (!owned) return 0;
    emu::launch(owned, 128, [&] {
        if (pack) shard_copy_kernel<true>(reinterpret_cast<float4*>(frame), reinterpret_cast<float4*>(packed), width, height, macro_x, n_macros, rank, size);
        else shard_copy_kernel<false>(reinterpret_cast<float4*>(frame), reinterpret_cast<float4*>(packed), width, height, macro_x, n_macros, rank, size);
    });
    return (int)owned;
}
EMU_API void emu_rgba8(const float* frame, uint32_t* out, uint64_t n) {
    emu::launch((unsigned)((n + 255) / 256), 256, [&] { rgba8_kernel(reinterpret_cast<const float4*>(frame), out, n); });
}

EMU_API void emu_stats(uint64_t out[2]) { out[0] = emu::collectives; out[1] = emu::switches; }

}  // extern "C"
