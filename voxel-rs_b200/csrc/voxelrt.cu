// voxelrt.cu — C ABI of libvoxelrt (include/voxelrt.h): context, streams, uploads and kernel launches.
// Device code lives in traverse.cuh (step machine, shading) and kernels.cuh (__global__ entry points).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -shared -Xcompiler -fPIC
// (--fmad=false is part of the numeric contract, see traverse.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/voxelrt.h"
#include "kernels.cuh"
#include "chunks.cuh"

// ================================================================== host ==

using namespace vx;

static thread_local std::string g_create_error;

struct GridCacheEntry;
// Ray binning of picker batches (vx_set_option 15): value = bits per axis of the origin cell's Z-order code (1..8) + 16 to append the
// direction octant; batches below VX_BIN_MIN_RAYS are traced in task order (the pre-pass is five launches). Default OFF: measured a loss
// on BASELINE configs[3] (profiles/r02_picker_binning.md: the pre-pass costs 0.50 ms and the picker kernel itself gets 10 % slower).
#ifndef VX_BIN_DEFAULT
#define VX_BIN_DEFAULT 0u
#endif
static constexpr uint64_t VX_BIN_MIN_RAYS = 1ull << 16;

struct VxCtx {
    VxConfig cfg{};
    int sm_count = 0;
    std::string err;

    cudaStream_t s_render = nullptr, s_upload = nullptr, s_picker = nullptr, s_copy = nullptr, s_pick_in = nullptr;
    cudaStream_t s_aux = nullptr;         // shade_kernel of the overlapped wavefront: runs next to trace_primary_kernel, ordered by strip flags
    cudaEvent_t e_pre = nullptr, e_k2 = nullptr;
    unsigned int* d_strip_done = nullptr; // per 32x4-pixel strip: pixels whose hit record is written (overlapped wavefront)
    uint32_t* d_work_list = nullptr;      // Z-order list of macro blocks (vx_set_option 13)
    uint32_t morton_x = 0, morton_y = 0;
    uint32_t* d_bounds = nullptr;         // occupied box of the SVO in voxel units (svo_bounds_kernel), read by the trace kernels (Clip)
    cudaEvent_t e_band[16] = {};
    cudaStream_t own_streams[3] = {nullptr, nullptr, nullptr};   // the library's own streams while caller streams are installed
    cudaEvent_t e_upload = nullptr, e_render = nullptr, e_picker = nullptr;
    cudaEvent_t t0_render = nullptr, t1_render = nullptr, t0_picker = nullptr, t1_picker = nullptr;

    uint32_t fmt = VX_FMT_ESVO;       // SVO type of the world buffer (VX_FLAG_SVO_CSVO)
    size_t head = 24;                 // bytes in front of the RangeBuffer image: f32 scale + 20-B preamble (ESVO) / + u32 root offset (CSVO)
    uint8_t* d_world_raw = nullptr;   // allocation; GL byte 0 lives at d_world_raw + 8 so that records are 16-B aligned
    uint8_t* d_world = nullptr;
    uint8_t* h_mirror = nullptr;      // pinned, capacity bytes
    uint8_t* h_stage = nullptr;       // pinned staging block for async dirty uploads
    uint8_t* d_stage = nullptr;       // its device twin (allocated on first use)
    uint8_t* h_stage_b = nullptr; uint8_t* d_stage_b = nullptr;   // second pair (first use): vx_svo_commit alternates, so a commit never waits
    cudaEvent_t e_stage[2] = {nullptr, nullptr};                  // for the upload in front of it — only for the one two commits ago
    cudaStream_t s_stage = nullptr;                               // the staging DMA's own stream (always the library's): the packed dirty set
    cudaEvent_t e_staged[2] = {nullptr, nullptr};                 // travels while the previous frame still renders; only the scatter waits
    bool stage_used[2] = {false, false};
    uint32_t stage_idx = 1;
    size_t stage_cap = 0;
    uint64_t hot_off = 0, hot_len = 0;
    bool l2_installed = false;
    uint64_t l2_off = 0, l2_len = 0, l2_opt = 0;
    cudaStream_t l2_render = nullptr, l2_picker = nullptr;
    bool have_svo = false;

    Material* d_materials = nullptr;
    uint32_t n_materials = 0;
    uint32_t* d_texels = nullptr;
    TexInfo* d_texinfo = nullptr;     // texture array description read by the device-side sampler
    float* d_unorm = nullptr;         // b / 255.0f table
    uint32_t tex_layers = 0;

    float4* d_frame = nullptr;
    uint32_t* d_frame8 = nullptr;
    // vx_render_read_rgba8_begin / _end with two frames in flight: the second RGBA8 device frame (allocated on first use), the frame the
    // last vx_render_read_rgba8* rendered into, a "frame is in host memory" event per frame in flight, frames issued / waited for
    uint32_t* d_frame8_b = nullptr;
    uint32_t* last_frame8 = nullptr;
    cudaEvent_t e_copied[2] = {nullptr, nullptr};
    uint64_t rr_issued = 0, rr_waited = 0;
    bool frame32_stale = false;       // the last frame was rendered as RGBA8 only (vx_render_read_rgba8 / option 8)
    uint32_t frame_w = 0, frame_h = 0;
    uint32_t last_shard_rank = 0, last_shard_size = 1, last_shard_rows = 0;   // shard of the last render (vx_read_hit_records)
    float4* frame_target = nullptr;   // where finished pixels go: d_frame, or a peer GPU's framebuffer (vx_open_peer_frame)
    uint32_t* frame8_target = nullptr;    // RGBA8 output mode (vx_set_option 8): d_frame8, or a peer GPU's RGBA8 frame (vx_open_peer_frame)
    unsigned int* d_flags = nullptr;      // 64 frame flags of this ctx (the root's are mapped by its peers); [63] = wait timeouts, [62] = dirty ranges the scatter kernel refused
    unsigned int* flags_target = nullptr; // the flags this ctx signals / gates on: d_flags, or the root's (vx_open_peer_sync)
    bool gate_armed = false;              // next vx_render: wait for flags_target[gate_slot] >= gate_value between trace and shade
    cudaEvent_t gate_event = nullptr;     // next vx_render: wait for this event (another device's, vx_group_render) at the same place
    unsigned int gate_slot = 0, gate_value = 0;

    // wavefront buffers of the render path (kernels.cuh), sized for the padded pixel count of the largest frame seen
    float4 *d_hit0 = nullptr, *d_hit1 = nullptr, *d_sh0 = nullptr, *d_sh1 = nullptr;
    uint32_t* d_sh_pix = nullptr;
    size_t wave_slots = 0;
    cudaEvent_t t_wave[4] = {nullptr, nullptr, nullptr, nullptr};   // trace / shade / shadow boundaries of the last frame

    std::vector<VxMaterial> h_materials;
    unsigned long long opaque_layers = 0;   // per texture layer: every texel of every level has alpha > 0
    unsigned long long opaque_materials = 0;

    float4* d_tasks = nullptr;
    float4* d_results = nullptr;
    // ray binning scratch of the picker path (bin_count_kernel ...), allocated on first use
    uint32_t* d_bin_hist = nullptr; uint32_t bin_hist_bits = 0;
    uint32_t* d_bin_sums = nullptr;
    uint2* d_bin_keyrank = nullptr; uint32_t* d_bin_order = nullptr; uint64_t bin_cap = 0;

    Counters* d_counters = nullptr;   // [0] render, [1] raycast
    unsigned long long* d_work = nullptr;   // u64[4] picker run counter, u64[6] chunk bump pointer, u64[7] its overflow flag; from byte 64:
                                            // 16 bands x 8 u32 render counters ([0] primary tiles, [2] shadow runs, [4] shadow list length)
    VxFrameStats last_render{}, last_raycast{};
    bool render_timed = false, raycast_timed = false;

    // chunk serialization scratch (vx_serialize_chunks_esvo)
    uint32_t* d_chunk_in = nullptr; size_t chunk_in_cap = 0;
    uint32_t* d_chunk_out = nullptr; size_t chunk_out_cap = 0;
    ChunkOut* d_chunk_info = nullptr; uint8_t* d_chunk_lod = nullptr; size_t chunk_n_cap = 0;
    cudaEvent_t t0_chunk = nullptr, t1_chunk = nullptr;
    float last_chunk_ms = 0.0f;

    VxStats stats{};
    uint64_t launches = 0;
    std::vector<struct GridCacheEntry> grid_cache;

    // options (vx_set_option)
    uint64_t opt_count = 0, opt_ctas_per_sm = 0, opt_l2_window = 1, opt_refill = 1, opt_refill_picker = 20, opt_rgba8_out = 0, opt_tma = 0, opt_refill_shadow = 0, opt_overlap = 2, opt_clip = 1, opt_morton = 0, opt_lifo = 0, opt_bin = VX_BIN_DEFAULT;
    bool hits_discarded = false;      // the last frame's hit records were consumed destructively (LIFO hand-over, vx_set_option 14)
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
    const size_t desc_off = c->fmt == VX_FMT_CSVO ? 8 : 4;   // descriptors[] follows octree_scale (and root_ptr), svo.esvo.glsl:3-6 / svo.csvo.glsl:1-5
    s.desc = reinterpret_cast<const uint32_t*>(c->d_world + desc_off);
    s.desc_words = (uint32_t)((c->cfg.svo_capacity_bytes - desc_off) / 4);
    s.format = c->fmt;
    s.max_rec = s.desc_words - 12;
    s.opaque_materials = c->opaque_materials;
    s.materials = c->d_materials; s.n_materials = c->n_materials;
    s.tex = c->d_texinfo;
    s.unorm = c->d_unorm;
    uint32_t levels = c->stats.depth + (c->fmt == VX_FMT_CSVO ? 3 : 1);   // CSVO: the oracle's stack policy for out-of-spec descents
    s.stack_levels = levels < 2 ? 2 : (levels > VX_MAX_SCALE ? VX_MAX_SCALE : levels);
    s.stack_max_off = (s.stack_levels - 1u) * VX_STACK_STRIDE;
    s.bounds = c->opt_clip ? c->d_bounds : nullptr;
    return s;
}
static size_t stack_smem_bytes(const Scene& s) { return smem_bytes(s.stack_levels); }

// bit m: material m (< 64) has only fully opaque face textures (tex id -1 samples layer 0 like the GL layer clamp does)
static void update_opaque_materials(VxCtx* c) {
    c->opaque_materials = 0;
    if (!c->d_texels || c->h_materials.empty()) return;
    for (size_t m = 0; m < c->h_materials.size() && m < 64; ++m) {
        const int ids[3] = {c->h_materials[m].tex_top, c->h_materials[m].tex_side, c->h_materials[m].tex_bottom};
        bool ok = true;
        for (int id : ids) {
            int layer = id < 0 ? 0 : (id >= (int)c->tex_layers ? (int)c->tex_layers - 1 : id);
            ok = ok && layer < 64 && ((c->opaque_layers >> layer) & 1ull);
        }
        if (ok) c->opaque_materials |= 1ull << m;
    }
}

extern "C" {

const char* vx_build_info(void) { return "libvoxelrt sm_100a --fmad=false " __DATE__ " " __TIME__; }

const char* vx_last_error(const VxCtx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

uint64_t vx_launch_count(const VxCtx* ctx) { return ctx ? ctx->launches : 0; }

int vx_set_option(VxCtx* ctx, uint32_t option, uint64_t value);

int vx_create(const VxConfig* cfg, VxCtx** out) {
    if (!cfg || !out) return fail(nullptr, VX_E_ARG, "vx_create: null argument");
    if (cfg->svo_capacity_bytes < 256) return fail(nullptr, VX_E_ARG, "vx_create: svo_capacity_bytes too small");
    if (cfg->svo_capacity_bytes > (1ull << 34)) return fail(nullptr, VX_E_ARG, "vx_create: svo_capacity_bytes > 16 GiB (u32 word pointers)");
    if ((cfg->flags & VX_FLAG_SVO_CSVO) && cfg->svo_capacity_bytes > (1ull << 31))
        return fail(nullptr, VX_E_ARG, "vx_create: a CSVO buffer is limited to 2 GiB (31-bit byte pointers, csvo.rs:82,121)");
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
    CUC(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->s_pick_in, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->s_stage, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&c->e_pre, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->e_k2, cudaEventDisableTiming));
    for (int i = 0; i < 16; ++i) CUC(cudaEventCreateWithFlags(&c->e_band[i], cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->e_upload, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CUC(cudaEventCreateWithFlags(&c->e_copied[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CUC(cudaEventCreateWithFlags(&c->e_stage[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CUC(cudaEventCreateWithFlags(&c->e_staged[i], cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->e_render, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->e_picker, cudaEventDisableTiming));
    CUC(cudaEventCreate(&c->t0_render)); CUC(cudaEventCreate(&c->t1_render));
    CUC(cudaEventCreate(&c->t0_picker)); CUC(cudaEventCreate(&c->t1_picker));
    for (int i = 0; i < 4; ++i) CUC(cudaEventCreate(&c->t_wave[i]));
    const size_t cap = (size_t)cfg->svo_capacity_bytes;
    CUC(cudaMalloc(&c->d_world_raw, cap + 64));
    c->d_world = c->d_world_raw + 8;
    CUC(cudaMemsetAsync(c->d_world_raw, 0, cap + 64, c->s_upload));
    CUC(cudaHostAlloc(&c->h_mirror, cap, cudaHostAllocPortable));   // portable: every device of a VxGroup uploads from device 0's mirror
    std::memset(c->h_mirror, 0, cap < (1u << 20) ? cap : (1u << 20));
    c->stage_cap = cap < (64u << 20) ? cap : (64u << 20);
    CUC(cudaHostAlloc(&c->h_stage, c->stage_cap, cudaHostAllocPortable));
    if (cfg->max_width && cfg->max_height) {
        const size_t px = (size_t)cfg->max_width * cfg->max_height;
        CUC(cudaMalloc(&c->d_frame, px * sizeof(float4)));
        CUC(cudaMalloc(&c->d_frame8, px * 4));
    }
    if (cfg->max_rays) {
        CUC(cudaMalloc(&c->d_tasks, (size_t)cfg->max_rays * 48));
        CUC(cudaMalloc(&c->d_results, (size_t)cfg->max_rays * 48));
    }
    CUC(cudaMalloc(&c->d_flags, 64 * sizeof(unsigned int)));
    CUC(cudaMemsetAsync(c->d_flags, 0, 64 * sizeof(unsigned int), c->s_upload));
    CUC(cudaMalloc(&c->d_bounds, 8 * sizeof(uint32_t)));
    CUC(cudaMemsetAsync(c->d_bounds, 0, 8 * sizeof(uint32_t), c->s_upload));   // min = max = 0: an empty world until the first commit
    CUC(cudaMalloc(&c->d_unorm, 256 * sizeof(float)));
    unorm_kernel<<<1, 256, 0, c->s_upload>>>(c->d_unorm);
    // one block: [0..7] u64 (picker run counter at [4], chunk bump pointer at [6..7]) | 16 bands x 8 u32 render work counters | the
    // render Counters | the raycast Counters — the per-frame reset of work counters + render Counters is ONE memset
    static_assert(sizeof(Counters) == 48, "Counters layout");
    CUC(cudaMalloc(&c->d_work, 64 + 16 * 32 + 2 * sizeof(Counters)));
    CUC(cudaMemsetAsync(c->d_work, 0, 64 + 16 * 32 + 2 * sizeof(Counters), c->s_upload));
    c->d_counters = reinterpret_cast<Counters*>(reinterpret_cast<uint8_t*>(c->d_work) + 64 + 16 * 32);
    CUC(cudaEventRecord(c->e_upload, c->s_upload));
    CUC(cudaStreamSynchronize(c->s_upload));
    c->stats.capacity_bytes = cap;
    c->fmt = (cfg->flags & VX_FLAG_SVO_CSVO) ? VX_FMT_CSVO : VX_FMT_ESVO;
    c->head = c->fmt == VX_FMT_CSVO ? 8 : 24;
    c->opt_l2_window = (cfg->flags & VX_FLAG_NO_L2_WINDOW) ? 0 : 1;
#undef CUC
    *out = c;
    return VX_OK;
}

void vx_destroy(VxCtx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    if (c->frame_target) cudaIpcCloseMemHandle(c->frame_target);
    if (c->frame8_target) cudaIpcCloseMemHandle(c->frame8_target);
    if (c->flags_target) cudaIpcCloseMemHandle(c->flags_target);
    if (c->d_flags) cudaFree(c->d_flags);
    if (c->d_chunk_in) cudaFree(c->d_chunk_in);
    if (c->d_chunk_out) cudaFree(c->d_chunk_out);
    if (c->d_chunk_info) cudaFree(c->d_chunk_info);
    if (c->d_chunk_lod) cudaFree(c->d_chunk_lod);
    if (c->t0_chunk) cudaEventDestroy(c->t0_chunk);
    if (c->t1_chunk) cudaEventDestroy(c->t1_chunk);
    if (c->d_world_raw) cudaFree(c->d_world_raw);
    if (c->h_mirror) cudaFreeHost(c->h_mirror);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->h_stage_b) cudaFreeHost(c->h_stage_b);
    if (c->d_stage_b) cudaFree(c->d_stage_b);
    for (cudaEvent_t& e : c->e_stage) if (e) cudaEventDestroy(e);
    for (cudaEvent_t& e : c->e_staged) if (e) cudaEventDestroy(e);
    if (c->s_stage) cudaStreamDestroy(c->s_stage);
    if (c->d_materials) cudaFree(c->d_materials);
    if (c->d_texels) cudaFree(c->d_texels);
    if (c->d_texinfo) cudaFree(c->d_texinfo);
    if (c->d_unorm) cudaFree(c->d_unorm);
    if (c->d_bounds) cudaFree(c->d_bounds);
    if (c->d_work_list) cudaFree(c->d_work_list);
    if (c->d_frame) cudaFree(c->d_frame);
    if (c->d_frame8) cudaFree(c->d_frame8);
    if (c->d_frame8_b) cudaFree(c->d_frame8_b);
    for (cudaEvent_t& e : c->e_copied) if (e) cudaEventDestroy(e);
    if (c->d_hit0) cudaFree(c->d_hit0);
    if (c->d_hit1) cudaFree(c->d_hit1);
    if (c->d_sh0) cudaFree(c->d_sh0);
    if (c->d_sh1) cudaFree(c->d_sh1);
    if (c->d_sh_pix) cudaFree(c->d_sh_pix);
    if (c->d_bin_hist) cudaFree(c->d_bin_hist);
    if (c->d_bin_sums) cudaFree(c->d_bin_sums);
    if (c->d_bin_keyrank) cudaFree(c->d_bin_keyrank);
    if (c->d_bin_order) cudaFree(c->d_bin_order);
    for (cudaEvent_t ev : c->t_wave) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->e_band) if (ev) cudaEventDestroy(ev);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    if (c->s_pick_in) cudaStreamDestroy(c->s_pick_in);
    if (c->s_aux) cudaStreamDestroy(c->s_aux);
    if (c->e_pre) cudaEventDestroy(c->e_pre);
    if (c->e_k2) cudaEventDestroy(c->e_k2);
    if (c->d_strip_done) cudaFree(c->d_strip_done);
    if (c->d_tasks) cudaFree(c->d_tasks);
    if (c->d_results) cudaFree(c->d_results);
    if (c->d_work) cudaFree(c->d_work);
    cudaEvent_t evs[] = {c->e_upload, c->e_render, c->e_picker, c->t0_render, c->t1_render, c->t0_picker, c->t1_picker};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    // only streams the library created are destroyed (caller-owned ones installed by vx_set_streams are not)
    if (c->own_streams[0]) { c->s_render = c->own_streams[0]; c->s_upload = c->own_streams[1]; c->s_picker = c->own_streams[2]; }
    cudaStream_t ss[] = {c->s_render, c->s_upload, c->s_picker};
    for (cudaStream_t s : ss) if (s) cudaStreamDestroy(s);
    delete c;
}

// Runtime knobs for A/B measurements (not part of the reference surface); the list with defaults is in voxelrt.h.
int vx_set_option(VxCtx* ctx, uint32_t option, uint64_t value) {
    if (!ctx) return VX_E_ARG;
    switch (option) {
        case 3: ctx->opt_count = value; break;
        case 4: ctx->opt_ctas_per_sm = value; break;
        case 5: ctx->opt_l2_window = value; break;
        case 6: ctx->opt_refill = value < 1 ? 1 : (value > 32 ? 32 : value); break;
        case 7: ctx->opt_refill_picker = value < 1 ? 1 : (value > 32 ? 32 : value); break;
        case 8: ctx->opt_rgba8_out = value ? 1 : 0; break;
        case 9: ctx->opt_tma = value ? 1 : 0; break;
        case 10: ctx->opt_refill_shadow = value > 32 ? 32 : value; break;
        case 11: ctx->opt_overlap = value > 2 ? 2 : value; break;
        case 12: ctx->opt_clip = value ? 1 : 0; break;
        case 13: ctx->opt_morton = value ? 1 : 0; break;
        case 14: ctx->opt_lifo = value > 7 ? 7 : value; break;
        case 15:
            if (value && ((value & 15u) < 1 || (value & 15u) > 8 || (value >> 5) || bin_key_bits((uint32_t)(value & 15u), (uint32_t)((value >> 4) & 1u)) > 24))
                return fail(ctx, VX_E_ARG, "vx_set_option 15: bits per axis 1..8 (+16 for the direction octant), at most 24 key bits");
            ctx->opt_bin = value;
            break;
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
    c->h_materials.assign(materials, materials + count);
    update_opaque_materials(c);
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
    // texture array description + per-layer "every texel is opaque" bits (lets shadow rays skip the sampler)
    TexInfo ti{};
    ti.texels = c->d_texels; ti.w = width; ti.h = height; ti.layers = layers; ti.levels = levels;
    for (int i = 0; i < 16; ++i) ti.off[i] = off[i];
    ti.opaque_layers = layers >= 64 ? ~0ull : ((1ull << layers) - 1ull);
    if (!c->d_texinfo) CU(c, cudaMalloc(&c->d_texinfo, sizeof(TexInfo)));
    CU(c, cudaMemcpyAsync(c->d_texinfo, &ti, sizeof(TexInfo), cudaMemcpyHostToDevice, c->s_upload));
    for (uint32_t l = 0; l < levels; ++l) {
        uint32_t wl = width >> l, hl = height >> l;
        wl = wl ? wl : 1; hl = hl ? hl : 1;
        const uint32_t n = wl * hl * layers;
        opaque_kernel<<<(n + 255) / 256, 256, 0, c->s_upload>>>(c->d_texels + off[l], wl * hl, layers, &c->d_texinfo->opaque_layers);
        c->launches++;
    }
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->s_upload));
    CU(c, cudaMemcpy(&c->opaque_layers, &c->d_texinfo->opaque_layers, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    c->tex_layers = layers;
    update_opaque_materials(c);
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
    // (re)installed only when the window or the streams changed: cudaDeviceSetLimit + two stream attributes per commit were
    // tens of microseconds of host time on every frame of the multi-GPU loop
    if (c->l2_installed && c->l2_off == c->hot_off && c->l2_len == c->hot_len && c->l2_opt == c->opt_l2_window && c->l2_render == c->s_render &&
        c->l2_picker == c->s_picker) return;
    c->l2_installed = true; c->l2_off = c->hot_off; c->l2_len = c->hot_len; c->l2_opt = c->opt_l2_window; c->l2_render = c->s_render; c->l2_picker = c->s_picker;
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
        attr.accessPolicyWindow.base_ptr = c->d_world + c->head + c->hot_off;
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

// The occupied box of the world for ray clipping (traverse.cuh Clip): recomputed on the upload stream after every change of the
// world buffer, before the upload event that frames and ray batches wait for. 1024 CTAs, most threads stop after a read or two.
static int refresh_bounds(VxCtx* c, uint32_t depth) {
    bounds_init_kernel<<<1, 32, 0, c->s_upload>>>(c->d_bounds);   // (a kernel, not a 32-byte H2D copy: no copy-engine round trip in front of every frame)
    c->launches++;
    Scene s = make_scene(c);
    const uint32_t L = depth < 6 ? depth : 6;   // (the kernel derives the depth from the buffer's scale itself; this only sizes the grid)
    const unsigned blocks = depth >= 1 && depth <= 23 ? ((1u << (3 * L)) + 255) / 256 : 1024;
    if (c->fmt == VX_FMT_CSVO) svo_bounds_kernel<VX_FMT_CSVO><<<blocks, 256, 0, c->s_upload>>>(s, c->d_bounds);
    else svo_bounds_kernel<VX_FMT_ESVO><<<blocks, 256, 0, c->s_upload>>>(s, c->d_bounds);
    c->launches++;
    CU(c, cudaGetLastError());
    return VX_OK;
}

// Bulk (re)load of ranges straight from a pinned mirror (this context's, or — in a VxGroup — device 0's): one DMA per range, then
// wait, because the caller may rewrite the mirror as soon as the call returns.
static int vx_svo_commit_from(VxCtx* c, const uint8_t* mirror, float octree_scale, const VxRange* dirty, uint32_t n_dirty, uint64_t used_bytes,
                              uint32_t depth) {
    (void)octree_scale;   // already at byte 0 of the mirror
    CU(c, cudaSetDevice(c->cfg.device));
    for (uint32_t i = 0; i < n_dirty; ++i)
        if (dirty[i].offset + dirty[i].length + c->head > c->cfg.svo_capacity_bytes)
            return fail(c, VX_E_CAPACITY, "dst is not large enough: len=%llu range_start=%llu range_length=%llu", (unsigned long long)c->cfg.svo_capacity_bytes,
                        (unsigned long long)dirty[i].offset, (unsigned long long)dirty[i].length);
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_render, 0));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_picker, 0));
    if (n_dirty) {
        CU(c, cudaMemcpyAsync(c->d_world, mirror, c->head, cudaMemcpyHostToDevice, c->s_upload));
        for (uint32_t i = 0; i < n_dirty; ++i)
            CU(c, cudaMemcpyAsync(c->d_world + c->head + dirty[i].offset, mirror + c->head + dirty[i].offset, dirty[i].length, cudaMemcpyHostToDevice,
                                  c->s_upload));
        CU(c, cudaStreamSynchronize(c->s_upload));
        c->have_svo = true;
        { const int rb_ = refresh_bounds(c, depth); if (rb_) return rb_; }
    }
    CU(c, cudaEventRecord(c->e_upload, c->s_upload));
    c->stats.used_bytes = used_bytes; c->stats.depth = depth;
    install_l2_window(c);
    return VX_OK;
}

int vx_svo_commit(VxCtx* c, float octree_scale, const VxRange* dirty, uint32_t n_dirty, uint64_t used_bytes, uint32_t depth) {
    if (!c || (n_dirty && !dirty)) return fail(c, VX_E_ARG, "vx_svo_commit: null argument");
    const uint64_t cap = c->cfg.svo_capacity_bytes;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_dirty; ++i) {
        // reference: assert!(start + length < dst_len) with dst_len = capacity - 1 relative to byte 4 (esvo.rs:328-331, svo.rs:180-181)
        if (dirty[i].offset + dirty[i].length + c->head > cap)
            return fail(c, VX_E_CAPACITY, "dst is not large enough: len=%llu range_start=%llu range_length=%llu", (unsigned long long)cap,
                        (unsigned long long)dirty[i].offset, (unsigned long long)dirty[i].length);
        total += dirty[i].length;
    }
    CU(c, cudaSetDevice(c->cfg.device));
    std::memcpy(c->h_mirror, &octree_scale, 4);                                  // svo.rs:173-175
    // (the scatter kernel is byte-granular: CSVO ranges, which are not word-aligned, take the staged path too)
    const bool staged = total + c->head + (uint64_t)n_dirty * sizeof(VxRange) <= c->stage_cap;
    // Two staging pairs, used alternately: this commit only has to wait for the upload + scatter of the commit BEFORE the previous one
    // (an event of its own; normally long done), never for the stream — which, when the caller runs everything on one stream
    // (vx_set_streams), would mean waiting for the frame in flight (the reference's render_fence.wait() stall, svo.rs:178).
    uint8_t* h_stage = c->h_stage;
    uint8_t** d_stage_p = &c->d_stage;
    uint32_t sidx = 0;
    if (n_dirty && staged) {
        sidx = (c->stage_idx ^= 1u);
        if (sidx == 1) {
            if (!c->h_stage_b) CU(c, cudaHostAlloc(&c->h_stage_b, c->stage_cap, cudaHostAllocPortable));
            h_stage = c->h_stage_b; d_stage_p = &c->d_stage_b;
        }
        if (c->stage_used[sidx]) CU(c, cudaEventSynchronize(c->e_stage[sidx]));
    }
    // do not tear a frame / ray batch in flight (render_fence.wait(), svo.rs:178) — on the GPU timeline, not the CPU's
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_render, 0));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_picker, 0));
    if (n_dirty) {
        if (staged) {
            // the caller may rewrite the mirror as soon as we return: snapshot the dirty bytes into the pinned staging block,
            // packed as [n VxRange headers | head bytes | range bytes...], move it with ONE async DMA and let the scatter kernel
            // put the ranges in place (one copy + one launch instead of a DMA per range)
            const size_t hdr_bytes = (size_t)n_dirty * sizeof(VxRange);
            if (hdr_bytes + c->head + total > c->stage_cap) return fail(c, VX_E_CAPACITY, "vx_svo_commit: %u dirty ranges do not fit the staging block", n_dirty);
            if (!*d_stage_p) CU(c, cudaMalloc(d_stage_p, c->stage_cap));
            uint8_t* const d_stage = *d_stage_p;
            std::memcpy(h_stage, dirty, hdr_bytes);
            size_t off = hdr_bytes;
            std::memcpy(h_stage + off, c->h_mirror, c->head);
            off += c->head;
            for (uint32_t i = 0; i < n_dirty; ++i) {
                std::memcpy(h_stage + off, c->h_mirror + c->head + dirty[i].offset, dirty[i].length);
                off += dirty[i].length;
            }
            // the DMA goes to the staging stream: it does not wait for the frame in flight (nothing reads d_stage but the scatter
            // below), so it is over long before the upload stream gets to the scatter kernel — and it does not queue behind that
            // frame's read-back either (measured: in-stream it cost 0.14 ms per frame of the pipelined loop)
            CU(c, cudaMemcpyAsync(d_stage, h_stage, off, cudaMemcpyHostToDevice, c->s_stage));
            CU(c, cudaEventRecord(c->e_staged[sidx], c->s_stage));
            CU(c, cudaStreamWaitEvent(c->s_upload, c->e_staged[sidx], 0));
            const unsigned long long pb = c->head + total;
            const int blocks = (int)((pb / 4 + 255) / 256 < 4096 ? (pb / 4 + 255) / 256 : 4096);
            scatter_ranges_kernel<<<blocks > 0 ? blocks : 1, 256, 0, c->s_upload>>>(c->d_world, d_stage, n_dirty, pb, (uint32_t)c->head,
                                                                                     c->cfg.svo_capacity_bytes, c->d_flags + 62);
            c->launches++;
            CU(c, cudaGetLastError());
            CU(c, cudaEventRecord(c->e_stage[sidx], c->s_upload));
            c->stage_used[sidx] = true;
        } else {
            CU(c, cudaMemcpyAsync(c->d_world, c->h_mirror, c->head, cudaMemcpyHostToDevice, c->s_upload));
            for (uint32_t i = 0; i < n_dirty; ++i)
                CU(c, cudaMemcpyAsync(c->d_world + c->head + dirty[i].offset, c->h_mirror + c->head + dirty[i].offset, dirty[i].length,
                                      cudaMemcpyHostToDevice, c->s_upload));
            CU(c, cudaStreamSynchronize(c->s_upload));   // bulk (re)load straight from the mirror: must finish before the caller reuses it
        }
        c->have_svo = true;
        { const int rb_ = refresh_bounds(c, depth); if (rb_) return rb_; }
    }
    CU(c, cudaEventRecord(c->e_upload, c->s_upload));
    c->stats.used_bytes = used_bytes; c->stats.depth = depth;
    install_l2_window(c);
    return VX_OK;
}

int64_t vx_svo_pack_dirty(VxCtx* c, const VxRange* dirty, uint32_t n_dirty, void* out, uint64_t out_cap) {
    if (!c || (n_dirty && !dirty)) return VX_E_ARG;
    uint64_t need = (uint64_t)n_dirty * sizeof(VxRange) + c->head;
    for (uint32_t i = 0; i < n_dirty; ++i) {
        if (dirty[i].offset + dirty[i].length + c->head > c->cfg.svo_capacity_bytes) return fail(c, VX_E_CAPACITY, "vx_svo_pack_dirty: range outside buffer");
        need += dirty[i].length;
    }
    if (!out) return (int64_t)need;
    if (need > out_cap) return fail(c, VX_E_CAPACITY, "vx_svo_pack_dirty: need %llu bytes, have %llu", (unsigned long long)need, (unsigned long long)out_cap);
    uint8_t* p = (uint8_t*)out;
    std::memcpy(p, dirty, (size_t)n_dirty * sizeof(VxRange));
    p += (size_t)n_dirty * sizeof(VxRange);
    std::memcpy(p, c->h_mirror, c->head);
    p += c->head;
    for (uint32_t i = 0; i < n_dirty; ++i) { std::memcpy(p, c->h_mirror + c->head + dirty[i].offset, dirty[i].length); p += dirty[i].length; }
    return (int64_t)need;
}

int vx_svo_commit_packed_device(VxCtx* c, const void* packed_dev, uint32_t n_dirty, uint64_t payload_bytes, uint64_t used_bytes, uint32_t depth) {
    if (!c || !packed_dev) return fail(c, VX_E_ARG, "vx_svo_commit_packed_device: null argument");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_render, 0));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_picker, 0));
    const unsigned long long pb = payload_bytes;   // head bytes + range bytes
    // what can be checked without the headers (they are in device memory; the kernel checks every range against the capacity and
    // counts the ones it had to skip: vx_svo_scatter_errors)
    if (pb < c->head || pb - c->head > c->cfg.svo_capacity_bytes - c->head)
        return fail(c, VX_E_CAPACITY, "vx_svo_commit_packed_device: payload of %llu bytes does not fit a %llu-byte buffer", pb,
                    (unsigned long long)c->cfg.svo_capacity_bytes);
    if ((uint64_t)n_dirty > pb) return fail(c, VX_E_ARG, "vx_svo_commit_packed_device: %u ranges in a %llu-byte payload", n_dirty, pb);
    if (reinterpret_cast<uintptr_t>(packed_dev) & 7u) return fail(c, VX_E_ARG, "vx_svo_commit_packed_device: packed_dev must be 8-byte aligned (VxRange headers)");
    const int blocks = (int)((pb / 4 + 255) / 256 < 4096 ? (pb / 4 + 255) / 256 : 4096);
    scatter_ranges_kernel<<<blocks > 0 ? blocks : 1, 256, 0, c->s_upload>>>(c->d_world, (const uint8_t*)packed_dev, n_dirty, pb, (uint32_t)c->head,
                                                                             c->cfg.svo_capacity_bytes, c->d_flags + 62);
    c->launches++;
    CU(c, cudaGetLastError());
    { const int rb_ = refresh_bounds(c, depth); if (rb_) return rb_; }
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

static int persistent_grid_query(VxCtx* c, const void* kernel, int threads, size_t smem, int* grid);
static int check_scene(VxCtx* c, const char* who) {
    if (!c->have_svo) return fail(c, VX_E_STATE, "%s: no SVO committed (vx_svo_commit)", who);
    if (!c->d_materials) return fail(c, VX_E_STATE, "%s: no materials (vx_set_materials)", who);
    if (!c->d_texels) return fail(c, VX_E_STATE, "%s: no textures (vx_set_textures)", who);
    return VX_OK;
}

// Sets the dynamic shared memory limit of a persistent kernel and sizes its grid (SMs x resident CTAs). Both are cached per
// (kernel, smem, CTAs/SM option): cudaFuncSetAttribute + the occupancy query cost tens of microseconds per launch otherwise.
struct GridCacheEntry { const void* kernel; size_t smem; uint64_t opt; int grid; };
static int persistent_grid(VxCtx* c, const void* kernel, int threads, size_t smem, int* grid) {
    for (const GridCacheEntry& e : c->grid_cache)
        if (e.kernel == kernel && e.smem == smem && e.opt == c->opt_ctas_per_sm) { *grid = e.grid; return VX_OK; }
    {
        cudaError_t e_ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e_ != cudaSuccess) return fail(c, VX_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e_));
    }
    int rc_ = persistent_grid_query(c, kernel, threads, smem, grid);
    if (rc_ == VX_OK) c->grid_cache.push_back(GridCacheEntry{kernel, smem, c->opt_ctas_per_sm, *grid});
    return rc_;
}
static int persistent_grid_query(VxCtx* c, const void* kernel, int threads, size_t smem, int* grid) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return fail(c, VX_E_CUDA, "occupancy query: %s", cudaGetErrorString(e));
    if (per_sm < 1) per_sm = 1;
    if (c->opt_ctas_per_sm && (int)c->opt_ctas_per_sm < per_sm) per_sm = (int)c->opt_ctas_per_sm;
    *grid = per_sm * c->sm_count;
    return VX_OK;
}

// (Re)allocates the wavefront buffers for `slots` pixel slots (padded pixel count of the frame).
static int ensure_wave_buffers(VxCtx* c, size_t slots) {
    if (slots <= c->wave_slots) return VX_OK;
    CU(c, cudaStreamSynchronize(c->s_render));
    float4** bufs[] = {&c->d_hit0, &c->d_hit1, &c->d_sh0, &c->d_sh1};
    for (float4** b : bufs) {
        if (*b) cudaFree(*b);
        *b = nullptr;
        CU(c, cudaMalloc(b, slots * sizeof(float4)));
    }
    if (c->d_sh_pix) cudaFree(c->d_sh_pix);
    c->d_sh_pix = nullptr;
    CU(c, cudaMalloc(&c->d_sh_pix, slots * sizeof(uint32_t)));
    if (c->d_strip_done) cudaFree(c->d_strip_done);
    c->d_strip_done = nullptr;
    CU(c, cudaMalloc(&c->d_strip_done, (slots / 128) * sizeof(unsigned int)));
    c->wave_slots = slots;
    return VX_OK;
}

#ifndef VX_TOP_MINB
#define VX_TOP_MINB 8   // resident CTAs per SM the default build of the trace kernels is compiled for (64 registers); A/B builds set 9 (56)
#endif
// register budget variant of the persistent trace kernels: CTAs/SM the allocator must allow (vx_set_option 4)
static int pick_minb(const VxCtx* c) { return c->opt_ctas_per_sm == 0 ? 8 : (c->opt_ctas_per_sm <= 5 ? 5 : (c->opt_ctas_per_sm <= 7 ? 6 : 8)); }

// One wavefront (trace -> shade -> shadow) over the macro-block rows [row0, row1) of the frame, on the render stream.
// `band` selects the set of work counters (each band of a frame needs its own zeroed set).
static int launch_wavefront(VxCtx* c, RenderArgs a, bool shadows, uint32_t band, uint32_t row0, uint32_t row1, bool timed) {
    shard_band(a, row0, row1);
    const uint32_t owned = a.n_owned;
    a.work_list = nullptr;
    if (c->opt_morton && a.shard_size == 1 && row0 == 0 && row1 == a.macro_y) {   // A/B: Z-order enumeration of a whole frame
        if (c->morton_x != a.macro_x || c->morton_y != a.macro_y) {
            // the frame's macro blocks along a Z-order curve: cells of the enclosing 2^b x 2^b grid in curve order, those outside dropped
            std::vector<uint32_t> list;
            list.reserve((size_t)a.macro_x * a.macro_y);
            uint32_t m = a.macro_x > a.macro_y ? a.macro_x : a.macro_y, bits = 0;
            while ((1u << bits) < m) ++bits;
            for (uint32_t k = 0; k < (1u << (2 * bits)); ++k) {
                uint32_t x = 0, y = 0;
                for (uint32_t b = 0; b < bits; ++b) { x |= ((k >> (2 * b)) & 1u) << b; y |= ((k >> (2 * b + 1)) & 1u) << b; }
                if (x < a.macro_x && y < a.macro_y) list.push_back(y * a.macro_x + x);
            }
            CU(c, cudaStreamSynchronize(c->s_render));
            if (c->d_work_list) cudaFree(c->d_work_list);
            c->d_work_list = nullptr;
            CU(c, cudaMalloc(&c->d_work_list, list.size() * sizeof(uint32_t)));
            CU(c, cudaMemcpy(c->d_work_list, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            c->morton_x = a.macro_x; c->morton_y = a.macro_y;
        }
        a.work_list = c->d_work_list;
    }
    unsigned int* work = reinterpret_cast<unsigned int*>(c->d_work) + 16 + band * 8;   // [0] primary strips, [2] shadow runs, [4] shadow list length
    a.shadow_count = work + 4;
    const size_t smem = stack_smem_bytes(a.scene);
    const bool count = c->opt_count != 0;
    const int minb = pick_minb(c);
    void (*k1)(RenderArgs) = nullptr;
    void (*k3)(RenderArgs) = nullptr;
#define VX_PICK(K, F, C) (minb == 5 ? K<F, C, 5> : (minb == 6 ? K<F, C, 6> : K<F, C, VX_TOP_MINB>))
    if (c->fmt == VX_FMT_CSVO) {
        if (count) { k1 = VX_PICK(trace_primary_kernel, VX_FMT_CSVO, true); k3 = VX_PICK(trace_shadow_kernel, VX_FMT_CSVO, true); }
        else { k1 = VX_PICK(trace_primary_kernel, VX_FMT_CSVO, false); k3 = VX_PICK(trace_shadow_kernel, VX_FMT_CSVO, false); }
    } else {
        if (count) { k1 = VX_PICK(trace_primary_kernel, VX_FMT_ESVO, true); k3 = VX_PICK(trace_shadow_kernel, VX_FMT_ESVO, true); }
        else { k1 = VX_PICK(trace_primary_kernel, VX_FMT_ESVO, false); k3 = VX_PICK(trace_shadow_kernel, VX_FMT_ESVO, false); }
    }
#undef VX_PICK
    if (timed) CU(c, cudaEventRecord(c->t_wave[0], c->s_render));
    if (owned) {
        // 1. primary rays -> hit records
        int grid = 0;
        int rc = persistent_grid(c, (const void*)k1, VX_THREADS, smem, &grid);
        if (rc) return rc;
        const int need = (int)owned * 4;   // 16 warp tiles per macro block, 4 warps per CTA: never more CTAs than there is work for
        a.work_counter = work;
        {   // tiles per work fetch: 1 (32 rays) up to 8K frames — claiming 2 per atomicAdd measured 6 % slower at 4K (longer tail);
            // only beyond ~500 tiles per resident warp does the single counter need relief
            const uint64_t per_warp = (uint64_t)owned * 16 / ((uint64_t)(grid < need ? grid : need) * (VX_THREADS / 32));
            a.fetch_tiles = per_warp >= 512 ? 2 : 1;
        }
        // Overlapped wavefront: the shade kernel goes to its own stream, ordered behind everything BEFORE the tracing kernel but not
        // behind the tracing kernel itself; its CTAs wait per strip (RenderArgs::strip_done). The tracing kernel's CTAs are all
        // resident first and fetch work dynamically, so whatever SM slots shade CTAs take, tracing finishes and feeds them.
        cudaStream_t s2 = a.strip_done ? c->s_aux : c->s_render;
        if (a.strip_done) {
            CU(c, cudaEventRecord(c->e_pre, c->s_render));
            CU(c, cudaStreamWaitEvent(c->s_aux, c->e_pre, 0));
        }
        k1<<<grid < need ? grid : need, VX_THREADS, smem, c->s_render>>>(a);
        c->launches++;
        if (timed) CU(c, cudaEventRecord(c->t_wave[1], c->s_render));
        // (multi-GPU, peer frame) the pixels go to GPU 0's framebuffer: not before GPU 0 released the previous frame
        if (c->gate_armed) {
            unsigned int* f = c->flags_target ? c->flags_target : c->d_flags;
            flag_wait_kernel<<<1, 32, 0, s2>>>(f, c->gate_slot, 1, c->gate_value, c->d_flags + 63);
            c->launches++;
            c->gate_armed = false;
        }
        if (c->gate_event) {   // single-process group: the root's "previous frame consumed" event, no kernel needed
            CU(c, cudaStreamWaitEvent(s2, c->gate_event, 0));
            c->gate_event = nullptr;
        }
        // 2. shading -> final pixels + shadow ray list
        // Overlapped: shade CTAs WAIT for the tracing kernel, so the tracing kernel must always be able to run. Nothing orders the
        // dispatch of two kernels on two streams (measured with bands: both become runnable on an idle GPU at the same instant;
        // when the block scheduler took the shade grid first, its waiting CTAs filled the SMs and the grid's undispatched rest kept
        // the tracing kernel out: every wait timed out, once in ~300 frames). So in this mode the shade kernel is persistent as
        // well: a grid of at most 5 CTAs per SM that is resident AS A WHOLE — nothing of it is ever left to dispatch — and padded
        // with shared memory so that what remains of every SM always holds one tracing CTA (5 x (pad + static + 1 KB) + trace + 1 KB
        // <= 227 KB, a sixth does not fit; registers 5 x 6 K + 8 K of 64 K). Its CTAs claim strips in work order from a counter; the
        // tracing kernel fetches its work dynamically too, so any one resident tracing CTA completes all tracing work.
        size_t smem2 = smem_bytes(0, false);
        unsigned shade_grid = owned * 4;
        a.shade_blocks = owned * 4;
        a.shade_counter = nullptr;
        if (a.strip_done) {
            const size_t sm_total = (size_t)227 * 1024, per_cta_extra = 3584;   // shade: ~2.1 KB static + 1 KB the system reserves per CTA (+ slack)
            const size_t trace_cta = smem + 1024;
            size_t pad = sm_total > trace_cta + 5 * per_cta_extra ? (sm_total - trace_cta) / 5 - per_cta_extra : 0;
            pad = pad > 47 * 1024 ? 47 * 1024 : pad / 128 * 128;                 // <= 48 KB: no opt-in attribute needed
            if (pad > smem2) smem2 = pad;
            a.shade_counter = work + 6;
            const unsigned resident = (unsigned)c->sm_count * 5u;
            if (shade_grid > resident) shade_grid = resident;
        }
        if (a.shade_counter) {
            if (count) shade_kernel<true, true><<<shade_grid, VX_THREADS, smem2, s2>>>(a);
            else shade_kernel<false, true><<<shade_grid, VX_THREADS, smem2, s2>>>(a);
        } else {
            shade_grid = (shade_grid + VX_SHADE_STRIPS - 1u) / VX_SHADE_STRIPS;
            if (count) shade_kernel<true, false><<<shade_grid, VX_THREADS, smem2, s2>>>(a);
            else shade_kernel<false, false><<<shade_grid, VX_THREADS, smem2, s2>>>(a);
        }
        c->launches++;
        if (timed) CU(c, cudaEventRecord(c->t_wave[2], s2));
        if (a.strip_done) {
            CU(c, cudaEventRecord(c->e_k2, c->s_aux));
            CU(c, cudaStreamWaitEvent(c->s_render, c->e_k2, 0));
        }
        // 3. shadow rays -> final pixels (world.glsl:80-84)
        if (shadows) {
            rc = persistent_grid(c, (const void*)k3, VX_THREADS, smem, &grid);
            if (rc) return rc;
            a.work_counter = work + 2;
            k3<<<grid < need ? grid : need, VX_THREADS, smem, c->s_render>>>(a);
            c->launches++;
        }
        CU(c, cudaGetLastError());
    } else if (timed) {
        CU(c, cudaEventRecord(c->t_wave[1], c->s_render));
        CU(c, cudaEventRecord(c->t_wave[2], c->s_render));
    }
    c->gate_armed = false;   // a gate belongs to ONE frame: a rank that owns no macro block of it must not carry it into the next
    c->gate_event = nullptr;
    if (timed) CU(c, cudaEventRecord(c->t_wave[3], c->s_render));
    return VX_OK;
}

#define VX_MAX_BANDS 16
#define VX_OVERLAP_TILES 24

// Argument checks + everything of RenderArgs that does not depend on the band.
static int prepare_render(VxCtx* c, const VxRenderParams* p, uint32_t width, uint32_t height, const VxShard* shard, RenderArgs& a, const char* who) {
    if (!c || !p || !width || !height) return fail(c, VX_E_ARG, "%s: null/empty argument", who);
    if (!c->d_frame || (uint64_t)width * height > (uint64_t)c->cfg.max_width * c->cfg.max_height)
        return fail(c, VX_E_CAPACITY, "%s: %ux%u exceeds the %ux%u framebuffer reserved at vx_create", who, width, height, c->cfg.max_width,
                    c->cfg.max_height);
    if (shard && ((shard->world_size & ~VX_SHARD_ROWS) == 0 || shard->rank >= (shard->world_size & ~VX_SHARD_ROWS))) return fail(c, VX_E_ARG, "%s: bad shard", who);
    int rc = check_scene(c, who);
    if (rc) return rc;
    CU(c, cudaSetDevice(c->cfg.device));
    a.scene = make_scene(c);
    std::memcpy(a.u.view, p->view, sizeof(a.u.view));
    a.u.tan_half_fov = tanf(p->fov_y_rad * 0.5f);                         // world.glsl:115, hoisted
    a.u.aspect = p->aspect_ratio; a.u.ambient = p->ambient_intensity;
    a.u.lx = p->light_dir[0]; a.u.ly = p->light_dir[1]; a.u.lz = p->light_dir[2];
    a.u.cx = p->cam_pos[0]; a.u.cy = p->cam_pos[1]; a.u.cz = p->cam_pos[2];
    a.u.hx = p->highlight_pos[0]; a.u.hy = p->highlight_pos[1]; a.u.hz = p->highlight_pos[2];
    a.u.render_shadows = p->render_shadows; a.u.shadow_distance = p->shadow_distance;
    a.u.width = width; a.u.height = height;
    set_macro_grid(a, width, height);
    rc = ensure_wave_buffers(c, (size_t)a.macro_x * a.macro_y * 512);
    if (rc) return rc;
    a.frame = c->frame_target ? c->frame_target : c->d_frame;
    a.frame8 = c->opt_rgba8_out ? (c->frame8_target ? c->frame8_target : c->d_frame8) : nullptr;
    a.hit0 = c->d_hit0; a.hit1 = c->d_hit1; a.sh0 = c->d_sh0; a.sh1 = c->d_sh1; a.sh_pix = c->d_sh_pix;
    a.counters = c->d_counters;
    a.shard_rank = shard ? shard->rank : 0; a.shard_size = shard ? (shard->world_size & ~VX_SHARD_ROWS) : 1;
    a.shard_rows = shard ? ((shard->world_size & VX_SHARD_ROWS) ? 1u : 0u) : 0u;
    c->last_shard_rank = a.shard_rank; c->last_shard_size = a.shard_size; c->last_shard_rows = a.shard_rows;
    a.refill_threshold = (uint32_t)c->opt_refill;
    a.shadow_refill = (uint32_t)(c->opt_refill_shadow ? c->opt_refill_shadow : c->opt_refill);
    a.tma_writeback = (c->opt_tma && !c->frame_target) ? 1u : 0u;   // bulk stores only into the local framebuffer
    a.lifo = (uint32_t)c->opt_lifo;
    c->hits_discarded = (a.lifo & 2u) != 0;
    CU(c, cudaStreamWaitEvent(c->s_render, c->e_upload, 0));
    CU(c, cudaMemsetAsync(reinterpret_cast<unsigned int*>(c->d_work) + 16, 0, VX_MAX_BANDS * 32 + sizeof(Counters), c->s_render));   // work counters + render Counters
    // overlapped wavefront (vx_set_option 11, default on; needs "finish all 32 rays, then refill" = refill threshold 1 only for its
    // cost model, the flags themselves count pixels and work with any threshold)
    // (vx_set_option 11: 0 off, 1 on, 2 = default: on when the launch is small enough for kernel tails to matter — a shard of a
    // frame, a small frame: fewer than VX_OVERLAP_TILES warp tiles per resident warp; a whole 4K frame on one GPU has 55 and loses
    // 3 % to the signalling, 1/8 of it has 7 and gains 6 %, profiles/r02_overlap.md)
    {
        const uint32_t size = a.shard_size ? a.shard_size : 1;
        const uint64_t tiles = (uint64_t)a.macro_x * a.macro_y * 16 / size;
        const uint64_t resident_warps = (uint64_t)c->sm_count * 8 * (VX_THREADS / 32);
        const bool on = c->opt_overlap == 1 || (c->opt_overlap == 2 && tiles < VX_OVERLAP_TILES * resident_warps);
        a.strip_done = on ? c->d_strip_done : nullptr;
    }
    a.sync_errors = c->d_flags + 61;
    if (a.strip_done) CU(c, cudaMemsetAsync(c->d_strip_done, 0, (size_t)a.macro_x * a.macro_y * 4 * sizeof(unsigned int), c->s_render));
    return VX_OK;
}

int vx_render(VxCtx* c, const VxRenderParams* p, uint32_t width, uint32_t height, const VxShard* shard, float* rgba32f_out) {
    RenderArgs a{};
    int rc = prepare_render(c, p, width, height, shard, a, "vx_render");
    if (rc) return rc;
    CU(c, cudaEventRecord(c->t0_render, c->s_render));
    rc = launch_wavefront(c, a, p->render_shadows != 0, 0, 0, a.macro_y, true);
    if (rc) return rc;
    c->frame32_stale = c->opt_rgba8_out != 0;
    c->last_frame8 = c->d_frame8;
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

// vx_render + Framebuffer::read_pixels in one call, pipelined: the frame is rendered in `bands` bands of macro-block rows
// (bottom to top); as soon as a band is finished it is converted to RGBA8 and copied to the host on the copy stream while the
// next band is being traced, so only the last band's copy is exposed. rgba8_out should be pinned (cudaHostAlloc /
// cudaHostRegister) for the overlap to happen. Returns when the whole frame is in rgba8_out.
// Everything of vx_render_read_rgba8 except waiting for the copies (vx_group_render_read_rgba8 issues one of these per device
// and waits for all of them afterwards).
static int render_read_rgba8_issue(VxCtx* c, const VxRenderParams* p, uint32_t width, uint32_t height, const VxShard* shard, uint8_t* rgba8_out,
                                   uint32_t bands) {
    if (!rgba8_out) return fail(c, VX_E_ARG, "vx_render_read_rgba8: null output");
    if (c->rr_issued - c->rr_waited >= 2)
        return fail(c, VX_E_STATE, "vx_render_read_rgba8_begin: two frames are in flight already — vx_render_read_rgba8_end first");
    RenderArgs a{};
    int rc = prepare_render(c, p, width, height, shard, a, "vx_render_read_rgba8");
    if (rc) return rc;
    // A frame whose copies may still be running keeps its device frame: the next one renders into the other (double buffering; the
    // kernels of frame k+1 then run under the read-back of frame k with nothing shared between them).
    uint32_t* frame8 = c->d_frame8;
    if (c->rr_issued != c->rr_waited && c->last_frame8 == c->d_frame8) {
        if (!c->d_frame8_b) CU(c, cudaMalloc(&c->d_frame8_b, (size_t)c->cfg.max_width * c->cfg.max_height * 4));
        frame8 = c->d_frame8_b;
    }
    c->last_frame8 = frame8;
    a.frame8 = frame8;        // the caller wants RGBA8 on the host: the shade / shadow kernels store the rounded pixels themselves (same
                              // bytes as converting the RGBA32F frame, a quarter of the frame traffic, no conversion pass) — into THIS
                              // device's frame even while a peer frame is open (vx_open_peer_frame*): the copy below reads it
    a.tma_writeback = 0;
    if (bands < 1) bands = 1;
    if (bands > VX_MAX_BANDS) bands = VX_MAX_BANDS;
    if (bands > a.macro_y) bands = a.macro_y;
    CU(c, cudaEventRecord(c->t0_render, c->s_render));
    // Band sizes shrink geometrically in render order (top of the image first, ratio VX_BAND_RATIO): the copy of band k runs
    // under the tracing of band k+1, so only the copy of the LAST, smallest band is exposed. The ratio is the measured
    // copy-time / render-time of a 4K frame on this box (33 MB at ~33 GB/s vs 1.5 ms).
    const double ratio = 0.6;
    double wsum = 0.0, wk = 1.0;
    for (uint32_t k = 0; k < bands; ++k) { wsum += wk; wk *= ratio; }
    uint32_t edge[VX_MAX_BANDS + 1];
    {
        double acc = 0.0; wk = 1.0;
        edge[0] = a.macro_y;                                   // render order k = 0 is the top band: rows [edge[k+1], edge[k])
        for (uint32_t k = 0; k < bands; ++k) {
            acc += wk; wk *= ratio;
            uint32_t e = a.macro_y - (uint32_t)((double)a.macro_y * acc / wsum + 0.5);
            if (k + 1 == bands) e = 0;
            if (e > edge[k]) e = edge[k];
            edge[k + 1] = e;
        }
    }
    for (uint32_t b = 0; b < bands; ++b) {
        const uint32_t row0 = edge[b + 1], row1 = edge[b];
        if (row1 == row0) continue;
        rc = launch_wavefront(c, a, p->render_shadows != 0, b, row0, row1, false);
        if (rc) return rc;
        CU(c, cudaEventRecord(c->e_band[b], c->s_render));
        CU(c, cudaStreamWaitEvent(c->s_copy, c->e_band[b], 0));
        if (a.shard_rows && a.shard_size > 1) {
            // this shard's stripes of the band: macro rows first, first + size, ... — each a contiguous run of 16 * width pixels,
            // a constant stride apart: ONE strided DMA (plus one for a ragged last stripe) into the host frame all shards share
            RenderArgs t = a;
            shard_band(t, row0, row1);
            uint32_t n_rows = t.n_owned / a.macro_x;
            if (n_rows) {
                const size_t stripe = (size_t)16 * width * 4, pitch = stripe * a.shard_size, off = (size_t)t.first_owned * stripe;
                const uint32_t last = t.first_owned + (n_rows - 1) * a.shard_size;
                if (last * 16 + 16 > height) {   // ragged last stripe of the frame
                    const size_t lo = (size_t)last * stripe;
                    CU(c, cudaMemcpyAsync(rgba8_out + lo, reinterpret_cast<const uint8_t*>(frame8) + lo, (size_t)(height - last * 16) * width * 4,
                                          cudaMemcpyDeviceToHost, c->s_copy));
                    --n_rows;
                }
                if (n_rows)
                    CU(c, cudaMemcpy2DAsync(rgba8_out + off, pitch, reinterpret_cast<const uint8_t*>(frame8) + off, pitch, stripe, n_rows,
                                            cudaMemcpyDeviceToHost, c->s_copy));
            }
        } else {
            const uint32_t y0 = row0 * 16, y1 = row1 * 16 < height ? row1 * 16 : height;
            const unsigned long long px0 = (unsigned long long)y0 * width, n = (unsigned long long)(y1 - y0) * width;
            CU(c, cudaMemcpyAsync(rgba8_out + px0 * 4, frame8 + px0, n * 4, cudaMemcpyDeviceToHost, c->s_copy));
        }
    }
    CU(c, cudaEventRecord(c->t1_render, c->s_render));
    CU(c, cudaEventRecord(c->e_render, c->s_render));
    CU(c, cudaEventRecord(c->e_copied[c->rr_issued & 1u], c->s_copy));   // behind this frame's last copy
    c->rr_issued++;
    c->frame_w = width; c->frame_h = height;
    c->render_timed = false;
    c->frame32_stale = true;
    return VX_OK;
}

int vx_render_read_rgba8(VxCtx* c, const VxRenderParams* p, uint32_t width, uint32_t height, const VxShard* shard, uint8_t* rgba8_out,
                         uint32_t bands) {
    int rc = render_read_rgba8_issue(c, p, width, height, shard, rgba8_out, bands);
    if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->s_copy));
    c->rr_waited = c->rr_issued;
    return VX_OK;
}

int vx_render_read_rgba8_begin(VxCtx* c, const VxRenderParams* p, uint32_t width, uint32_t height, const VxShard* shard, uint8_t* rgba8_out,
                               uint32_t bands) {
    return render_read_rgba8_issue(c, p, width, height, shard, rgba8_out, bands);
}

int vx_render_read_rgba8_end(VxCtx* c) {
    if (!c) return VX_E_ARG;
    CU(c, cudaSetDevice(c->cfg.device));
    if (c->rr_waited < c->rr_issued) {   // the OLDEST frame in flight
        CU(c, cudaEventSynchronize(c->e_copied[c->rr_waited & 1u]));
        c->rr_waited++;
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
    if (c->opt_rgba8_out || c->frame32_stale)
        return fail(c, VX_E_STATE, "vx_read_frame_rgba32f: the last frame was rendered as RGBA8 only (vx_set_option 8 / vx_render_read_rgba8)");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaMemcpyAsync(out, c->d_frame, (size_t)c->frame_w * c->frame_h * sizeof(float4), cudaMemcpyDeviceToHost, c->s_render));
    CU(c, cudaStreamSynchronize(c->s_render));
    return VX_OK;
}

int vx_read_frame_rgba8(VxCtx* c, uint8_t* out) {
    if (!c || !out || !c->frame_w) return fail(c, VX_E_ARG, "vx_read_frame_rgba8: nothing rendered / null");
    CU(c, cudaSetDevice(c->cfg.device));
    const unsigned long long n = (unsigned long long)c->frame_w * c->frame_h;
    if (!c->opt_rgba8_out && !c->frame32_stale) {   // in RGBA8 output mode the kernels already stored the rounded pixels
        rgba8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->s_render>>>(c->d_frame, c->d_frame8, n);
        c->launches++;
        CU(c, cudaGetLastError());
    }
    CU(c, cudaMemcpyAsync(out, (c->frame32_stale && c->last_frame8) ? c->last_frame8 : c->d_frame8, n * 4, cudaMemcpyDeviceToHost, c->s_render));
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

int vx_read_hit_records(VxCtx* c, VxHitRecord* out) {
    if (!c || !out || !c->frame_w || !c->d_hit0) return fail(c, VX_E_ARG, "vx_read_hit_records: nothing rendered / null");
    if (c->hits_discarded)
        return fail(c, VX_E_STATE, "vx_read_hit_records: the frame was rendered with the LIFO hand-over, which discards the hit records as "
                                   "they are shaded; vx_set_option(ctx, 14, 0) before the render keeps them");
    CU(c, cudaSetDevice(c->cfg.device));
    const uint32_t w = c->frame_w, h = c->frame_h, macro_x = (w + 31) / 32, macro_y = (h + 15) / 16;
    const size_t slots = (size_t)macro_x * macro_y * 512;
    std::vector<float4> h0(slots), h1(slots);
    CU(c, cudaMemcpyAsync(h0.data(), c->d_hit0, slots * sizeof(float4), cudaMemcpyDeviceToHost, c->s_render));
    CU(c, cudaMemcpyAsync(h1.data(), c->d_hit1, slots * sizeof(float4), cudaMemcpyDeviceToHost, c->s_render));
    CU(c, cudaStreamSynchronize(c->s_render));
    const uint32_t size = c->last_shard_size ? c->last_shard_size : 1, rank = c->last_shard_rank;
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const uint32_t macro = (y / 16) * macro_x + x / 32;
            const size_t slot = ((size_t)macro * 4 + (y % 16) / 4) * 128 + ((x % 32) / 8) * 32 + (y % 4) * 8 + (x % 8);   // kernels.cuh strip_pixel
            VxHitRecord& o = out[(size_t)y * w + x];
            std::memset(&o, 0, sizeof(o));
            if ((c->last_shard_rows ? (y / 16) % size : macro % size) != rank) { o.t = -2.0f; continue; }
            uint32_t flags;
            std::memcpy(&flags, &h1[slot].w, 4);
            if (!(flags & 8u)) { o.t = -1.0f; continue; }
            o.t = h0[slot].x;
            std::memcpy(&o.value, &h0[slot].y, 4);
            o.face_id = (int32_t)(flags & 7u);
            o.pos[0] = h1[slot].x; o.pos[1] = h1[slot].y; o.pos[2] = h1[slot].z;
            o.uv[0] = h0[slot].z; o.uv[1] = h0[slot].w;
        }
    return VX_OK;
}

// One launch of the picker kernel over n tasks. A batch that is cut into several launches (vx_raycast's pipelined path) clears the
// ray counters and starts the timer with its first launch and stops it with its last.
static int launch_raycast(VxCtx* c, const float4* tasks_dev, uint64_t n, float4* results_dev, bool first = true, bool last = true) {
    RaycastArgs a{};
    a.scene = make_scene(c);
    a.tasks = tasks_dev; a.results = results_dev; a.n = n;
    a.counters = c->d_counters + 1;
    a.work_counter = c->d_work + 4;
    a.refill_threshold = (uint32_t)c->opt_refill_picker;
    const size_t smem = smem_bytes(a.scene.stack_levels, true, false);   // the picker kernel has no unorm table
    auto k = c->fmt == VX_FMT_CSVO ? (c->opt_count ? trace_picker_kernel<VX_FMT_CSVO, true> : trace_picker_kernel<VX_FMT_CSVO, false>)
                                   : (c->opt_count ? trace_picker_kernel<VX_FMT_ESVO, true> : trace_picker_kernel<VX_FMT_ESVO, false>);
    int grid = 0;
    int rc = persistent_grid(c, (const void*)k, VX_THREADS, smem, &grid);
    if (rc) return rc;
    const uint64_t need = (n + VX_THREADS - 1) / VX_THREADS;
    if ((uint64_t)grid > need) grid = (int)need;
    if (first) CU(c, cudaMemsetAsync(c->d_counters + 1, 0, sizeof(Counters), c->s_picker));
    CU(c, cudaMemsetAsync(c->d_work + 4, 0, sizeof(unsigned long long), c->s_picker));
    if (first) CU(c, cudaEventRecord(c->t0_picker, c->s_picker));
    // Ray binning (vx_set_option 15): large batches are traced in Z-order of their origin cells. Inside the timed region.
    if (c->opt_bin && n >= VX_BIN_MIN_RAYS && n < (1ull << 32)) {
        const uint32_t bits_axis = (uint32_t)(c->opt_bin & 15u), with_octant = (uint32_t)((c->opt_bin >> 4) & 1u);
        const uint32_t key_bits = bin_key_bits(bits_axis, with_octant);
        const size_t bins = (size_t)1 << (key_bits < 12 ? 12 : key_bits);   // a multiple of VX_SCAN_TILE
        if (c->bin_hist_bits < key_bits || !c->d_bin_hist) {
            CU(c, cudaStreamSynchronize(c->s_picker));
            if (c->d_bin_hist) cudaFree(c->d_bin_hist);
            if (c->d_bin_sums) cudaFree(c->d_bin_sums);
            c->d_bin_hist = nullptr; c->d_bin_sums = nullptr;
            CU(c, cudaMalloc(&c->d_bin_hist, bins * sizeof(uint32_t)));
            CU(c, cudaMalloc(&c->d_bin_sums, (bins / VX_SCAN_TILE) * sizeof(uint32_t)));
            c->bin_hist_bits = key_bits < 12 ? 12 : key_bits;
        }
        if (c->bin_cap < n) {
            CU(c, cudaStreamSynchronize(c->s_picker));
            if (c->d_bin_keyrank) cudaFree(c->d_bin_keyrank);
            if (c->d_bin_order) cudaFree(c->d_bin_order);
            c->d_bin_keyrank = nullptr; c->d_bin_order = nullptr; c->bin_cap = 0;
            const uint64_t cap = n > c->cfg.max_rays ? n : c->cfg.max_rays;
            CU(c, cudaMalloc(&c->d_bin_keyrank, cap * sizeof(uint2)));
            CU(c, cudaMalloc(&c->d_bin_order, cap * sizeof(uint32_t)));
            c->bin_cap = cap;
        }
        const uint32_t n_tiles = (uint32_t)(bins / VX_SCAN_TILE);
        const int sweep = c->sm_count * 8;
        const uint64_t blocks = (n + 255) / 256;
        const int g = blocks < (uint64_t)sweep ? (int)blocks : sweep;
        CU(c, cudaMemsetAsync(c->d_bin_hist, 0, bins * sizeof(uint32_t), c->s_picker));
        bin_count_kernel<<<g, 256, 0, c->s_picker>>>(tasks_dev, n, a.scene.desc - (c->fmt == VX_FMT_CSVO ? 2 : 1), bits_axis, with_octant, c->d_bin_hist,
                                                     c->d_bin_keyrank);
        bin_scan_reduce_kernel<<<n_tiles, 256, 0, c->s_picker>>>(c->d_bin_hist, c->d_bin_sums);
        bin_scan_sums_kernel<<<1, 256, 0, c->s_picker>>>(c->d_bin_sums, n_tiles);
        bin_scan_apply_kernel<<<n_tiles, 256, 0, c->s_picker>>>(c->d_bin_hist, c->d_bin_sums);
        bin_scatter_kernel<<<g, 256, 0, c->s_picker>>>(c->d_bin_keyrank, n, c->d_bin_hist, c->d_bin_order);
        c->launches += 5;
        a.order = c->d_bin_order;
    }
    k<<<grid, VX_THREADS, smem, c->s_picker>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    if (last) {
        CU(c, cudaEventRecord(c->t1_picker, c->s_picker));
        CU(c, cudaEventRecord(c->e_picker, c->s_picker));
        c->raycast_timed = true;
    }
    return VX_OK;
}

// Batches above this many rays are cut into slices: slice i's tasks go up (s_pick_in) while slice i-1 is traced (s_picker) and
// slice i-2's results come down (s_copy) — PCIe is full duplex, so a large batch costs about max(H2D, D2H) instead of their sum.
#ifndef VX_PICK_SLICE_LOG2
#define VX_PICK_SLICE_LOG2 20
#endif
static constexpr uint64_t VX_PICK_SLICE = 1ull << VX_PICK_SLICE_LOG2;   // 1 Mi rays = 48 MiB each way; 16 Mi rays end to end: 2^19 / 2^20 / 2^21 / 2^22 ->
                                                                        // 18.90 / 18.78 / 19.97 / 20.90 ms (profiles/r02_step_variants.md)

static int raycast_pipelined(VxCtx* c, const VxPickerTask* tasks, uint64_t n, VxPickerResult* results) {
    const uint64_t n_slices = (n + VX_PICK_SLICE - 1) / VX_PICK_SLICE;
    std::vector<cudaEvent_t> ev(2 * n_slices, nullptr);
    int rc = VX_OK;
    for (auto& e : ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { rc = fail(c, VX_E_CUDA, "vx_raycast: cudaEventCreate failed"); break; }
    for (uint64_t i = 0; i < n_slices && rc == VX_OK; ++i) {
        const uint64_t off = i * VX_PICK_SLICE, m = std::min(VX_PICK_SLICE, n - off);
        cudaError_t e = cudaMemcpyAsync(c->d_tasks + off * 3, tasks + off, (size_t)m * 48, cudaMemcpyHostToDevice, c->s_pick_in);
        if (e == cudaSuccess) e = cudaEventRecord(ev[2 * i], c->s_pick_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->s_picker, ev[2 * i], 0);
        if (e != cudaSuccess) { rc = fail(c, VX_E_CUDA, "vx_raycast: upload of slice %llu: %s", (unsigned long long)i, cudaGetErrorString(e)); break; }
        rc = launch_raycast(c, c->d_tasks + off * 3, m, c->d_results + off * 3, i == 0, i + 1 == n_slices);
        if (rc) break;
        e = cudaEventRecord(ev[2 * i + 1], c->s_picker);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->s_copy, ev[2 * i + 1], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(results + off, c->d_results + off * 3, (size_t)m * 48, cudaMemcpyDeviceToHost, c->s_copy);
        if (e != cudaSuccess) { rc = fail(c, VX_E_CUDA, "vx_raycast: read-back of slice %llu: %s", (unsigned long long)i, cudaGetErrorString(e)); break; }
    }
    // the fence of the reference (svo.rs:248-249); also on the error path, so that no copy is in flight into the caller's buffers
    const cudaError_t e0 = cudaStreamSynchronize(c->s_pick_in), e1 = cudaStreamSynchronize(c->s_picker), e2 = cudaStreamSynchronize(c->s_copy);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (rc == VX_OK && (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess))
        rc = fail(c, VX_E_CUDA, "vx_raycast: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2)));
    return rc;
}

int vx_raycast(VxCtx* c, const VxPickerTask* tasks, uint64_t n, VxPickerResult* results) {
    if (!c || (n && (!tasks || !results))) return fail(c, VX_E_ARG, "vx_raycast: null argument");
    if (n == 0) return VX_OK;
    if (n > c->cfg.max_rays) return fail(c, VX_E_CAPACITY, "vx_raycast: %llu rays exceed max_rays=%llu", (unsigned long long)n, (unsigned long long)c->cfg.max_rays);
    int rc = check_scene(c, "vx_raycast");
    if (rc) return rc;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamWaitEvent(c->s_picker, c->e_upload, 0));
    if (n > VX_PICK_SLICE) return raycast_pipelined(c, tasks, n, results);
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
    // one scratch block {result | frame count | frames}, freed on every way out
    const uint32_t cap = frames ? frames_cap : 0;
    const size_t off_n = (sizeof(VxOctreeResult) + 15) & ~(size_t)15, off_frames = off_n + 16;
    struct Scratch { uint8_t* p = nullptr; ~Scratch() { if (p) cudaFree(p); } } scratch;
    CU(c, cudaMalloc(&scratch.p, off_frames + sizeof(VxDebugFrame) * (cap ? cap : 1)));
    VxOctreeResult* d_res = reinterpret_cast<VxOctreeResult*>(scratch.p);
    uint32_t* d_n = reinterpret_cast<uint32_t*>(scratch.p + off_n);
    VxDebugFrame* d_frames = reinterpret_cast<VxDebugFrame*>(scratch.p + off_frames);
    DebugArgs a{};
    a.scene = make_scene(c);
    for (int k = 0; k < 3; ++k) { a.pos[k] = pos[k]; a.dir[k] = dir[k]; }
    a.max_dst = max_dst; a.cast_translucent = cast_translucent;
    a.result = d_res; a.frames = d_frames; a.frames_cap = cap; a.n_frames = d_n;
    CU(c, cudaStreamWaitEvent(c->s_picker, c->e_upload, 0));
    if (c->fmt == VX_FMT_CSVO) debug_cast_kernel<VX_FMT_CSVO><<<1, VX_THREADS, stack_smem_bytes(a.scene), c->s_picker>>>(a);
    else debug_cast_kernel<VX_FMT_ESVO><<<1, VX_THREADS, stack_smem_bytes(a.scene), c->s_picker>>>(a);
    c->launches++;
    CU(c, cudaGetLastError());
    uint32_t n = 0;
    CU(c, cudaMemcpyAsync(result, d_res, sizeof(VxOctreeResult), cudaMemcpyDeviceToHost, c->s_picker));
    CU(c, cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, c->s_picker));
    if (cap) CU(c, cudaMemcpyAsync(frames, d_frames, sizeof(VxDebugFrame) * cap, cudaMemcpyDeviceToHost, c->s_picker));
    CU(c, cudaStreamSynchronize(c->s_picker));
    if (n_frames) *n_frames = n;
    return VX_OK;
}

static void shard_geometry(uint32_t width, uint32_t height, const VxShard* shard, uint32_t* macro_x, uint32_t* n_macros, uint32_t* owned) {
    const uint32_t tiles_x = (width + 7) / 8, tiles_y = (height + 3) / 4;
    *macro_x = (tiles_x + 3) / 4;
    *n_macros = *macro_x * ((tiles_y + 3) / 4);
    const uint32_t size = shard ? shard->world_size : 1, rank = shard ? shard->rank : 0;
    *owned = *n_macros > rank ? (*n_macros - rank + size - 1) / size : 0;
}

uint64_t vx_shard_bytes(uint32_t width, uint32_t height, const VxShard* shard) {
    if (shard && (shard->world_size == 0 || shard->rank >= shard->world_size)) return 0;
    uint32_t mx, n, owned;
    shard_geometry(width, height, shard, &mx, &n, &owned);
    return (uint64_t)owned * 512 * sizeof(float4);
}

static int shard_copy(VxCtx* c, const VxShard* shard, void* packed_dev, bool pack) {
    if (!c || !packed_dev || !c->frame_w) return fail(c, VX_E_ARG, "vx_%s_shard: null argument / nothing rendered", pack ? "pack" : "unpack");
    if (shard && (shard->world_size == 0 || shard->rank >= shard->world_size))   // (VX_SHARD_ROWS shards are not packed: their stripes are contiguous already)
        return fail(c, VX_E_ARG, "vx_%s_shard: bad shard", pack ? "pack" : "unpack");
    CU(c, cudaSetDevice(c->cfg.device));
    uint32_t mx, n, owned;
    shard_geometry(c->frame_w, c->frame_h, shard, &mx, &n, &owned);
    if (!owned) return VX_OK;
    const uint32_t size = shard ? shard->world_size : 1, rank = shard ? shard->rank : 0;
    if (pack) shard_copy_kernel<true><<<owned, 128, 0, c->s_render>>>(c->d_frame, (float4*)packed_dev, c->frame_w, c->frame_h, mx, n, rank, size);
    else shard_copy_kernel<false><<<owned, 128, 0, c->s_render>>>(c->d_frame, (float4*)packed_dev, c->frame_w, c->frame_h, mx, n, rank, size);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->e_render, c->s_render));
    return VX_OK;
}
int vx_pack_shard(VxCtx* c, const VxShard* shard, void* packed_dev) { return shard_copy(c, shard, packed_dev, true); }
int vx_unpack_shard(VxCtx* c, const VxShard* shard, const void* packed_dev) { return shard_copy(c, shard, const_cast<void*>(packed_dev), false); }

int vx_frame_ipc_handle(VxCtx* c, uint8_t handle_out[64]) {
    if (!c || !handle_out || !c->d_frame) return fail(c, VX_E_ARG, "vx_frame_ipc_handle: null argument / no framebuffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU(c, cudaSetDevice(c->cfg.device));
    cudaIpcMemHandle_t h;
    CU(c, cudaIpcGetMemHandle(&h, c->d_frame));
    std::memcpy(handle_out, &h, 64);
    return VX_OK;
}

int vx_frame8_ipc_handle(VxCtx* c, uint8_t handle_out[64]) {
    if (!c || !handle_out || !c->d_frame8) return fail(c, VX_E_ARG, "vx_frame8_ipc_handle: null argument / no framebuffer");
    CU(c, cudaSetDevice(c->cfg.device));
    cudaIpcMemHandle_t h;
    CU(c, cudaIpcGetMemHandle(&h, c->d_frame8));
    std::memcpy(handle_out, &h, 64);
    return VX_OK;
}

int vx_open_peer_frame8(VxCtx* c, const uint8_t handle[64]) {
    if (!c || !handle) return fail(c, VX_E_ARG, "vx_open_peer_frame8: null argument");
    CU(c, cudaSetDevice(c->cfg.device));
    if (c->frame8_target) return fail(c, VX_E_STATE, "vx_open_peer_frame8: a peer RGBA8 frame is already open");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    CU(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->frame8_target = (uint32_t*)p;
    return VX_OK;
}

int vx_open_peer_frame(VxCtx* c, const uint8_t handle[64]) {
    if (!c || !handle) return fail(c, VX_E_ARG, "vx_open_peer_frame: null argument");
    CU(c, cudaSetDevice(c->cfg.device));
    if (c->frame_target) return fail(c, VX_E_STATE, "vx_open_peer_frame: a peer frame is already open");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    CU(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->frame_target = (float4*)p;
    return VX_OK;
}

int vx_close_peer_frame(VxCtx* c) {
    if (!c) return VX_E_ARG;
    if (!c->frame_target && !c->frame8_target) return VX_OK;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->s_render));
    if (c->frame_target) CU(c, cudaIpcCloseMemHandle(c->frame_target));
    if (c->frame8_target) CU(c, cudaIpcCloseMemHandle(c->frame8_target));
    c->frame_target = nullptr; c->frame8_target = nullptr;
    return VX_OK;
}

// ---- chunk serialization on the GPU ---------------------------------------------------------------------
static_assert(sizeof(ChunkOut) == sizeof(VxChunkInfo), "VxChunkInfo layout");

int vx_serialize_chunks_esvo(VxCtx* c, const uint32_t* blocks, uint32_t n_chunks, const uint8_t* lods, VxChunkInfo* infos_out, void* records_out,
                             uint64_t records_capacity, uint64_t* total_bytes) {
    if (!c || !blocks || !n_chunks || !infos_out) return fail(c, VX_E_ARG, "vx_serialize_chunks_esvo: null/empty argument");
    if (n_chunks > 16384) return fail(c, VX_E_CAPACITY, "vx_serialize_chunks_esvo: at most 16384 chunks per call");
    CU(c, cudaSetDevice(c->cfg.device));
    cudaStream_t st = c->s_upload;
    if (!c->t0_chunk) { CU(c, cudaEventCreate(&c->t0_chunk)); CU(c, cudaEventCreate(&c->t1_chunk)); }
    cudaPointerAttributes pa{};
    const bool on_device = cudaPointerGetAttributes(&pa, blocks) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    const size_t in_bytes = (size_t)n_chunks * 32768 * 4;
    const uint32_t* d_in = blocks;
    if (!on_device) {
        if (c->chunk_in_cap < in_bytes) {
            if (c->d_chunk_in) cudaFree(c->d_chunk_in);
            c->d_chunk_in = nullptr; c->chunk_in_cap = 0;
            CU(c, cudaMalloc(&c->d_chunk_in, in_bytes));
            c->chunk_in_cap = in_bytes;
        }
        CU(c, cudaMemcpyAsync(c->d_chunk_in, blocks, in_bytes, cudaMemcpyHostToDevice, st));
        d_in = c->d_chunk_in;
    }
    const size_t out_bytes = (size_t)n_chunks * 4681 * 48;   // every octant of every chunk present: the worst case
    if (c->chunk_out_cap < out_bytes) {
        if (c->d_chunk_out) cudaFree(c->d_chunk_out);
        c->d_chunk_out = nullptr; c->chunk_out_cap = 0;
        CU(c, cudaMalloc(&c->d_chunk_out, out_bytes));
        c->chunk_out_cap = out_bytes;
    }
    if (c->chunk_n_cap < n_chunks) {
        if (c->d_chunk_info) cudaFree(c->d_chunk_info);
        if (c->d_chunk_lod) cudaFree(c->d_chunk_lod);
        c->d_chunk_info = nullptr; c->d_chunk_lod = nullptr; c->chunk_n_cap = 0;
        CU(c, cudaMalloc(&c->d_chunk_info, (size_t)n_chunks * sizeof(ChunkOut)));
        CU(c, cudaMalloc(&c->d_chunk_lod, n_chunks));
        c->chunk_n_cap = n_chunks;
    }
    if (lods) CU(c, cudaMemcpyAsync(c->d_chunk_lod, lods, n_chunks, cudaMemcpyHostToDevice, st));
    unsigned long long* bump = c->d_work + 6;                       // [6] bump pointer (words), [7] overflow count
    CU(c, cudaMemsetAsync(bump, 0, 16, st));
    CU(c, cudaFuncSetAttribute(serialize_chunks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChunkSmem)));
    CU(c, cudaEventRecord(c->t0_chunk, st));
    serialize_chunks_kernel<<<n_chunks, VX_CHUNK_THREADS, sizeof(ChunkSmem), st>>>(d_in, lods ? c->d_chunk_lod : nullptr, n_chunks, c->d_chunk_out,
                                                                                   (unsigned long long)(c->chunk_out_cap / 4), bump,
                                                                                   c->d_chunk_info, reinterpret_cast<unsigned int*>(bump + 1));
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->t1_chunk, st));
    unsigned long long h[2] = {0, 0};
    CU(c, cudaMemcpyAsync(h, bump, 16, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(infos_out, c->d_chunk_info, (size_t)n_chunks * sizeof(ChunkOut), cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    CU(c, cudaEventElapsedTime(&c->last_chunk_ms, c->t0_chunk, c->t1_chunk));
    if ((unsigned int)h[1]) return fail(c, VX_E_CAPACITY, "vx_serialize_chunks_esvo: output scratch overflow");
    const uint64_t total = h[0] * 4ull;
    if (total_bytes) *total_bytes = total;
    if (records_out) {
        if (total > records_capacity) return fail(c, VX_E_CAPACITY, "vx_serialize_chunks_esvo: %llu bytes of records, %llu given", (unsigned long long)total,
                                                  (unsigned long long)records_capacity);
        CU(c, cudaMemcpyAsync(records_out, c->d_chunk_out, total, cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
    }
    return VX_OK;
}

int vx_serialize_chunks_result(VxCtx* c, void** records_dev, float* kernel_ms) {
    if (!c) return VX_E_ARG;
    if (records_dev) *records_dev = c->d_chunk_out;
    if (kernel_ms) *kernel_ms = c->last_chunk_ms;
    return VX_OK;
}

int vx_svo_write_device(VxCtx* c, uint64_t range_offset, const void* src_dev, uint64_t length) {
    if (!c || !src_dev) return fail(c, VX_E_ARG, "vx_svo_write_device: null argument");
    if (range_offset + length + c->head > c->cfg.svo_capacity_bytes)
        return fail(c, VX_E_CAPACITY, "dst is not large enough: len=%llu range_start=%llu range_length=%llu", (unsigned long long)c->cfg.svo_capacity_bytes,
                    (unsigned long long)range_offset, (unsigned long long)length);
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_render, 0));
    CU(c, cudaStreamWaitEvent(c->s_upload, c->e_picker, 0));
    CU(c, cudaMemcpyAsync(c->d_world + c->head + range_offset, src_dev, length, cudaMemcpyDeviceToDevice, c->s_upload));
    { const int rb_ = refresh_bounds(c, c->stats.depth ? c->stats.depth : 23); if (rb_) return rb_; }
    CU(c, cudaEventRecord(c->e_upload, c->s_upload));
    return VX_OK;
}

int vx_sync_ipc_handle(VxCtx* c, uint8_t handle_out[64]) {
    if (!c || !handle_out) return fail(c, VX_E_ARG, "vx_sync_ipc_handle: null argument");
    CU(c, cudaSetDevice(c->cfg.device));
    cudaIpcMemHandle_t h;
    CU(c, cudaIpcGetMemHandle(&h, c->d_flags));
    std::memcpy(handle_out, &h, 64);
    return VX_OK;
}

int vx_open_peer_sync(VxCtx* c, const uint8_t handle[64]) {
    if (!c || !handle) return fail(c, VX_E_ARG, "vx_open_peer_sync: null argument");
    CU(c, cudaSetDevice(c->cfg.device));
    if (c->flags_target) return fail(c, VX_E_STATE, "vx_open_peer_sync: peer flags are already open");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    CU(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->flags_target = (unsigned int*)p;
    return VX_OK;
}

int vx_close_peer_sync(VxCtx* c) {
    if (!c) return VX_E_ARG;
    if (!c->flags_target) return VX_OK;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->s_render));
    CU(c, cudaIpcCloseMemHandle(c->flags_target));
    c->flags_target = nullptr;
    return VX_OK;
}

int vx_frame_signal(VxCtx* c, uint32_t slot, uint32_t value) {
    if (!c || slot >= 61) return fail(c, VX_E_ARG, "vx_frame_signal: bad slot");
    CU(c, cudaSetDevice(c->cfg.device));
    unsigned int* f = c->flags_target ? c->flags_target : c->d_flags;
    flag_signal_kernel<<<1, 1, 0, c->s_render>>>(f + slot, value);
    c->launches++;
    CU(c, cudaGetLastError());
    return VX_OK;
}

int vx_frame_wait(VxCtx* c, uint32_t first_slot, uint32_t n_slots, uint32_t value) {
    if (!c || n_slots == 0 || n_slots > 32 || first_slot + n_slots > 61) return fail(c, VX_E_ARG, "vx_frame_wait: bad slot range");
    CU(c, cudaSetDevice(c->cfg.device));
    unsigned int* f = c->flags_target ? c->flags_target : c->d_flags;
    flag_wait_kernel<<<1, 32, 0, c->s_render>>>(f, first_slot, n_slots, value, c->d_flags + 63);
    c->launches++;
    CU(c, cudaGetLastError());
    return VX_OK;
}

int vx_frame_gate(VxCtx* c, uint32_t slot, uint32_t value) {
    if (!c || slot >= 61) return fail(c, VX_E_ARG, "vx_frame_gate: bad slot");
    c->gate_armed = true; c->gate_slot = slot; c->gate_value = value;
    return VX_OK;
}

int vx_frame_sync_errors(VxCtx* c, uint32_t* out) {
    if (!c || !out) return VX_E_ARG;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->s_render));
    unsigned int h[3] = {0, 0, 0};   // [61] strip waits of the overlapped wavefront, [62] refused dirty ranges (vx_svo_scatter_errors), [63] frame-flag waits
    CU(c, cudaMemcpy(h, c->d_flags + 61, sizeof(h), cudaMemcpyDeviceToHost));
    *out = h[0] + h[2];
    return VX_OK;
}

int vx_svo_scatter_errors(VxCtx* c, uint32_t* out) {
    if (!c || !out) return VX_E_ARG;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->s_upload));
    CU(c, cudaMemcpy(out, c->d_flags + 62, 4, cudaMemcpyDeviceToHost));
    return VX_OK;
}

int vx_frame_flags_reset(VxCtx* c) {
    if (!c) return VX_E_ARG;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->s_render));
    CU(c, cudaMemset(c->d_flags, 0, 64 * sizeof(unsigned int)));
    c->gate_armed = false;
    return VX_OK;
}

int vx_probe_read_bandwidth(VxCtx* c, uint64_t bytes, uint32_t passes, float* gb_per_s) {
    if (!c || !gb_per_s || bytes < (1u << 20) || passes == 0) return fail(c, VX_E_ARG, "vx_probe_read_bandwidth: bad argument");
    CU(c, cudaSetDevice(c->cfg.device));
    uint4* buf = nullptr;
    uint32_t* sink = nullptr;
    const unsigned long long n16 = bytes / 16;
    CU(c, cudaMalloc(&buf, n16 * 16));
    if (cudaMalloc(&sink, 4) != cudaSuccess) { cudaFree(buf); return fail(c, VX_E_CUDA, "vx_probe_read_bandwidth: cudaMalloc failed"); }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemsetAsync(buf, 1, n16 * 16, c->s_render);
    const int grid = c->sm_count * 8;
    read_probe_kernel<<<grid, 256, 0, c->s_render>>>(buf, n16, 1, sink);          // warm: the buffer is in the L2 now if it fits
    float best = 0.0f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c->s_render);
        read_probe_kernel<<<grid, 256, 0, c->s_render>>>(buf, n16, passes, sink);
        cudaEventRecord(e1, c->s_render);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const float gbs = ms > 0.0f ? (float)((double)n16 * 16.0 * passes / (ms * 1e-3) / 1e9) : 0.0f;
        if (gbs > best) best = gbs;
    }
    c->launches += 6;
    const cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    if (e != cudaSuccess) return fail(c, VX_E_CUDA, "vx_probe_read_bandwidth: %s", cudaGetErrorString(e));
    *gb_per_s = best;
    return VX_OK;
}

int vx_set_streams(VxCtx* c, void* render_stream, void* upload_stream, void* picker_stream) {
    if (!c) return VX_E_ARG;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaDeviceSynchronize());
    if (!c->own_streams[0]) { c->own_streams[0] = c->s_render; c->own_streams[1] = c->s_upload; c->own_streams[2] = c->s_picker; }
    c->s_render = render_stream ? (cudaStream_t)render_stream : c->own_streams[0];
    c->s_upload = upload_stream ? (cudaStream_t)upload_stream : c->own_streams[1];
    c->s_picker = picker_stream ? (cudaStream_t)picker_stream : c->own_streams[2];
    // re-arm the ordering events on the new streams
    CU(c, cudaEventRecord(c->e_upload, c->s_upload));
    CU(c, cudaEventRecord(c->e_render, c->s_render));
    CU(c, cudaEventRecord(c->e_picker, c->s_picker));
    install_l2_window(c);
    return VX_OK;
}

int vx_stream(VxCtx* c, int which, void** out_stream) {
    if (!c || !out_stream || which < 0 || which > 2) return VX_E_ARG;
    *out_stream = which == 0 ? (void*)c->s_render : which == 1 ? (void*)c->s_upload : (void*)c->s_picker;
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
        if (which == 0) {
            CU(c, cudaEventElapsedTime(&st.trace_ms, c->t_wave[0], c->t_wave[1]));
            CU(c, cudaEventElapsedTime(&st.shade_ms, c->t_wave[1], c->t_wave[2]));
            CU(c, cudaEventElapsedTime(&st.shadow_ms, c->t_wave[2], c->t_wave[3]));
        }
    }
    *out = st;
    return VX_OK;
}

}  // extern "C"

#include "group.inl"
