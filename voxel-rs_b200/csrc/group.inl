// group.inl — vx_group_*: one process, several GPUs (included at the end of voxelrt.cu; needs VxCtx's internals).
//
// The reference engine is ONE process (src/gamelogic/game.rs:102-160) driving one graphics::Svo (src/graphics/svo.rs). A VxGroup is
// that Svo spread over the GPUs of a box: one VxCtx per device, every device holds a full replica of the SVO, a
// frame is cut into image-space shards, and the group's calls have the shape of the single-GPU ones:
//
//   vx_group_svo_commit          dirty ranges: pinned mirror -> ONE packed H2D to device 0 -> ncclBroadcast over NVLink to every
//                                replica's staging buffer -> scatter kernel on each (Svo::update, svo.rs:171-189)
//   vx_group_render              frame kept in GPU memory: every device traces its interleaved macro blocks and its shade / shadow
//                                kernels store the finished pixels straight into device 0's framebuffer (peer memory over NVLink);
//                                CUDA events order the devices (no flag kernels, no collective)
//   vx_group_render_read_rgba8   frame wanted on the host: every device traces whole 16-pixel stripes (VX_SHARD_ROWS) and DMAs them
//                                into the caller's host frame itself — N PCIe links in parallel instead of a gather to GPU 0
//   vx_group_raycast             contiguous slices of the task array, one per device
//
// Launches are issued by one worker thread per device (a frame is ~10 launches; issued from one thread the last of 8 devices would
// start ~0.3 ms late, as long as its whole share of a 4K frame takes). NCCL is opened at run time (dlopen "libnccl.so.2"): the
// library has no link-time dependency on it and a single-GPU user never loads it. No NCCL / no peer access => VX_E_NCCL /
// VX_E_STATE from the call that needs it; nothing falls back to the CPU.
#include <dlfcn.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace {

// the few NCCL entry points used, declared here so that nccl.h is not needed to build (ABI of NCCL 2.x: ncclResult_t and
// ncclDataType_t are ints, ncclUint8 == 1)
typedef struct ncclComm* vx_ncclComm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(vx_ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(vx_ncclComm_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, vx_ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& why) {
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { why = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define VX_NCCL_SYM(field, name)                                                             \
        field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                         \
        if (!field) { why = std::string("libnccl: no symbol ") + name; return false; }
        VX_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        VX_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        VX_NCCL_SYM(Broadcast, "ncclBroadcast")
        VX_NCCL_SYM(GroupStart, "ncclGroupStart")
        VX_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        VX_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef VX_NCCL_SYM
        return true;
    }
};

}  // namespace

struct VxGroup {
    std::vector<VxCtx*> ctx;
    std::vector<int> dev;
    uint32_t n = 0;
    std::string err;
    NcclApi nccl;
    std::vector<vx_ncclComm_t> comms;
    bool peer_frame = false;           // every device can store into device 0's memory
    cudaEvent_t e_released = nullptr;  // device 0: everything enqueued before this frame (consumers of the previous one) is done
    std::vector<cudaEvent_t> e_done;   // device i: its share of the frame is out
    uint8_t* h_frame8 = nullptr;       // group-owned pinned portable host frame (vx_group_host_frame)
    size_t h_frame8_bytes = 0;

    // one worker per device
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    uint64_t epoch = 0;
    uint32_t pending = 0;
    bool quit = false;
    std::function<int(uint32_t)> job;
    std::vector<int> rc;

    int run_all(std::function<int(uint32_t)> f) {
        if (n == 1) return f(0);
        {
            std::lock_guard<std::mutex> l(m);
            job = std::move(f);
            pending = n;
            ++epoch;
        }
        cv_go.notify_all();
        std::unique_lock<std::mutex> l(m);
        cv_done.wait(l, [&] { return pending == 0; });
        for (uint32_t i = 0; i < n; ++i)
            if (rc[i] != VX_OK) { err = "device " + std::to_string(dev[i]) + ": " + ctx[i]->err; return rc[i]; }
        return VX_OK;
    }
    void worker(uint32_t i) {
        cudaSetDevice(dev[i]);
        uint64_t seen = 0;
        for (;;) {
            std::function<int(uint32_t)> f;
            {
                std::unique_lock<std::mutex> l(m);
                cv_go.wait(l, [&] { return quit || epoch != seen; });
                if (quit) return;
                seen = epoch;
                f = job;
            }
            const int r = f(i);
            {
                std::lock_guard<std::mutex> l(m);
                rc[i] = r;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
};

static thread_local std::string g_group_create_error;
static int gfail(VxGroup* g, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (g) g->err = buf; else g_group_create_error = buf;
    return code;
}
#define GCU(g, call)                                                                                                       \
    do {                                                                                                                   \
        cudaError_t e_ = (call);                                                                                           \
        if (e_ != cudaSuccess) return gfail(g, VX_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" {

const char* vx_group_last_error(const VxGroup* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }
uint32_t vx_group_size(const VxGroup* g) { return g ? g->n : 0; }
VxCtx* vx_group_ctx(VxGroup* g, uint32_t index) { return (g && index < g->n) ? g->ctx[index] : nullptr; }

void vx_group_destroy(VxGroup* g) {
    if (!g) return;
    if (!g->workers.empty()) {
        { std::lock_guard<std::mutex> l(g->m); g->quit = true; }
        g->cv_go.notify_all();
        for (std::thread& t : g->workers) t.join();
    }
    for (uint32_t i = 0; i < g->ctx.size(); ++i) {
        if (!g->ctx[i]) continue;
        cudaSetDevice(g->dev[i]);
        cudaDeviceSynchronize();
    }
    for (uint32_t i = 0; i < g->comms.size(); ++i)
        if (g->comms[i]) g->nccl.CommDestroy(g->comms[i]);
    for (uint32_t i = 0; i < g->ctx.size(); ++i) {
        if (!g->ctx[i]) { if (i < g->e_done.size() && g->e_done[i]) cudaEventDestroy(g->e_done[i]); continue; }
        g->ctx[i]->frame_target = nullptr;   // plain peer pointers into device 0's frame, not IPC mappings: nothing to close
        g->ctx[i]->gate_event = nullptr;
        cudaSetDevice(g->dev[i]);
        if (i < g->e_done.size() && g->e_done[i]) cudaEventDestroy(g->e_done[i]);
        vx_destroy(g->ctx[i]);
    }
    if (!g->ctx.empty() && g->ctx[0]) cudaSetDevice(g->dev[0]);
    if (g->e_released) cudaEventDestroy(g->e_released);
    if (g->h_frame8) cudaFreeHost(g->h_frame8);
    cudaGetLastError();   // (a failed vx_group_create must not leave its CUDA error for the next call to trip over)
    delete g;
}

int vx_group_create(const VxConfig* cfg, const int* devices, uint32_t n_devices, VxGroup** out) {
    if (!cfg || !out || !devices || n_devices == 0 || n_devices > 64) return gfail(nullptr, VX_E_ARG, "vx_group_create: null / empty argument");
    *out = nullptr;
    for (uint32_t i = 0; i < n_devices; ++i)
        for (uint32_t j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return gfail(nullptr, VX_E_ARG, "vx_group_create: device %d listed twice", devices[i]);
    VxGroup* g = new VxGroup();
    g->n = n_devices;
    g->dev.assign(devices, devices + n_devices);
    g->ctx.assign(n_devices, nullptr);
    g->rc.assign(n_devices, VX_OK);
    g->e_done.assign(n_devices, nullptr);
    auto bail = [&](int code, const std::string& msg) { g_group_create_error = msg; vx_group_destroy(g); return code; };
    for (uint32_t i = 0; i < n_devices; ++i) {
        VxConfig ci = *cfg;
        ci.device = devices[i];
        const int rc = vx_create(&ci, &g->ctx[i]);
        if (rc != VX_OK) return bail(rc, std::string("vx_group_create: ") + vx_last_error(nullptr));
        if (cudaEventCreateWithFlags(&g->e_done[i], cudaEventDisableTiming) != cudaSuccess) return bail(VX_E_CUDA, "vx_group_create: cudaEventCreate failed");
    }
    cudaSetDevice(devices[0]);
    if (cudaEventCreateWithFlags(&g->e_released, cudaEventDisableTiming) != cudaSuccess) return bail(VX_E_CUDA, "vx_group_create: cudaEventCreate failed");
    if (n_devices > 1) {
        // peer access to device 0's memory: lets vx_group_render gather the frame by plain stores over NVLink
        g->peer_frame = true;
        for (uint32_t i = 1; i < n_devices; ++i) {
            int ok = 0;
            cudaDeviceCanAccessPeer(&ok, devices[i], devices[0]);
            if (!ok) { g->peer_frame = false; continue; }
            cudaSetDevice(devices[i]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) g->peer_frame = false;
            cudaGetLastError();
        }
        std::string why;
        if (!g->nccl.load(why)) return bail(VX_E_NCCL, "vx_group_create: " + why + " (the dirty-range broadcast of a multi-GPU group needs NCCL)");
        g->comms.assign(n_devices, nullptr);
        const int r = g->nccl.CommInitAll(g->comms.data(), (int)n_devices, devices);
        if (r != 0) return bail(VX_E_NCCL, std::string("vx_group_create: ncclCommInitAll: ") + g->nccl.GetErrorString(r));
        for (uint32_t i = 0; i < n_devices; ++i) g->workers.emplace_back(&VxGroup::worker, g, i);
    }
    *out = g;
    return VX_OK;
}

int vx_group_set_materials(VxGroup* g, const VxMaterial* materials, uint32_t count) {
    if (!g) return VX_E_ARG;
    return g->run_all([=](uint32_t i) { return vx_set_materials(g->ctx[i], materials, count); });
}

int vx_group_set_textures(VxGroup* g, const uint8_t* rgba8, uint32_t width, uint32_t height, uint32_t layers, uint32_t mip_levels) {
    if (!g) return VX_E_ARG;
    return g->run_all([=](uint32_t i) { return vx_set_textures(g->ctx[i], rgba8, width, height, layers, mip_levels); });
}

uint8_t* vx_group_svo_host_mirror(VxGroup* g) { return g ? g->ctx[0]->h_mirror : nullptr; }

uint8_t* vx_group_host_frame(VxGroup* g, uint64_t bytes) {
    if (!g || !bytes) return nullptr;
    if (g->h_frame8_bytes < bytes) {
        cudaSetDevice(g->dev[0]);
        if (g->h_frame8) cudaFreeHost(g->h_frame8);
        g->h_frame8 = nullptr; g->h_frame8_bytes = 0;
        if (cudaHostAlloc(&g->h_frame8, bytes, cudaHostAllocPortable) != cudaSuccess) { gfail(g, VX_E_CUDA, "vx_group_host_frame: cudaHostAlloc of %llu bytes failed", (unsigned long long)bytes); return nullptr; }
        g->h_frame8_bytes = bytes;
    }
    return g->h_frame8;
}

int vx_group_svo_commit(VxGroup* g, float octree_scale, const VxRange* dirty, uint32_t n_dirty, uint64_t used_bytes, uint32_t depth) {
    if (!g || (n_dirty && !dirty)) return gfail(g, VX_E_ARG, "vx_group_svo_commit: null argument");
    VxCtx* c0 = g->ctx[0];
    if (g->n == 1) {
        const int rc = vx_svo_commit(c0, octree_scale, dirty, n_dirty, used_bytes, depth);
        if (rc) g->err = c0->err;
        return rc;
    }
    const uint64_t cap = c0->cfg.svo_capacity_bytes;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_dirty; ++i) {
        if (dirty[i].offset + dirty[i].length + c0->head > cap)   // the reference's assert!, esvo.rs:328-331
            return gfail(g, VX_E_CAPACITY, "dst is not large enough: len=%llu range_start=%llu range_length=%llu", (unsigned long long)cap,
                         (unsigned long long)dirty[i].offset, (unsigned long long)dirty[i].length);
        total += dirty[i].length;
    }
    std::memcpy(c0->h_mirror, &octree_scale, 4);                                  // svo.rs:173-175
    const size_t hdr_bytes = (size_t)n_dirty * sizeof(VxRange);
    const size_t bytes = hdr_bytes + c0->head + total;
    if (n_dirty && bytes > c0->stage_cap) {
        // bulk (re)load, e.g. the first commit of a world: straight from the (portable, pinned) mirror to every replica
        const int rc = g->run_all([=](uint32_t i) { return vx_svo_commit_from(g->ctx[i], c0->h_mirror, octree_scale, dirty, n_dirty, used_bytes, depth); });
        return rc;
    }
    if (n_dirty) {
        GCU(g, cudaSetDevice(g->dev[0]));
        GCU(g, cudaStreamSynchronize(c0->s_upload));                              // the staging block is reused (normally long drained)
        std::memcpy(c0->h_stage, dirty, hdr_bytes);
        size_t off = hdr_bytes;
        std::memcpy(c0->h_stage + off, c0->h_mirror, c0->head);
        off += c0->head;
        for (uint32_t i = 0; i < n_dirty; ++i) {
            std::memcpy(c0->h_stage + off, c0->h_mirror + c0->head + dirty[i].offset, dirty[i].length);
            off += dirty[i].length;
        }
    }
    // every device's share — wait for its rays in flight, (device 0: the H2D copy,) its end of the broadcast, scatter, box, event —
    // is issued by that device's worker thread: eight devices' launches in parallel instead of one after the other
    const uint64_t hot_off = c0->hot_off, hot_len = c0->hot_len;
    const int rc = g->run_all([=](uint32_t i) -> int {
        VxCtx* c = g->ctx[i];
        CU(c, cudaSetDevice(g->dev[i]));
        if (n_dirty) {
            if (!c->d_stage) CU(c, cudaMalloc(&c->d_stage, c->stage_cap));
            CU(c, cudaStreamWaitEvent(c->s_upload, c->e_render, 0));              // do not tear a frame / ray batch in flight
            CU(c, cudaStreamWaitEvent(c->s_upload, c->e_picker, 0));
            if (i == 0) CU(c, cudaMemcpyAsync(c->d_stage, c->h_stage, bytes, cudaMemcpyHostToDevice, c->s_upload));
            const int r = g->nccl.Broadcast(c->d_stage, c->d_stage, bytes, /*ncclUint8*/ 1, 0, g->comms[i], c->s_upload);
            if (r != 0) return fail(c, VX_E_NCCL, "ncclBroadcast: %s", g->nccl.GetErrorString(r));
            const unsigned long long pb = c->head + total;
            const int blocks = (int)((pb / 4 + 255) / 256 < 4096 ? (pb / 4 + 255) / 256 : 4096);
            scatter_ranges_kernel<<<blocks > 0 ? blocks : 1, 256, 0, c->s_upload>>>(c->d_world, c->d_stage, n_dirty, pb, (uint32_t)c->head,
                                                                                     c->cfg.svo_capacity_bytes, c->d_flags + 62);
            c->launches++;
            CU(c, cudaGetLastError());
            c->have_svo = true;
            const int rb = refresh_bounds(c, depth);
            if (rb) return rb;
        }
        CU(c, cudaEventRecord(c->e_upload, c->s_upload));
        c->stats.used_bytes = used_bytes; c->stats.depth = depth;
        c->hot_off = hot_off; c->hot_len = hot_len;
        install_l2_window(c);
        return VX_OK;
    });
    if (rc == VX_E_NCCL) return gfail(g, VX_E_NCCL, "vx_group_svo_commit: %s", g->err.c_str());
    return rc;
}

int vx_group_svo_set_hot_range(VxGroup* g, uint64_t offset, uint64_t length) {
    if (!g) return VX_E_ARG;
    for (VxCtx* c : g->ctx) { c->hot_off = offset; c->hot_len = length; }
    return VX_OK;
}

int vx_group_stats(const VxGroup* g, VxStats* out) { return g ? vx_stats(g->ctx[0], out) : VX_E_ARG; }

int vx_group_render(VxGroup* g, const VxRenderParams* p, uint32_t width, uint32_t height) {
    if (!g || !p) return gfail(g, VX_E_ARG, "vx_group_render: null argument");
    VxCtx* c0 = g->ctx[0];
    if (g->n == 1) {
        const int rc = vx_render(c0, p, width, height, nullptr, nullptr);
        if (rc) g->err = c0->err;
        return rc;
    }
    if (!g->peer_frame) return gfail(g, VX_E_STATE, "vx_group_render: no peer access to device %d's memory (use vx_group_render_read_rgba8)", g->dev[0]);
    GCU(g, cudaSetDevice(g->dev[0]));
    // whatever was enqueued on device 0's render stream so far reads the PREVIOUS frame (vx_read_frame_*, a consumer's kernels on
    // that stream): the other devices may trace their primary rays at once but hold their pixels back until it is done
    GCU(g, cudaEventRecord(g->e_released, c0->s_render));
    int rc = g->run_all([=](uint32_t i) {
        VxCtx* c = g->ctx[i];
        if (i) { c->frame_target = c0->d_frame; c->gate_event = g->e_released; }
        const VxShard sh{i, g->n};
        const int r = vx_render(c, p, width, height, &sh, nullptr);
        if (r == VX_OK && i && cudaEventRecord(g->e_done[i], c->s_render) != cudaSuccess) return fail(c, VX_E_CUDA, "vx_group_render: cudaEventRecord failed");
        return r;
    });
    if (rc) return rc;
    GCU(g, cudaSetDevice(g->dev[0]));
    for (uint32_t i = 1; i < g->n; ++i) GCU(g, cudaStreamWaitEvent(c0->s_render, g->e_done[i], 0));   // the frame is whole for whatever follows on this stream
    GCU(g, cudaEventRecord(c0->e_render, c0->s_render));
    return VX_OK;
}

int vx_group_render_read_rgba8(VxGroup* g, const VxRenderParams* p, uint32_t width, uint32_t height, uint8_t* rgba8_out, uint32_t bands) {
    if (!g || !p || !rgba8_out) return gfail(g, VX_E_ARG, "vx_group_render_read_rgba8: null argument");
    return g->run_all([=](uint32_t i) {
        VxCtx* c = g->ctx[i];
        c->frame_target = nullptr; c->gate_event = nullptr;   // pixels are stored locally and leave over this device's own PCIe link
        const VxShard sh{i, g->n | VX_SHARD_ROWS};
        return vx_render_read_rgba8(c, p, width, height, g->n > 1 ? &sh : nullptr, rgba8_out, bands);
    });
}

int vx_group_wait(VxGroup* g) {
    if (!g) return VX_E_ARG;
    return g->run_all([=](uint32_t i) { return vx_render_wait(g->ctx[i]); });
}

int vx_group_read_frame_rgba8(VxGroup* g, uint8_t* out) { return g ? vx_read_frame_rgba8(g->ctx[0], out) : VX_E_ARG; }
int vx_group_read_frame_rgba32f(VxGroup* g, float* out) { return g ? vx_read_frame_rgba32f(g->ctx[0], out) : VX_E_ARG; }

int vx_group_raycast(VxGroup* g, const VxPickerTask* tasks, uint64_t n, VxPickerResult* results) {
    if (!g || (n && (!tasks || !results))) return gfail(g, VX_E_ARG, "vx_group_raycast: null argument");
    if (n == 0) return VX_OK;
    // contiguous slices (multiples of a 128-ray work unit); a batch of a few rays goes to device 0 alone
    uint64_t per = (n + g->n - 1) / g->n;
    per = (per + 127) / 128 * 128;
    if (n <= 4096) per = n;
    return g->run_all([=](uint32_t i) {
        const uint64_t off = (uint64_t)i * per;
        if (off >= n) return (int)VX_OK;
        const uint64_t cnt = n - off < per ? n - off : per;
        return vx_raycast(g->ctx[i], tasks + off, cnt, results + off);
    });
}

int vx_group_set_option(VxGroup* g, uint32_t option, uint64_t value) {
    if (!g) return VX_E_ARG;
    for (VxCtx* c : g->ctx) { const int rc = vx_set_option(c, option, value); if (rc) { g->err = c->err; return rc; } }
    return VX_OK;
}

}  // extern "C"
