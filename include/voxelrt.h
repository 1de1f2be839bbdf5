/*
 * voxelrt.h — C ABI of libvoxelrt, the B200 (sm_100a) replacement for voxel-rs's
 * GLSL ray-cast path. Everything below is plain C: opaque context, POD structs,
 * raw pointers and sizes. No C++/torch types cross this boundary.
 *
 * Each entry point names the reference interface it replaces; paths are relative
 * to the voxel-rs checkout (tim-oster/voxel-rs).
 *
 *   reference seam                                   replaced by
 *   -----------------------------------------------  ---------------------------
 *   graphics::Svo::new          src/graphics/svo.rs:109-149   vx_create + vx_set_textures + vx_set_materials
 *   graphics::Svo::update       src/graphics/svo.rs:171-189   vx_svo_host_mirror + vx_svo_commit
 *   graphics::Svo::get_stats    src/graphics/svo.rs:191-193   vx_stats
 *   graphics::Svo::render       src/graphics/svo.rs:196-229   vx_render (+ vx_read_frame*)
 *   graphics::Svo::raycast      src/graphics/svo.rs:233-255   vx_raycast
 *   svo.test.glsl debug harness assets/shaders/svo.test.glsl  vx_debug_cast
 *   Framebuffer::read_pixels    src/graphics/framebuffer.rs:97-105  vx_read_frame_rgba8
 *
 * Threading: calls on one VxCtx must be serialised by the caller (the reference
 * issues all GL calls from the main thread, README.md:116-119). Internally the
 * context owns three CUDA streams (render, upload, picker) ordered by events.
 *
 * Errors: the reference panics (unwrap / assert!, e.g. src/world/hds/esvo.rs:328-331).
 * The FFI never unwinds: every call returns VX_OK or a negative VX_E_* code and
 * vx_last_error() holds a human-readable message for the last failure on that ctx.
 */
#ifndef VOXELRT_H
#define VOXELRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX_OK            0
#define VX_E_ARG        -1   /* bad argument (null pointer, size out of range) */
#define VX_E_CAPACITY   -2   /* dirty range / frame / ray batch exceeds what vx_create reserved
                                (reference: assert! in esvo.rs:328-331) */
#define VX_E_CUDA       -3   /* CUDA runtime error; message has cudaGetErrorString */
#define VX_E_NCCL       -4   /* NCCL error */
#define VX_E_STATE      -5   /* call out of order (e.g. render before textures/materials/SVO) */

typedef struct VxCtx VxCtx;

/* ------------------------------------------------------------------ config -- */

/* flags for VxConfig.flags */
#define VX_FLAG_NO_L2_WINDOW   (1u << 0)  /* do not install the persisting-L2 access-policy window */
#define VX_FLAG_SVO_CSVO       (1u << 2)  /* the world buffer holds the CSVO format (world::hds::csvo, the reference's default
                                             feature `use-csvo`, Cargo.toml:39-45) instead of ESVO: = the SvoType argument of
                                             graphics::Svo::new (svo.rs:109). Buffer = f32 2^-depth, u32 root offset, bytes
                                             (svo.csvo.glsl:1-5); dirty ranges live at byte 8 + offset */

typedef struct VxConfig {
    int32_t  device;              /* CUDA device ordinal this context owns */
    uint32_t flags;
    uint64_t svo_capacity_bytes;  /* = size_mb * 1000 * 1000 of graphics::Svo::new (svo.rs:133) */
    uint32_t max_width;           /* largest framebuffer this ctx will render (0 => no frame buffer) */
    uint32_t max_height;
    uint64_t max_rays;            /* largest vx_raycast batch (reference caps at 100, svo_picker.rs:5) */
} VxConfig;

/* --------------------------------------------------------------- materials -- */

/* = #[repr(C)] MaterialInstance, src/graphics/svo_registry.rs:29-40
 * = GLSL struct Material, assets/shaders/svo.glsl:48-59 (std430, stride 32). -1 = no texture. */
typedef struct VxMaterial {
    float   specular_pow;
    float   specular_strength;
    int32_t tex_top;
    int32_t tex_side;
    int32_t tex_bottom;
    int32_t tex_top_normal;
    int32_t tex_side_normal;
    int32_t tex_bottom_normal;
} VxMaterial;

/* ------------------------------------------------------------------ picker -- */

/* = #[repr(C)] PickerTask, src/graphics/svo_picker.rs:13-19 (AlignedPoint3/AlignedVec3 are
 * 16-byte aligned, src/graphics/macros.rs:64-90) = GLSL PickerTask, picker.glsl:19-23. 48 bytes. */
typedef struct VxPickerTask {
    float max_dst;  float _pad0[3];
    float pos[3];   float _pad1;
    float dir[3];   float _pad2;
} VxPickerTask;

/* = #[repr(C)] PickerResult, src/graphics/svo_picker.rs:21-32 = GLSL PickerResult,
 * picker.glsl:9-14. inside_voxel is a GLSL bool (4 bytes, 0/1); Rust reads its first byte. */
typedef struct VxPickerResult {
    float    dst;
    uint32_t inside_voxel;
    float    _pad0[2];
    float    pos[3];    float _pad1;
    float    normal[3]; float _pad2;
} VxPickerResult;

/* ------------------------------------------------------------------ render -- */

/* The ten uniforms of world.glsl:12-25 as set by graphics::Svo::render (svo.rs:201-215).
 * `view` is the ALREADY INVERTED look_to_rh matrix (camera -> world), column-major, exactly the
 * value the reference uploads as u_view (svo.rs:197,204) so device code never re-derives it.
 * highlight_pos = NaN,NaN,NaN when no voxel is selected (svo.rs:211-215). */
typedef struct VxRenderParams {
    float    view[16];
    float    fov_y_rad;
    float    aspect_ratio;
    float    ambient_intensity;
    float    light_dir[3];
    float    cam_pos[3];
    float    highlight_pos[3];
    uint32_t render_shadows;
    float    shadow_distance;
} VxRenderParams;

/* Image-space shard for multi-GPU rendering. The frame is cut into macro blocks of 32x16 pixels, row-major; this ctx renders
 * only the macro blocks m with m % world_size == rank (interleaved blocks: the finest mix of cheap and expensive image regions,
 * for frames that stay in GPU memory), or — with VX_SHARD_ROWS or'ed into world_size — the macro ROWS r with r % world_size ==
 * rank (whole 16-pixel-high stripes: a shard's pixels are contiguous runs of the frame, so vx_render_read_rgba8 can DMA them
 * into a host frame shared by all shards, each GPU over its own PCIe link). {0,1} renders the whole frame. */
#define VX_SHARD_ROWS 0x80000000u
typedef struct VxShard {
    uint32_t rank;
    uint32_t world_size;   /* | VX_SHARD_ROWS */
} VxShard;

/* Per-frame counters of the last vx_render / vx_raycast (device-accumulated). */
typedef struct VxFrameStats {
    uint64_t primary_rays;   /* rays traced from the camera (== pixels of this shard) */
    uint64_t shadow_rays;    /* secondary rays actually cast (world.glsl:80-84) */
    uint64_t steps;          /* traversal loop iterations (svo.esvo.glsl:152), both ray kinds */
    uint64_t pushes;         /* PUSH phases (svo.esvo.glsl:281-311) */
    uint64_t leaf_tests;     /* HIT blocks entered (svo.esvo.glsl:185-265) */
    uint64_t tex_fetches;    /* texels read (1 per NEAREST sample, 8 per trilinear) */
    float    kernel_ms;      /* CUDA-event time of the frame's kernels (trace + shade + shadow) / the picker kernel */
    float    trace_ms;       /* vx_render only: primary-ray trace kernel */
    float    shade_ms;       /* vx_render only: shading kernel */
    float    shadow_ms;      /* vx_render only: shadow-ray trace kernel */
} VxFrameStats;

/* = graphics::svo::Stats, src/graphics/svo.rs:75-83 */
typedef struct VxStats {
    uint64_t used_bytes;
    uint64_t capacity_bytes;
    uint32_t depth;
} VxStats;

/* A dirty byte range of the serialized SVO, relative to the RangeBuffer start
 * (= world::hds::internal::Range, src/world/hds/internal.rs:150-154). In the GPU buffer the
 * range lives at byte 24 + offset (4-byte octree_scale + 20-byte preamble, svo.esvo.glsl:3-6,
 * esvo.rs:134,179-188). */
typedef struct VxRange {
    uint64_t offset;
    uint64_t length;
} VxRange;

/* ---------------------------------------------------------- debug ray cast -- */

/* = StackFrame, assets/shaders/svo.test.glsl:23-33 / src/graphics/svo_shader_tests.rs:51-63 */
typedef struct VxDebugFrame {
    float    t_min;
    uint32_t ptr;
    uint32_t idx;
    uint32_t parent_octant_idx;
    int32_t  scale;
    int32_t  is_child;
    int32_t  is_leaf;
    int32_t  crossed_boundary;   /* CSVO only; always 0 for ESVO */
    uint32_t next_ptr;           /* CSVO only; always 0 for ESVO */
} VxDebugFrame;

/* = OctreeResult, assets/shaders/svo.glsl:31-40 (plus lod, which svo.test.glsl drops) */
typedef struct VxOctreeResult {
    float    t;
    uint32_t value;
    int32_t  face_id;
    float    pos[3];
    float    uv[2];
    float    color[4];
    float    lod;
    uint32_t inside_voxel;
} VxOctreeResult;

/* ------------------------------------------------------------- entry points -- */

/* graphics::Svo::new (svo.rs:109-149): allocate device world buffer, pinned host mirror,
 * framebuffer, picker buffers, streams. */
int vx_create(const VxConfig* cfg, VxCtx** out);
void vx_destroy(VxCtx* ctx);
const char* vx_last_error(const VxCtx* ctx);   /* ctx may be NULL: returns the last create error */

/* VoxelRegistry::build_material_buffer (svo_registry.rs:135-165): table indexed by BlockId. */
int vx_set_materials(VxCtx* ctx, const VxMaterial* materials, uint32_t count);

/* VoxelRegistry::build_texture_array / TextureArrayBuilder::build (svo_registry.rs:122-133,
 * texture_array.rs:83-153): `rgba8` holds `layers` images of width*height RGBA8 texels, level 0
 * only, ALREADY vertically flipped the way the reference flips on load (texture_array.rs:92,126,
 * 155-176). mip_levels is clamped to min(mip_levels, ilog2(min(w,h))) like texture_array.rs:105;
 * the library generates the chain on the GPU (2x2 box filter, round-to-nearest — the
 * glGenerateMipmap stand-in, texture_array.rs:258-260). Sampler state is fixed to the
 * reference's: S clamp, T repeat, MIN linear-mipmap-linear, MAG nearest (texture_array.rs:200-203). */
int vx_set_textures(VxCtx* ctx, const uint8_t* rgba8, uint32_t width, uint32_t height,
                    uint32_t layers, uint32_t mip_levels);

/* Pinned host mirror of the whole GPU world buffer (capacity = svo_capacity_bytes). Byte 0 is
 * the f32 octree_scale; the Rust side passes `mirror + 4` as `dst` to the unchanged
 * Esvo::write_changes_to (svo.rs:180-181, esvo.rs:310-339). */
uint8_t* vx_svo_host_mirror(VxCtx* ctx);

/* graphics::Svo::update (svo.rs:171-189) after write_changes_to filled the mirror: waits for the
 * frame in flight (the render_fence of svo.rs:178), then copies byte 0..24 (scale + preamble) and
 * every dirty range (offset relative to RangeBuffer start) host->device with cudaMemcpyAsync on
 * the upload stream; the next render/raycast waits on the upload event. `octree_scale` is written
 * to mirror byte 0 by this call (svo.rs:173-175). n_dirty==0 with used_bytes>0 uploads nothing but
 * still refreshes stats. */
int vx_svo_commit(VxCtx* ctx, float octree_scale, const VxRange* dirty, uint32_t n_dirty,
                  uint64_t used_bytes, uint32_t depth);

/* Optional hint: RangeBuffer byte range of the world-root octree (octant_to_range[u64::MAX],
 * esvo.rs:270) — pinned in L2 with a persisting access-policy window together with the preamble. */
int vx_svo_set_hot_range(VxCtx* ctx, uint64_t offset, uint64_t length);

/* Multi-GPU replicas: same as vx_svo_commit but the packed dirty payload comes from a device
 * buffer that was just broadcast (NCCL) from rank 0: `packed` = n_dirty VxRange headers followed
 * by the concatenated range bytes, first 24 bytes of the world buffer in `head24`. Applied by a
 * scatter kernel on the upload stream. */
int vx_svo_commit_packed_device(VxCtx* ctx, const void* packed_dev, uint32_t n_dirty,
                                uint64_t payload_bytes, uint64_t used_bytes, uint32_t depth);
/* Rank-0 helper: pack mirror ranges into `out` (host, pinned or not) in the layout above.
 * Returns bytes written or a negative error; call with out==NULL to query the size. */
int64_t vx_svo_pack_dirty(VxCtx* ctx, const VxRange* dirty, uint32_t n_dirty, void* out, uint64_t out_cap);
/* Ranges of packed dirty sets that the scatter kernel refused so far because offset + length would leave the buffer (the headers
 * of vx_svo_commit_packed_device live in device memory, so they are checked there; a refused range is skipped, never written out
 * of bounds). Waits for the upload stream. 0 in a healthy run. */
int vx_svo_scatter_errors(VxCtx* ctx, uint32_t* out);

/* Measurement aid (no counterpart in the reference): read bandwidth of this GPU in GB/s, best of 5 — every thread streams 16-byte
 * words of a scratch buffer of `bytes` bytes `passes` times with L2-only caching. bytes well below the L2 size (e.g. 32 MiB)
 * measures the L2, bytes several times the L2 (e.g. 2 GiB, passes 1) measures HBM: the denominators bench.py reports the ray
 * caster's node-fetch traffic against ("report against both"). */
int vx_probe_read_bandwidth(VxCtx* ctx, uint64_t bytes, uint32_t passes, float* gb_per_s);

int vx_stats(const VxCtx* ctx, VxStats* out);

/* graphics::Svo::render (svo.rs:196-229) = world.glsl main(): one primary ray per pixel, shading,
 * optional shadow ray, sky; writes the context's RGBA32F device framebuffer (row 0 = bottom row,
 * world.glsl:112-115,140). Asynchronous on the render stream unless rgba32f_out != NULL, in which
 * case the frame is copied to that HOST buffer (width*height*16 bytes) and the call returns after
 * the copy completed. shard may be NULL (= whole frame). */
int vx_render(VxCtx* ctx, const VxRenderParams* params, uint32_t width, uint32_t height,
              const VxShard* shard, float* rgba32f_out);

/* graphics::Svo::render + Framebuffer::read_pixels (svo.rs:196-229 + framebuffer.rs:97-105) in one pipelined call: the frame is
 * rendered in `bands` bands of macro-block rows; each finished band is converted to RGBA8 and copied to the HOST buffer
 * rgba8_out (width*height*4 bytes, row 0 = bottom; pinned memory for the copy to overlap) while the next band is traced.
 * Returns when the whole frame is in rgba8_out. bands is clamped to 1..16. The kernels store the RGBA8 pixels directly
 * (bit-identical to vx_render + vx_read_frame_rgba8); the RGBA32F frame is NOT produced by this call. */
int vx_render_read_rgba8(VxCtx* ctx, const VxRenderParams* params, uint32_t width, uint32_t height,
                         const VxShard* shard, uint8_t* rgba8_out, uint32_t bands);
/* The same in two halves, for a caller that prepares the next frame's inputs while this one renders: _begin returns as soon as
 * the frame and its band copies are enqueued, _end returns when the OLDEST frame begun and not yet ended (this shard's stripes of
 * it) is in its rgba8_out. Up to TWO frames may be in flight (the PBO-style double-buffered read-back of a render loop):
 * _begin(k+1) before _end(k) renders frame k+1 into a second device frame while frame k's copy still runs, so a frame's read-back
 * hides entirely under the next frame's tracing; each frame needs its own host buffer until its _end returned. A third _begin is
 * VX_E_STATE. vx_render_read_rgba8 waits for everything in flight. */
int vx_render_read_rgba8_begin(VxCtx* ctx, const VxRenderParams* params, uint32_t width, uint32_t height,
                               const VxShard* shard, uint8_t* rgba8_out, uint32_t bands);
int vx_render_read_rgba8_end(VxCtx* ctx);

/* Block until the last vx_render finished (the reference's render_fence.wait(), svo.rs:178). */
int vx_render_wait(VxCtx* ctx);

/* Framebuffer::read_pixels (framebuffer.rs:97-105): RGBA8 = round(clamp(c,0,1)*255), row 0 = bottom. */
int vx_read_frame_rgba8(VxCtx* ctx, uint8_t* rgba8_out);
int vx_read_frame_rgba32f(VxCtx* ctx, float* rgba32f_out);
/* Device pointer of the RGBA32F framebuffer of the last render (for zero-copy consumers /
 * NCCL tile gather). Valid until vx_destroy. */
int vx_frame_device_ptr(VxCtx* ctx, void** out_ptr, uint32_t* width, uint32_t* height);

/* Parity instrument: what intersect_octree (svo.esvo.glsl:50-393) returned for the PRIMARY ray of every pixel of the last
 * vx_render — the OctreeResult fields world.glsl reads (:131-137), before any shading. out = width * height records, row 0 =
 * bottom, like the frame. A miss has t = -1 and zeros. Pixels of macro blocks another shard owns are not filled (t = -2).
 * The reference keeps these values in shader registers; here they are the hit records the wavefront hands from
 * trace_primary_kernel to shade_kernel, so the bit-exactness bar of the hit voxel / material / face / distance can be tested
 * directly instead of through colours. */
typedef struct VxHitRecord {
    float    t;          /* res.t (voxel units), -1 = miss */
    uint32_t value;      /* res.value: block id */
    int32_t  face_id;    /* res.face_id 0..5 */
    float    pos[3];     /* res.pos */
    float    uv[2];      /* res.uv */
} VxHitRecord;
int vx_read_hit_records(VxCtx* ctx, VxHitRecord* out);

/* graphics::Svo::raycast (svo.rs:233-255) = picker.glsl main(): tasks/results are HOST arrays of n
 * records; synchronous like the reference (fence place+wait, svo.rs:248-249). n is not capped at
 * 100 (svo_picker.rs:5) — only by VxConfig.max_rays. Batches above 1 Mi rays are traced in slices so that
 * the upload of the tasks, the tracing and the read-back of the results overlap (give it pinned host
 * memory for that); the call still returns only when every result is in `results`. */
int vx_raycast(VxCtx* ctx, const VxPickerTask* tasks, uint64_t n, VxPickerResult* results);

/* Same kernel on DEVICE-resident task/result arrays (no copies, asynchronous on the picker
 * stream; vx_raycast_wait to join). Used by benchmarks to time the kernel with inputs in HBM. */
int vx_raycast_device(VxCtx* ctx, const VxPickerTask* tasks_dev, uint64_t n, VxPickerResult* results_dev);
int vx_raycast_wait(VxCtx* ctx);

/* svo.test.glsl main(): cast one ray and record every loop iteration (svo.esvo.glsl:175).
 * frames_cap frames at most are stored; *n_frames receives the total count (stack_ptr + 1). */
int vx_debug_cast(VxCtx* ctx, const float pos[3], const float dir[3], float max_dst,
                  uint32_t cast_translucent, VxOctreeResult* result,
                  VxDebugFrame* frames, uint32_t frames_cap, uint32_t* n_frames);

/* ---- chunk serialization on the GPU (reference: SerializedChunk::new + serialize_octant,
 * src/world/hds/esvo.rs:353-383, 439-512, run by the job system in src/systems/worldsvo.rs:90-99) ----
 * blocks: n_chunks dense 32^3 BlockId arrays (index x + 32*(y + 32*z), 0 = air), HOST or DEVICE memory. lods: one byte per chunk
 * (0 = full detail), host memory, may be NULL. The records of chunk i (12-word octant records, depth-first, byte-identical to the
 * host serializer) land at infos_out[i].offset_bytes of the context's scratch buffer; chunks are laid out in completion order.
 * child_mask / leaf_mask / depth = the SerializationResult of the chunk's root octant (what Esvo::serialize_root needs,
 * esvo.rs:151-175). records_out (host, optional) receives the *total_bytes of records. */
typedef struct VxChunkInfo {
    uint64_t offset_bytes;
    uint64_t length_bytes;
    uint8_t  child_mask, leaf_mask, depth, _pad[5];
} VxChunkInfo;
int vx_serialize_chunks_esvo(VxCtx* ctx, const uint32_t* blocks, uint32_t n_chunks, const uint8_t* lods, VxChunkInfo* infos_out,
                             void* records_out, uint64_t records_capacity, uint64_t* total_bytes);
/* Device pointer of the records of the last vx_serialize_chunks_esvo and its kernel time (CUDA events). */
int vx_serialize_chunks_result(VxCtx* ctx, void** records_dev, float* kernel_ms);
/* Copies `length` bytes of device memory (records built on the GPU) into the world buffer at RangeBuffer offset `range_offset`,
 * ordered like vx_svo_commit (after the frame in flight, before the next one): dirty chunks without a host round trip. */
int vx_svo_write_device(VxCtx* ctx, uint64_t range_offset, const void* src_dev, uint64_t length);

/* ---- multi-GPU plumbing (no reference counterpart: the reference is single-GPU) ----
 * One process per GPU. Each rank renders its VxShard into its own full-size framebuffer; the shard's
 * pixels are then packed into a contiguous device buffer ([owned macro block][16 rows][32 px] RGBA32F),
 * moved to rank 0 by the caller's collective (NCCL over NVLink) and unpacked into rank 0's framebuffer. */
uint64_t vx_shard_bytes(uint32_t width, uint32_t height, const VxShard* shard);   /* packed size of one shard */
int vx_pack_shard(VxCtx* ctx, const VxShard* shard, void* packed_dev);            /* frame -> packed, on the render stream */
int vx_unpack_shard(VxCtx* ctx, const VxShard* shard, const void* packed_dev);    /* packed -> frame, on the render stream */

/* Fused tile gather over NVLink peer memory: a non-root rank maps the root GPU's framebuffer (CUDA IPC) and installs it as
 * its frame target; its shade / shadow kernels then store every finished pixel of the shard straight into the root's frame
 * (write-only 16-byte stores, no pack / send / recv / unpack), overlapping the transfer with tracing. The root learns that
 * the frame is complete from the caller's barrier on the render streams.
 *   vx_frame_ipc_handle   root: 64-byte cudaIpcMemHandle_t of this ctx's framebuffer
 *   vx_open_peer_frame    other ranks: map that handle and write finished pixels there from now on
 *   vx_close_peer_frame   back to the local framebuffer, unmap */
int vx_frame_ipc_handle(VxCtx* ctx, uint8_t handle_out[64]);
int vx_open_peer_frame(VxCtx* ctx, const uint8_t handle[64]);
int vx_close_peer_frame(VxCtx* ctx);   /* closes both the RGBA32F and the RGBA8 peer frame */
/* The same for the RGBA8 frame, used with vx_set_option(8, 1) on every rank: pixels then cross NVLink as 4 bytes instead of
 * 16 (what vx_read_frame_rgba8 hands out anyway); vx_read_frame_rgba32f is not available in that mode. */
int vx_frame8_ipc_handle(VxCtx* ctx, uint8_t handle_out[64]);
int vx_open_peer_frame8(VxCtx* ctx, const uint8_t handle[64]);

/* Frame flags for the peer-memory gather: 63 32-bit frame counters per context; the root's are mapped by its peers (same IPC
 * scheme as the framebuffer), so ranks order their frames without a collective:
 *   vx_frame_signal(slot, v)     on the render stream: publish everything this GPU wrote so far (its pixels in the root's
 *                                frame), then raise flag `slot` (of the root if peer flags are open, else local) to v
 *   vx_frame_wait(first, n, v)   on the render stream: wait until flags [first, first+n) are all >= v
 *   vx_frame_gate(slot, v)       the next vx_render waits for flag `slot` >= v BETWEEN its trace and shade kernels: primary
 *                                rays of frame i+1 overlap the root's consumption of frame i, pixels are held back
 *   vx_frame_sync_errors         number of waits that gave up (~2 s): a lost peer is an error, not a hang */
int vx_sync_ipc_handle(VxCtx* ctx, uint8_t handle_out[64]);
int vx_open_peer_sync(VxCtx* ctx, const uint8_t handle[64]);
int vx_close_peer_sync(VxCtx* ctx);
int vx_frame_signal(VxCtx* ctx, uint32_t slot, uint32_t value);
int vx_frame_wait(VxCtx* ctx, uint32_t first_slot, uint32_t n_slots, uint32_t value);
int vx_frame_gate(VxCtx* ctx, uint32_t slot, uint32_t value);
int vx_frame_sync_errors(VxCtx* ctx, uint32_t* out);
/* Zeroes this context's frame flags (and error counters) after waiting for its render stream. The root calls it — followed by a
 * barrier among the ranks — before a new sequence of frames starts counting from 1 again. */
int vx_frame_flags_reset(VxCtx* ctx);

/* Run this context's work on caller-owned CUDA streams (cudaStream_t passed as void*) so that it orders
 * with the caller's collectives without host synchronisation. NULL keeps the library's own stream. */
int vx_set_streams(VxCtx* ctx, void* render_stream, void* upload_stream, void* picker_stream);
/* The cudaStream_t a kind of work is issued on: 0 = render, 1 = upload, 2 = picker. */
int vx_stream(VxCtx* ctx, int which, void** out_stream);

/* Counters + kernel time of the last vx_render (which=0) or vx_raycast* (which=1). Implies a wait. */
int vx_frame_stats(VxCtx* ctx, int which, VxFrameStats* out);

/* Runtime knobs for A/B measurements (no reference counterpart):
 *   3 = count steps/pushes/leaf tests/texels (0/1)     4 = CTAs per SM of the persistent trace kernels (0 = default 8;
 *                                                          <=5 / 6-7 / >=8 select the 96 / 80 / 64-register builds)
 *   5 = persisting-L2 access-policy window (0/1, default 1)
 *   6 = refill threshold of the render trace kernels: a warp leaves its walk loop to finish/refill rays when fewer
 *       than this many lanes are still walking (1..32; default 1 = run all rays of the warp to their end, then
 *       refill all 32 lanes — measured fastest on coherent frames, profiles/r01_v1_*)
 *   8 = RGBA8 output (0/1, default 0): the shade / shadow kernels store round(clamp(c)*255) pixels into the RGBA8 frame
 *       instead of the RGBA32F one — bit-identical to rendering RGBA32F and converting (framebuffer.rs:97-105), a quarter of
 *       the frame bytes (multi-GPU gather, read-back)
 *   9 = TMA write-back of framebuffer tiles in the shade kernel (0/1, default 0: measured 3 % slower): a strip's 4 x 32 RGBA32F pixels are staged in
 *       shared memory and written with four 512-byte bulk async copies instead of 128 16-byte stores
 *  10 = refill threshold of the shadow-ray kernel alone (0 = follow option 6)
 *  13 = A/B of the work order (default 0): enumerate the macro blocks of a whole, unsharded frame along a Z-order curve instead
 *       of row by row (north_star: "Morton/tile-ordered ray assignment"; measured: profiles/r02_ns1.md)
 *  15 = ray binning of picker batches (default 0 = off; SURVEY §2.2 "optional Morton/direction-octant binning for the 16 M incoherent
 *       config"): batches of at least 64 Ki rays are traced in Z-order of their ORIGIN cells — value & 15 = bits per axis of the cell
 *       grid over the octree (1..8), + 16 = the ray's direction octant is appended below the cell code; at most 24 key bits. A counting
 *       sort in one atomic pass (histogram + rank, exclusive scan, scatter: 5 small launches in front of the picker kernel, inside its
 *       timed region). Results stay in task order and are bit-identical. Measured on configs[3] (16 Mi random rays, 0.8 GB world): a
 *       LOSS — 4.87 vs 6.37 Grays/s: the pre-pass costs 0.50 ms, and the picker kernel, now gathering tasks and scattering results,
 *       2.83 instead of 2.57 ms (profiles/r02_picker_binning.md); hence off
 *  14 = LIFO hand-over of the wavefront buffers (default 0 = off; bit 0: hit records and shadow-list entries are written with the default
 *       L2 policy instead of streaming stores and the consuming kernel starts with what was written LAST; bit 1: shade_kernel discards
 *       every record line it has read (discard.global.L2: a dead dirty line needs no write-back) — vx_read_hit_records then refuses;
 *       bit 2: trace_shadow_kernel discards its list the same way). Pixels are bit-identical. Measured: 1.506 vs 1.471 ms per 4K frame —
 *       265 MB of records do not fit next to the 33 MB SVO in the 126 MB L2, they push the SVO out and trace_shadow_kernel pays for
 *       it (0.498 vs 0.469 ms); profiles/r02_lifo.md
 *  12 = clip rays against the occupied box of the world (default 1): after every commit a small kernel finds the box that holds
 *       every voxel (chunk granularity); a ray that has left it can only miss, so its traversal stops there instead of walking
 *       the empty cells to the edge of the octree. Results are bit-identical; only the iteration counters shrink. 0 = off
 *  11 = overlapped wavefront (0 off, 1 on, 2 = default: on for launches small enough that kernel tails matter): shade_kernel is launched on its own stream next to trace_primary_kernel and each of its
 *       CTAs waits for the hit records of ITS 32x4-pixel strip (a completion counter per strip) instead of for the whole kernel,
 *       so shading fills the SMs that the tracing kernel's tail leaves idle; 0 = strictly one kernel after the other
 *   7 = the same for the picker kernel (default 20; 16 / 20 / 24 / 28 measured 5.86 / 6.00 / 5.92 / 5.46 Grays/s in round 2: incoherent rays differ in length by 100x; 4.85 vs 2.84 Grays/s
 *       against threshold 1 on 16 Mi random rays, profiles/r01_v2_*) */
int vx_set_option(VxCtx* ctx, uint32_t option, uint64_t value);

/* How many kernels of this library were launched on this ctx since creation. */
uint64_t vx_launch_count(const VxCtx* ctx);

/* Library build id string (compile flags, arch). */
const char* vx_build_info(void);

/* ---- single process, several GPUs ----------------------------------------------------------------------------------------
 * The reference engine is one process (src/gamelogic/game.rs:102-160) with one graphics::Svo (src/graphics/svo.rs:109-255). A
 * VxGroup is that Svo spread over the GPUs of one box: one VxCtx per device, the SVO replicated on every device, a frame cut
 * into image-space shards. The calls mirror the single-GPU ones one to one; a group of one device forwards to them.
 *   vx_group_create            contexts on `devices`, peer access to devices[0]'s memory, NCCL communicators (ncclCommInitAll;
 *                              NCCL is opened with dlopen("libnccl.so.2"), a missing / failing NCCL is VX_E_NCCL)
 *   vx_group_svo_host_mirror   the pinned mirror Esvo::write_changes_to writes into (devices[0]'s)
 *   vx_group_svo_commit        = vx_svo_commit for every replica: ONE packed H2D copy to devices[0], ncclBroadcast over NVLink to
 *                              the others' staging buffers, a scatter kernel on each; ordered behind frames / ray batches in
 *                              flight on every device by events, the CPU does not wait
 *   vx_group_render            frame for GPU consumers: interleaved macro blocks, every device's shade / shadow kernels store their
 *                              finished pixels into devices[0]'s RGBA32F framebuffer over NVLink peer memory; afterwards that
 *                              context's render stream (vx_group_read_frame_*, vx_frame_device_ptr(vx_group_ctx(g, 0))) sees the
 *                              whole frame. VX_E_STATE if a device cannot map devices[0]'s memory.
 *   vx_group_render_read_rgba8 frame for the host (Framebuffer::read_pixels): whole 16-pixel stripes per device (VX_SHARD_ROWS),
 *                              each device DMAs its stripes into rgba8_out itself — one PCIe link per GPU instead of a gather
 *                              through one. rgba8_out should be pinned AND portable (vx_group_host_frame hands one out).
 *                              Returns when the whole frame is in rgba8_out.
 *   vx_group_raycast           contiguous slices of the task array, one per device; synchronous like vx_raycast
 * Threading: like a VxCtx, calls on one group are serialised by the caller; internally one worker thread per device issues
 * that device's launches. */
typedef struct VxGroup VxGroup;
int vx_group_create(const VxConfig* cfg /* .device ignored */, const int* devices, uint32_t n_devices, VxGroup** out);
void vx_group_destroy(VxGroup* group);
const char* vx_group_last_error(const VxGroup* group /* NULL: last vx_group_create failure of this thread */);
uint32_t vx_group_size(const VxGroup* group);
VxCtx* vx_group_ctx(VxGroup* group, uint32_t index);   /* the context of devices[index] (statistics, options, zero-copy frame access) */
int vx_group_set_materials(VxGroup* group, const VxMaterial* materials, uint32_t count);
int vx_group_set_textures(VxGroup* group, const uint8_t* rgba8, uint32_t width, uint32_t height, uint32_t layers, uint32_t mip_levels);
int vx_group_set_option(VxGroup* group, uint32_t option, uint64_t value);
uint8_t* vx_group_svo_host_mirror(VxGroup* group);
int vx_group_svo_set_hot_range(VxGroup* group, uint64_t offset, uint64_t length);
int vx_group_svo_commit(VxGroup* group, float octree_scale, const VxRange* dirty, uint32_t n_dirty, uint64_t used_bytes, uint32_t depth);
int vx_group_stats(const VxGroup* group, VxStats* out);
int vx_group_render(VxGroup* group, const VxRenderParams* params, uint32_t width, uint32_t height);
int vx_group_wait(VxGroup* group);
int vx_group_read_frame_rgba8(VxGroup* group, uint8_t* rgba8_out);
int vx_group_read_frame_rgba32f(VxGroup* group, float* rgba32f_out);
uint8_t* vx_group_host_frame(VxGroup* group, uint64_t bytes);   /* group-owned pinned portable host memory, at least `bytes` */
int vx_group_render_read_rgba8(VxGroup* group, const VxRenderParams* params, uint32_t width, uint32_t height, uint8_t* rgba8_out, uint32_t bands);
int vx_group_raycast(VxGroup* group, const VxPickerTask* tasks, uint64_t n, VxPickerResult* results);

#ifdef __cplusplus
}
#endif
#endif /* VOXELRT_H */
