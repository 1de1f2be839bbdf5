// capi.cpp — flat C entry points (vxh_*) over the C++ host mirror so that Python (ctypes) tests and
// bench.py can drive it. The host mirror is CPU-only code; the only GPU access is through vx_* of
// libvoxelrt inside vxh::Svo.
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>

#include <atomic>
#include <thread>
#include "esvo.hpp"
#include "picker.hpp"
#include "svo.hpp"
#include "world.hpp"
#include "chunkloader.hpp"

using namespace vxh;

static thread_local std::string g_err;
#define VXH_TRY try {
#define VXH_CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

const char* vxh_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------ world --

void* vxh_world_new(uint32_t radius, int32_t cx, int32_t cy, int32_t cz, uint32_t seed, int no_lod) {
    WorldSvo* w = new WorldSvo();
    w->space.dst = radius; w->space.center = ChunkPos{cx, cy, cz};
    w->terrain.seed = seed; w->no_lod = no_lod != 0;
    return w;
}
// SVO_TYPE of the world (0 = ESVO, 1 = CSVO); must be chosen before the first chunk is set.
void vxh_world_set_format(void* w, int format) { ((WorldSvo*)w)->format = format ? SvoFormat::Csvo : SvoFormat::Esvo; }
int vxh_world_format(void* w) { return (int)((WorldSvo*)w)->format; }
void vxh_world_free(void* w) { delete (WorldSvo*)w; }
uint64_t vxh_world_generate(void* w, int32_t y0, int32_t y1, int threads) { return ((WorldSvo*)w)->generate(y0, y1, threads); }
// terrain kind: 0 = stand-in noise, 1 = the reference's generator (noise 0.8.2 Perlin, gamelogic/worldgen.rs)
void vxh_world_set_terrain(void* w, int kind) { ((WorldSvo*)w)->terrain.set_kind(kind); }
// worldgen.rs noise_tests::get (88-101): Noise{frequency, octaves, spline (-1,0)..(1,1)}.get(&Perlin::new(seed), x, z)
double vxh_kat_worldgen_noise(uint32_t seed, float frequency, int octaves, double x, double z) {
    static const float PTS[2][2] = {{-1.0f, 0.0f}, {1.0f, 1.0f}};
    const RefPerlin p(seed);
    return RefNoise{frequency, octaves, PTS, 2}.get(p, x, z);
}
double vxh_kat_spline(const float* pts, int n, double x) { return RefNoise{1.0f, 1, (const float (*)[2])pts, n}.spline(x); }
int32_t vxh_world_height_at(void* w, int32_t x, int32_t z) { return ((WorldSvo*)w)->terrain.height_at(x, z); }
uint64_t vxh_world_chunk_count(void* w) { return ((WorldSvo*)w)->leaf_ids.size(); }

// Builds a chunk the way the reference's tests do (Chunk::new with pooled storage = expand_to(5),
// chunk.rs:20-38,104-123; optional storage.compact(), svo_shader_tests.rs:85) and puts it at SVO
// position (sx,sy,sz) directly. xyzid = n * (x, y, z, block id).
int vxh_world_set_leaf_blocks(void* wp, uint32_t sx, uint32_t sy, uint32_t sz, uint64_t uid, const uint32_t* xyzid, uint32_t n,
                              uint8_t lod, int compact) {
    VXH_TRY
    WorldSvo* w = (WorldSvo*)wp;
    Octree<BlockId> storage;
    storage.expand_to(5);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t* e = xyzid + 4 * (size_t)i;
        if (e[3] == 0) storage.remove_leaf(Position{e[0], e[1], e[2]});
        else storage.set_leaf(Position{e[0], e[1], e[2]}, e[3]);
    }
    if (compact) storage.compact();
    if (w->format == SvoFormat::Csvo) w->csvo.set_leaf(Position{sx, sy, sz}, CsvoChunk::from_octree(uid, storage, lod), true);
    else w->esvo.set_leaf(Position{sx, sy, sz}, SerializedChunk::from_octree(0, 0, 0, uid, storage, lod), true);
    return 0;
    VXH_CATCH(-1)
}

// dense 32^3 chunk (index x + 32*(y + 32*z)) at SVO position
int vxh_world_set_leaf_dense(void* wp, uint32_t sx, uint32_t sy, uint32_t sz, uint64_t uid, const uint32_t* blocks, uint8_t lod) {
    VXH_TRY
    WorldSvo* w = (WorldSvo*)wp;
    if (w->format == SvoFormat::Csvo) w->csvo.set_leaf(Position{sx, sy, sz}, CsvoChunk::from_dense(uid, blocks, lod), true);
    else w->esvo.set_leaf(Position{sx, sy, sz}, SerializedChunk::from_dense(0, 0, 0, uid, blocks, lod), true);
    return 0;
    VXH_CATCH(-1)
}

// The loaded chunks (world chunk coordinates), the dense 32^3 block array + LOD each one was serialized from, and where its
// records live in the RangeBuffer: what a device-side serializer needs to redo the host's work and be compared with it.
uint64_t vxh_world_chunk_list(void* wp, int32_t* xyz, uint64_t cap) {
    WorldSvo* w = (WorldSvo*)wp;
    uint64_t i = 0;
    for (auto& kv : w->leaf_ids) {
        if (i < cap) { xyz[3 * i] = std::get<0>(kv.first); xyz[3 * i + 1] = std::get<1>(kv.first); xyz[3 * i + 2] = std::get<2>(kv.first); }
        ++i;
    }
    return i;
}
int vxh_world_fill_chunk(void* wp, int32_t cx, int32_t cy, int32_t cz, uint32_t* blocks) {   // returns the chunk's LOD, -1 if empty
    WorldSvo* w = (WorldSvo*)wp;
    ChunkPos p{cx, cy, cz};
    std::vector<BlockId> b;
    if (!w->fill_chunk(p, w->make_column(cx, cz), b)) return -1;
    std::memcpy(blocks, b.data(), b.size() * sizeof(BlockId));
    return (int)w->lod_for(p);
}
int vxh_world_chunk_range(void* wp, int32_t cx, int32_t cy, int32_t cz, VxRange* out) {
    WorldSvo* w = (WorldSvo*)wp;
    auto& m = w->range_buffer().id_to_range;
    auto it = m.find(WorldSvo::chunk_uid(ChunkPos{cx, cy, cz}));
    if (it == m.end()) return 0;
    *out = VxRange{it->second.start, it->second.length};
    return 1;
}

// The CPU side of the chunk-serialization comparison: n dense chunks through serialize_dense_chunk on `threads` workers (the
// reference runs SerializedChunk::new on its job-system threads, worldsvo.rs:90-99). Returns the total bytes of records.
uint64_t vxh_serialize_dense_batch(const uint32_t* blocks, uint32_t n, const uint8_t* lods, int threads) {
    if (threads < 1) threads = 1;
    std::atomic<uint32_t> next{0};
    std::atomic<uint64_t> total{0};
    auto worker = [&]() {
        std::vector<uint32_t> dst;
        for (;;) {
            uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            dst.clear();
            serialize_dense_chunk(blocks + (size_t)i * 32768, dst, lods ? lods[i] : 0);
            total.fetch_add(dst.size() * 4);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    return total.load();
}

// world-space block edit on top of the terrain: re-serialises the owning chunk (dirty-range producer)
int vxh_world_edit_block(void* wp, int32_t wx, int32_t wy, int32_t wz, uint32_t id) {
    VXH_TRY
    WorldSvo* w = (WorldSvo*)wp;
    ChunkPos p{wx >> 5, wy >> 5, wz >> 5};
    w->edits[{p.x, p.y, p.z}].push_back({(uint32_t)(wx & 31), (uint32_t)(wy & 31), (uint32_t)(wz & 31), id});
    w->regenerate_chunk(p);
    return 0;
    VXH_CATCH(-1)
}

void vxh_world_serialize(void* w) { ((WorldSvo*)w)->serialize(); }
uint32_t vxh_world_depth(void* w) { return ((WorldSvo*)w)->depth(); }
uint64_t vxh_world_size_bytes(void* w) { return ((WorldSvo*)w)->size_in_bytes(); }
uint64_t vxh_world_header_bytes(void* w) { return ((WorldSvo*)w)->header_bytes(); }
uint64_t vxh_world_write_to(void* w, uint8_t* dst) { return ((WorldSvo*)w)->write_to(dst); }
int vxh_world_write_changes_to(void* w, uint8_t* dst, uint64_t dst_len, int reset) {
    return ((WorldSvo*)w)->write_changes_to(dst, dst_len, reset != 0) ? 0 : -1;
}
uint32_t vxh_world_dirty_ranges(void* w, VxRange* out, uint32_t cap) {
    auto& rs = ((WorldSvo*)w)->range_buffer().updated_ranges;
    for (uint32_t i = 0; i < rs.size() && i < cap; ++i) out[i] = VxRange{rs[i].start, rs[i].length};
    return (uint32_t)rs.size();
}
// Marks the whole RangeBuffer dirty so that the next write_changes_to / Svo::update re-uploads everything (a fresh
// graphics::Svo attached to an already serialized world; the reference never re-attaches, it owns one GL buffer).
void vxh_world_mark_all_dirty(void* w) {
    auto& b = ((WorldSvo*)w)->range_buffer();
    b.updated_ranges.clear();
    if (!b.bytes.empty()) b.updated_ranges.push_back(Range{0, b.bytes.size()});
}
void vxh_world_root_range(void* w, uint64_t* off, uint64_t* len) {
    Range r = ((WorldSvo*)w)->root_range();
    *off = r.start; *len = r.length;
}
void vxh_world_root_info(void* w, uint64_t* buf_offset, uint8_t masks_depth[3]) {
    auto& ri = ((WorldSvo*)w)->esvo.root_info;
    if (!ri) { *buf_offset = 0; masks_depth[0] = masks_depth[1] = masks_depth[2] = 0; return; }
    *buf_offset = ri->buf_offset;
    masks_depth[0] = ri->serialization.child_mask; masks_depth[1] = ri->serialization.leaf_mask; masks_depth[2] = ri->serialization.depth;
}
void vxh_world_cnv_block_pos(void* w, const float in[3], float out[3]) {
    Vec3 r = ((WorldSvo*)w)->space.cnv_block_pos(Vec3{in[0], in[1], in[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void vxh_world_cnv_svo_pos(void* w, const float in[3], float out[3]) {
    Vec3 r = ((WorldSvo*)w)->space.cnv_svo_pos(Vec3{in[0], in[1], in[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int vxh_world_cnv_chunk_pos(void* w, int32_t cx, int32_t cy, int32_t cz, uint32_t out[3]) {
    Position p;
    if (!((WorldSvo*)w)->space.cnv_chunk_pos(ChunkPos{cx, cy, cz}, p)) return 0;
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
    return 1;
}
uint8_t vxh_calculate_lod(int32_t ccx, int32_t ccy, int32_t ccz, int32_t px, int32_t py, int32_t pz) {
    return calculate_lod(ChunkPos{ccx, ccy, ccz}, ChunkPos{px, py, pz});
}

// ------------------------------------------------------- serializer KATs --

// Octree<BlockId>: set_leaf each (x,y,z,id), expand_to(expand), optional compact, serialise with lod.
// Mirrors the construction in esvo.rs:564-573,863-872. Returns words written (or needed).
uint64_t vxh_kat_block_octree(const uint32_t* xyzid, uint32_t n, uint8_t expand_to, int compact, uint8_t lod, uint32_t* out, uint64_t cap,
                              uint8_t result[3]) {
    Octree<BlockId> t;
    for (uint32_t i = 0; i < n; ++i) t.set_leaf(Position{xyzid[4 * i], xyzid[4 * i + 1], xyzid[4 * i + 2]}, xyzid[4 * i + 3]);
    t.expand_to(expand_to);
    if (compact) t.compact();
    std::vector<uint32_t> dst;
    SerializationResult r = serialize_block_octree(t, dst, lod);
    result[0] = r.child_mask; result[1] = r.leaf_mask; result[2] = r.depth;
    for (uint64_t i = 0; i < dst.size() && i < cap; ++i) out[i] = dst[i];
    return dst.size();
}

// CSVO: Octree<BlockId> built the same way, SerializedChunk::serialize_octant(root, tree.depth() - depth_minus, 0, materials)
// (csvo.rs:592-712). Returns node bytes written (or needed); *n_materials receives the material count.
uint64_t vxh_kat_csvo_octant(const uint32_t* xyzid, uint32_t n, uint8_t expand_to, int compact, uint8_t depth_minus, uint8_t* out, uint64_t cap,
                             uint32_t* materials, uint32_t materials_cap, uint32_t* n_materials) {
    Octree<BlockId> t;
    for (uint32_t i = 0; i < n; ++i) t.set_leaf(Position{xyzid[4 * i], xyzid[4 * i + 1], xyzid[4 * i + 2]}, xyzid[4 * i + 3]);
    t.expand_to(expand_to);
    if (compact) t.compact();
    std::vector<BlockId> mats;
    if (!t.root) { *n_materials = 0; return 0; }
    const uint32_t root = *t.root;
    std::vector<uint8_t> dst = csvo_serialize_octant(t, root, (uint8_t)(t.depth() - depth_minus), 0, mats);
    for (uint64_t i = 0; i < dst.size() && i < cap; ++i) out[i] = dst[i];
    for (uint32_t i = 0; i < mats.size() && i < materials_cap; ++i) materials[i] = mats[i];
    *n_materials = (uint32_t)mats.size();
    return dst.size();
}
// CSVO dense fast path vs Chunk::fill_with + generic serializer on the same 32^3 array: returns 1 when nodes and materials agree
int vxh_csvo_dense_equals_generic(const uint32_t* blocks, uint8_t lod) {
    Octree<BlockId> t;
    t.construct_octants_with(5, [blocks](Position p) -> std::optional<BlockId> {
        BlockId b = blocks[(size_t)p.x + 32 * ((size_t)p.y + 32 * (size_t)p.z)];
        if (b == 0) return std::nullopt;
        return b;
    });
    CsvoChunk a = CsvoChunk::from_octree(1, t, lod), b = CsvoChunk::from_dense(1, blocks, lod);
    return a.has_buffer == b.has_buffer && a.lod == b.lod && a.buffer == b.buffer && a.materials == b.materials;
}
// RangeBuffer image + root offset of the world's Csvo (csvo.rs:330-391)
uint64_t vxh_world_csvo_root_offset(void* w) { auto& r = ((WorldSvo*)w)->csvo.root_info; return r ? (uint64_t)*r : ~0ull; }
uint64_t vxh_world_range_bytes(void* w, uint8_t* out, uint64_t cap) {
    auto& b = ((WorldSvo*)w)->range_buffer().bytes;
    for (uint64_t i = 0; i < b.size() && i < cap; ++i) out[i] = b[i];
    return b.size();
}

// dense fast path vs Chunk::fill_with + generic serializer on the same 32^3 array
uint64_t vxh_serialize_dense(const uint32_t* blocks, uint8_t lod, uint32_t* out, uint64_t cap, uint8_t result[3]) {
    std::vector<uint32_t> dst;
    SerializationResult r = serialize_dense_chunk(blocks, dst, lod);
    result[0] = r.child_mask; result[1] = r.leaf_mask; result[2] = r.depth;
    for (uint64_t i = 0; i < dst.size() && i < cap; ++i) out[i] = dst[i];
    return dst.size();
}
uint64_t vxh_serialize_filled(const uint32_t* blocks, uint8_t lod, uint32_t* out, uint64_t cap, uint8_t result[3]) {
    Octree<BlockId> t;
    t.construct_octants_with(5, [blocks](Position p) -> std::optional<BlockId> {
        BlockId b = blocks[(size_t)p.x + 32 * ((size_t)p.y + 32 * (size_t)p.z)];
        if (b == 0) return std::nullopt;
        return b;
    });
    std::vector<uint32_t> dst;
    SerializationResult r = serialize_block_octree(t, dst, lod);
    result[0] = r.child_mask; result[1] = r.leaf_mask; result[2] = r.depth;
    for (uint64_t i = 0; i < dst.size() && i < cap; ++i) out[i] = dst[i];
    return dst.size();
}

// Esvo<u32> (the reference's test fake, worldsvo.rs:236-245) for esvo.rs:745-858
void* vxh_esvo32_new(void) { return new Esvo<U32Leaf>(); }
void vxh_esvo32_free(void* e) { delete (Esvo<U32Leaf>*)e; }
void vxh_esvo32_set_leaf(void* e, uint32_t x, uint32_t y, uint32_t z, uint32_t v, int serialize, uint32_t out_leaf[2]) {
    auto r = ((Esvo<U32Leaf>*)e)->set_leaf(Position{x, y, z}, U32Leaf{v}, serialize != 0);
    out_leaf[0] = r.first.parent; out_leaf[1] = r.first.idx;
}
// returns 1 and *old = replaced value if the target held a leaf
int vxh_esvo32_move_leaf(void* e, uint32_t parent, uint32_t idx, uint32_t x, uint32_t y, uint32_t z, uint32_t out_leaf[2], uint32_t* old) {
    auto r = ((Esvo<U32Leaf>*)e)->move_leaf(LeafId{parent, (uint8_t)idx}, Position{x, y, z});
    out_leaf[0] = r.first.parent; out_leaf[1] = r.first.idx;
    if (r.second) { *old = r.second->v; return 1; }
    return 0;
}
int vxh_esvo32_remove_leaf(void* e, uint32_t parent, uint32_t idx, uint32_t* old) {
    auto r = ((Esvo<U32Leaf>*)e)->remove_leaf(LeafId{parent, (uint8_t)idx});
    if (r) { *old = r->v; return 1; }
    return 0;
}
void vxh_esvo32_serialize(void* e) { ((Esvo<U32Leaf>*)e)->serialize(); }
void vxh_esvo32_root_info(void* e, uint64_t* buf_offset, uint8_t masks_depth[3]) {
    auto& ri = ((Esvo<U32Leaf>*)e)->root_info;
    *buf_offset = ri ? ri->buf_offset : 0;
    masks_depth[0] = ri ? ri->serialization.child_mask : 0; masks_depth[1] = ri ? ri->serialization.leaf_mask : 0;
    masks_depth[2] = ri ? ri->serialization.depth : 0;
}
uint64_t vxh_esvo32_bytes(void* e, uint8_t* out, uint64_t cap) {
    auto& b = ((Esvo<U32Leaf>*)e)->buffer.bytes;
    if (out) std::memcpy(out, b.data(), b.size() < cap ? b.size() : cap);
    return b.size();
}
// kind 0 = free_ranges, 1 = updated_ranges
uint32_t vxh_esvo32_ranges(void* e, int kind, VxRange* out, uint32_t cap) {
    auto& rs = kind == 0 ? ((Esvo<U32Leaf>*)e)->buffer.free_ranges : ((Esvo<U32Leaf>*)e)->buffer.updated_ranges;
    for (uint32_t i = 0; i < rs.size() && i < cap; ++i) out[i] = VxRange{rs[i].start, rs[i].length};
    return (uint32_t)rs.size();
}
int vxh_esvo32_range_of(void* e, uint64_t id, VxRange* out) {
    auto& m = ((Esvo<U32Leaf>*)e)->buffer.id_to_range;
    auto it = m.find(id);
    if (it == m.end()) return 0;
    *out = VxRange{it->second.start, it->second.length};
    return 1;
}
void vxh_esvo32_clear_updated(void* e) { ((Esvo<U32Leaf>*)e)->buffer.updated_ranges.clear(); }
uint64_t vxh_esvo32_write_to(void* e, uint8_t* dst) { return ((Esvo<U32Leaf>*)e)->write_to(dst); }
int vxh_esvo32_write_changes_to(void* e, uint8_t* dst, uint64_t dst_len, int reset) {
    return ((Esvo<U32Leaf>*)e)->write_changes_to(dst, dst_len, reset != 0) ? 0 : -1;
}

// RangeBuffer alone (internal.rs:279-455 tests)
void* vxh_rangebuf_new(void) { return new RangeBuffer(); }
void vxh_rangebuf_free(void* r) { delete (RangeBuffer*)r; }
uint64_t vxh_rangebuf_insert(void* r, uint64_t id, const uint8_t* buf, uint64_t len) { return ((RangeBuffer*)r)->insert(id, buf, len); }
void vxh_rangebuf_remove(void* r, uint64_t id) { ((RangeBuffer*)r)->remove(id); }
uint64_t vxh_rangebuf_bytes(void* r, uint8_t* out, uint64_t cap) {
    auto& b = ((RangeBuffer*)r)->bytes;
    if (out) std::memcpy(out, b.data(), b.size() < cap ? b.size() : cap);
    return b.size();
}
uint32_t vxh_rangebuf_ranges(void* r, int kind, VxRange* out, uint32_t cap) {
    auto& rs = kind == 0 ? ((RangeBuffer*)r)->free_ranges : ((RangeBuffer*)r)->updated_ranges;
    for (uint32_t i = 0; i < rs.size() && i < cap; ++i) out[i] = VxRange{rs[i].start, rs[i].length};
    return (uint32_t)rs.size();
}

void* vxh_rangebuf_with_capacity(uint64_t n) { return new RangeBuffer((size_t)n); }
// id -> range table, sorted by id: out[i] = {id, start, length}
uint32_t vxh_rangebuf_ids(void* r, uint64_t* out, uint32_t cap) {
    auto& m = ((RangeBuffer*)r)->id_to_range;
    std::vector<uint64_t> ids;
    for (auto& kv : m) ids.push_back(kv.first);
    std::sort(ids.begin(), ids.end());
    for (uint32_t i = 0; i < ids.size() && i < cap; ++i) { out[3 * i] = ids[i]; out[3 * i + 1] = m[ids[i]].start; out[3 * i + 2] = m[ids[i]].length; }
    return (uint32_t)ids.size();
}
// RangeBuffer::merge_ranges (internal.rs:252-272) on its own
uint32_t vxh_merge_ranges(VxRange* ranges, uint32_t n) {
    std::vector<Range> rs;
    for (uint32_t i = 0; i < n; ++i) rs.push_back(Range{(size_t)ranges[i].offset, (size_t)ranges[i].length});
    RangeBuffer::merge(rs);
    for (uint32_t i = 0; i < rs.size(); ++i) ranges[i] = VxRange{rs[i].start, rs[i].length};
    return (uint32_t)rs.size();
}

// Svo::shift_chunks on a bare Esvo<u32> (worldsvo.rs:249-375 tests). chunks = n x {x, y, z}, ids = n x {parent, idx} in/out;
// returns the number of chunks that are still inside the window (their rows come first, in key order).
uint32_t vxh_kat_shift_chunks(void* e, int32_t cx, int32_t cy, int32_t cz, uint32_t dst, int32_t* chunks, uint32_t* ids, uint32_t n) {
    std::map<std::tuple<int32_t, int32_t, int32_t>, LeafId> leaf_ids;
    for (uint32_t i = 0; i < n; ++i) leaf_ids[{chunks[3 * i], chunks[3 * i + 1], chunks[3 * i + 2]}] = LeafId{ids[2 * i], (uint8_t)ids[2 * i + 1]};
    SvoCoordSpace cs; cs.center = ChunkPos{cx, cy, cz}; cs.dst = dst;
    shift_chunks(cs, leaf_ids, *(Esvo<U32Leaf>*)e);
    uint32_t k = 0;
    for (auto& kv : leaf_ids) {
        chunks[3 * k] = std::get<0>(kv.first); chunks[3 * k + 1] = std::get<1>(kv.first); chunks[3 * k + 2] = std::get<2>(kv.first);
        ids[2 * k] = kv.second.parent; ids[2 * k + 1] = kv.second.idx;
        ++k;
    }
    return k;
}
int64_t vxh_esvo32_get_leaf(void* e, uint32_t x, uint32_t y, uint32_t z) {
    const U32Leaf* v = ((Esvo<U32Leaf>*)e)->get_leaf(Position{x, y, z});
    return v ? (int64_t)v->v : -1;
}
// storage / generator results arriving for one chunk (Svo::set_chunk after serialization, worldsvo.rs:90-99) and chunk unloads (:101-108)
int vxh_world_load_chunk(void* w, int32_t cx, int32_t cy, int32_t cz) { return ((WorldSvo*)w)->load_chunk(ChunkPos{cx, cy, cz}) ? 1 : 0; }
void vxh_world_remove_chunk(void* w, int32_t cx, int32_t cy, int32_t cz) { ((WorldSvo*)w)->remove_chunk(ChunkPos{cx, cy, cz}); }
// the player entered another chunk: re-centre the SVO window (1 if the centre changed)
int vxh_world_set_center(void* w, int32_t cx, int32_t cy, int32_t cz) { return ((WorldSvo*)w)->set_center(ChunkPos{cx, cy, cz}) ? 1 : 0; }

// ------------------------------------------------------------ chunk loader --
// systems::chunkloader::ChunkLoader. Events: rows of {kind (0 Load, 1 Unload, 2 LodChange), x, y, z, lod}.
void* vxh_chunkloader_new(uint32_t radius, int32_t start_y, int32_t end_y) { return start_y < end_y ? new ChunkLoader(radius, start_y, end_y) : nullptr; }
void vxh_chunkloader_free(void* l) { delete (ChunkLoader*)l; }
void vxh_chunkloader_set_radius(void* l, uint32_t r) { ((ChunkLoader*)l)->set_radius(r); }
int vxh_chunkloader_is_loaded(void* l, int32_t x, int32_t y, int32_t z) { return ((ChunkLoader*)l)->is_loaded(ChunkPos{x, y, z}) ? 1 : 0; }
void vxh_chunkloader_add_loaded(void* l, int32_t x, int32_t y, int32_t z, uint32_t lod) { ((ChunkLoader*)l)->add_loaded_chunk(ChunkPos{x, y, z}, (uint8_t)lod); }
uint64_t vxh_chunkloader_loaded_count(void* l) { return ((ChunkLoader*)l)->loaded_count(); }
// Returns the number of events; the loader's state always advances, `out` receives the first `cap` of them.
uint64_t vxh_chunkloader_update(void* l, float x, float y, float z, int32_t* out, uint64_t cap) {
    std::vector<ChunkEvent> ev = ((ChunkLoader*)l)->update(x, y, z);
    for (uint64_t i = 0; i < ev.size() && i < cap; ++i) {
        out[5 * i] = ev[i].kind; out[5 * i + 1] = ev[i].pos.x; out[5 * i + 2] = ev[i].pos.y; out[5 * i + 3] = ev[i].pos.z; out[5 * i + 4] = ev[i].lod;
    }
    return ev.size();
}

// ----------------------------------------------------------------- octree --
// Octree<u32> alone (src/world/hds/octree.rs:508-866 tests). Results: out[0] = parent, out[1] = idx, out[2] = 1 if a value was
// replaced / removed, out[3] = that value.
using U32Tree = Octree<uint32_t>;
void* vxh_octree_new(void) { return new U32Tree(); }
void vxh_octree_free(void* t) { delete (U32Tree*)t; }
void vxh_octree_set_leaf(void* t, uint32_t x, uint32_t y, uint32_t z, uint32_t v, int64_t out[4]) {
    auto r = ((U32Tree*)t)->set_leaf(Position{x, y, z}, v);
    out[0] = r.first.parent; out[1] = r.first.idx; out[2] = r.second ? 1 : 0; out[3] = r.second ? *r.second : 0;
}
void vxh_octree_move_leaf(void* t, uint32_t parent, uint32_t idx, uint32_t x, uint32_t y, uint32_t z, int64_t out[4]) {
    auto r = ((U32Tree*)t)->move_leaf(LeafId{parent, (uint8_t)idx}, Position{x, y, z});
    out[0] = r.first.parent; out[1] = r.first.idx; out[2] = r.second ? 1 : 0; out[3] = r.second ? *r.second : 0;
}
// out[0] = parent, out[1] = idx (or -1, -1 when nothing was there), out[2] / out[3] = removed value
void vxh_octree_remove_leaf(void* t, uint32_t x, uint32_t y, uint32_t z, int64_t out[4]) {
    auto r = ((U32Tree*)t)->remove_leaf(Position{x, y, z});
    out[0] = r.second ? (int64_t)r.second->parent : -1; out[1] = r.second ? (int64_t)r.second->idx : -1;
    out[2] = r.first ? 1 : 0; out[3] = r.first ? *r.first : 0;
}
int64_t vxh_octree_remove_leaf_by_id(void* t, uint32_t parent, uint32_t idx) {
    auto r = ((U32Tree*)t)->remove_leaf_by_id(LeafId{parent, (uint8_t)idx});
    return r ? (int64_t)*r : -1;
}
int64_t vxh_octree_get_leaf(void* t, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t* v = ((const U32Tree*)t)->get_leaf(Position{x, y, z});
    return v ? (int64_t)*v : -1;
}
void vxh_octree_compact(void* t) { ((U32Tree*)t)->compact(); }
// construct_octants_with(depth, f) with f = "value at the listed positions, None elsewhere"; cells = n x {x, y, z, value}
void vxh_octree_construct(void* t, uint32_t depth, const uint32_t* cells, uint32_t n) {
    ((U32Tree*)t)->construct_octants_with((uint8_t)depth, [&](Position p) -> std::optional<uint32_t> {
        for (uint32_t i = 0; i < n; ++i)
            if (cells[4 * i] == p.x && cells[4 * i + 1] == p.y && cells[4 * i + 2] == p.z) return cells[4 * i + 3];
        return std::nullopt;
    });
}
// State dump: head = {root or -1, depth, n_octants, n_free}; octants = n x 18 {parent or -1, children_count, kind[8], value[8]}
// (kind 0 None / 1 Octant / 2 Leaf; value = octant id resp. leaf value); free = the free list in order.
void vxh_octree_dump(void* t, int64_t head[4], int64_t* octants, uint32_t octants_cap, int64_t* free_ids, uint32_t free_cap) {
    const U32Tree& tr = *(const U32Tree*)t;
    head[0] = tr.root ? (int64_t)*tr.root : -1; head[1] = tr.depth(); head[2] = (int64_t)tr.octants.size(); head[3] = (int64_t)tr.free_list().size();
    for (uint32_t i = 0; i < tr.octants.size() && i < octants_cap; ++i) {
        const auto& o = tr.octants[i];
        int64_t* row = octants + 18 * (size_t)i;
        row[0] = o.parent; row[1] = o.children_count;
        for (int k = 0; k < 8; ++k) {
            row[2 + k] = (int64_t)o.kind[k];
            row[10 + k] = o.kind[k] == ChildKind::Leaf ? (int64_t)*tr.leaf_value(i, (uint8_t)k) : (o.kind[k] == ChildKind::Octant ? (int64_t)o.ref[k] : 0);
        }
    }
    for (uint32_t i = 0; i < tr.free_list().size() && i < free_cap; ++i) free_ids[i] = tr.free_list()[i];
}

// ----------------------------------------------------------------- picker --

static void fill_batch(PickerBatch& b, const float* rays, uint32_t n_rays, const float* aabbs, uint32_t n_aabbs) {
    for (uint32_t i = 0; i < n_rays; ++i) {
        const float* r = rays + 7 * (size_t)i;
        b.add_ray(Vec3{r[0], r[1], r[2]}, Vec3{r[3], r[4], r[5]}, r[6]);
    }
    for (uint32_t i = 0; i < n_aabbs; ++i) {
        const float* a = aabbs + 9 * (size_t)i;
        b.add_aabb(Aabb{Vec3{a[0], a[1], a[2]}, Vec3{a[3], a[4], a[5]}, Vec3{a[6], a[7], a[8]}});
    }
}
static void dump_result(const PickerBatchResult& res, float* ray_out, float* aabb_out) {
    for (size_t i = 0; i < res.rays.size(); ++i) {
        float* o = ray_out + 8 * i;
        const RayResult& r = res.rays[i];
        o[0] = r.dst; o[1] = r.inside_voxel ? 1.0f : 0.0f; o[2] = r.pos.x; o[3] = r.pos.y; o[4] = r.pos.z;
        o[5] = r.normal.x; o[6] = r.normal.y; o[7] = r.normal.z;
    }
    for (size_t i = 0; i < res.aabbs.size(); ++i) {
        float* o = aabb_out + 6 * i;
        const AabbResult& a = res.aabbs[i];
        o[0] = a.neg.x; o[1] = a.neg.y; o[2] = a.neg.z; o[3] = a.pos.x; o[4] = a.pos.y; o[5] = a.pos.z;
    }
}

// rays: n * (pos3, dir3, max_dst); aabbs: n * (pos3, offset3, extents3). Returns the task count.
uint64_t vxh_picker_serialize(const float* rays, uint32_t n_rays, const float* aabbs, uint32_t n_aabbs, VxPickerTask* out, uint64_t cap) {
    PickerBatch b;
    fill_batch(b, rays, n_rays, aabbs, n_aabbs);
    std::vector<VxPickerTask> tasks;
    b.serialize_tasks(tasks);
    for (uint64_t i = 0; i < tasks.size() && i < cap; ++i) out[i] = tasks[i];
    return tasks.size();
}
// ray_out: n_rays * (dst, inside, pos3, normal3); aabb_out: n_aabbs * (neg3, pos3)
void vxh_picker_deserialize(const float* rays, uint32_t n_rays, const float* aabbs, uint32_t n_aabbs, const VxPickerResult* results,
                            uint64_t n_results, float* ray_out, float* aabb_out) {
    PickerBatch b;
    fill_batch(b, rays, n_rays, aabbs, n_aabbs);
    PickerBatchResult res;
    b.deserialize_results(results, n_results, res);
    dump_result(res, ray_out, aabb_out);
}

// --------------------------------------------------------- registry / Svo --

void* vxh_registry_new(void) { return new VoxelRegistry(); }
void vxh_registry_free(void* r) { delete (VoxelRegistry*)r; }
int vxh_registry_add_texture(void* r, const char* name, uint32_t w, uint32_t h, const uint8_t* rgba_top_down) {
    VXH_TRY
    ((VoxelRegistry*)r)->add_texture(name, w, h, rgba_top_down);
    return 0;
    VXH_CATCH(-1)
}
void vxh_registry_set_mip_levels(void* r, uint8_t levels) { ((VoxelRegistry*)r)->set_mip_levels(levels); }
// top/side/bottom may be NULL (= no texture)
int vxh_registry_add_material(void* r, uint32_t block, float spec_pow, float spec_strength, const char* top, const char* side,
                              const char* bottom, int with_normals) {
    VXH_TRY
    Material m;
    m.specular(spec_pow, spec_strength);
    if (top) m.top(top);
    if (side) m.side(side);
    if (bottom) m.bottom(bottom);
    if (with_normals) m.with_normals();
    ((VoxelRegistry*)r)->add_material(block, m);
    return 0;
    VXH_CATCH(-1)
}
uint32_t vxh_registry_materials(void* r, VxMaterial* out, uint32_t cap) {
    VXH_TRY
    std::vector<VxMaterial> m = ((VoxelRegistry*)r)->build_material_buffer();
    for (uint32_t i = 0; i < m.size() && i < cap; ++i) out[i] = m[i];
    return (uint32_t)m.size();
    VXH_CATCH(0)
}
// level-0 texel data as uploaded (v-flipped); returns bytes, dims = {w, h, layers, mip_levels}
uint64_t vxh_registry_textures(void* r, uint8_t* out, uint64_t cap, uint32_t dims[4]) {
    VoxelRegistry* reg = (VoxelRegistry*)r;
    std::vector<uint8_t> t = reg->packed_level0();
    dims[0] = reg->width(); dims[1] = reg->height(); dims[2] = reg->layers(); dims[3] = reg->mip_levels();
    if (out) std::memcpy(out, t.data(), t.size() < cap ? t.size() : cap);
    return t.size();
}

void vxh_look_to_rh_inverted(const float eye[3], const float dir[3], const float up[3], float out[16]) { look_to_rh_inverted(eye, dir, up, out); }

#ifndef VXH_WORLD_ONLY   // libvoxelrs_world.so = everything above: the producers of the path's inputs, no dependency on libvoxelrt
void* vxh_svo_new(void* registry, uint64_t size_mb, uint32_t max_w, uint32_t max_h, uint64_t max_rays, int device, uint32_t flags) {
    VXH_TRY
    return new Svo(*(VoxelRegistry*)registry, size_mb, max_w, max_h, max_rays, device, flags);
    VXH_CATCH(nullptr)
}
void vxh_svo_free(void* s) { delete (Svo*)s; }
void* vxh_svo_ctx(void* s) { return ((Svo*)s)->ctx(); }
int vxh_svo_update(void* s, void* world) {
    VXH_TRY
    WorldSvo* w = (WorldSvo*)world;
    if (w->format == SvoFormat::Csvo) ((Svo*)s)->update(w->csvo); else ((Svo*)s)->update(w->esvo);
    return 0;
    VXH_CATCH(-1)
}
void vxh_svo_stats(void* s, uint64_t out[3]) {
    Stats st = ((Svo*)s)->get_stats();
    out[0] = st.used_bytes; out[1] = st.capacity_bytes; out[2] = st.depth;
}

// graphics::svo::RenderParams (svo.rs:85-106) as a flat record for ctypes
struct VxhRenderParams {
    float ambient_intensity;
    float light_dir[3];
    float cam_pos[3];
    float cam_fwd[3];
    float cam_up[3];
    float fov_y_rad;
    float aspect_ratio;
    int32_t has_selected_voxel;
    float selected_voxel[3];
    int32_t render_shadows;
    float shadow_distance;
};
int vxh_svo_render(void* s, const VxhRenderParams* p, uint32_t width, uint32_t height, const VxShard* shard) {
    VXH_TRY
    RenderParams rp;
    rp.ambient_intensity = p->ambient_intensity;
    for (int k = 0; k < 3; ++k) { rp.light_dir[k] = p->light_dir[k]; rp.cam_pos[k] = p->cam_pos[k]; rp.cam_fwd[k] = p->cam_fwd[k]; rp.cam_up[k] = p->cam_up[k]; }
    rp.fov_y_rad = p->fov_y_rad; rp.aspect_ratio = p->aspect_ratio;
    if (p->has_selected_voxel) rp.selected_voxel = Vec3{p->selected_voxel[0], p->selected_voxel[1], p->selected_voxel[2]};
    rp.render_shadows = p->render_shadows != 0; rp.shadow_distance = p->shadow_distance;
    ((Svo*)s)->render(rp, Framebuffer{width, height}, shard);
    return 0;
    VXH_CATCH(-1)
}
// systems::worldsvo::Svo::render (worldsvo.rs:397-409): same, with cam_pos / selected_voxel given in WORLD space
int vxh_worldsvo_render(void* s, void* world, const VxhRenderParams* p, uint32_t width, uint32_t height, const VxShard* shard) {
    VxhRenderParams q = *p;
    WorldSvo* w = (WorldSvo*)world;
    Vec3 c = w->space.cnv_block_pos(Vec3{p->cam_pos[0], p->cam_pos[1], p->cam_pos[2]});
    q.cam_pos[0] = c.x; q.cam_pos[1] = c.y; q.cam_pos[2] = c.z;
    if (p->has_selected_voxel) {
        Vec3 v = w->space.cnv_block_pos(Vec3{p->selected_voxel[0], p->selected_voxel[1], p->selected_voxel[2]});
        q.selected_voxel[0] = v.x; q.selected_voxel[1] = v.y; q.selected_voxel[2] = v.z;
    }
    return vxh_svo_render(s, &q, width, height, shard);
}
// graphics::Svo::raycast (svo.rs:233-255) on a batch of rays + AABBs in SVO space
int vxh_svo_raycast(void* s, const float* rays, uint32_t n_rays, const float* aabbs, uint32_t n_aabbs, float* ray_out, float* aabb_out) {
    VXH_TRY
    PickerBatch b;
    fill_batch(b, rays, n_rays, aabbs, n_aabbs);
    PickerBatchResult res;
    ((Svo*)s)->raycast(b, res);
    dump_result(res, ray_out, aabb_out);
    return 0;
    VXH_CATCH(-1)
}
// <worldsvo::Svo as Raycaster>::raycast (worldsvo.rs:419-435): inputs in world space, ray hit positions mapped back
int vxh_worldsvo_raycast(void* s, void* world, const float* rays, uint32_t n_rays, const float* aabbs, uint32_t n_aabbs, float* ray_out,
                         float* aabb_out) {
    VXH_TRY
    WorldSvo* w = (WorldSvo*)world;
    PickerBatch b;
    fill_batch(b, rays, n_rays, aabbs, n_aabbs);
    for (Ray& r : b.rays) r.pos = w->space.cnv_block_pos(r.pos);
    for (Aabb& a : b.aabbs) a.pos = w->space.cnv_block_pos(a.pos);
    PickerBatchResult res;
    ((Svo*)s)->raycast(b, res);
    for (RayResult& r : res.rays) r.pos = w->space.cnv_svo_pos(r.pos);
    dump_result(res, ray_out, aabb_out);
    return 0;
    VXH_CATCH(-1)
}

#endif  // VXH_WORLD_ONLY

}  // extern "C"
