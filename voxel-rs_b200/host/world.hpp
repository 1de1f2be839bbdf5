// world.hpp — world-level SVO assembly above the serializer: SVO coordinate space, LOD rule, and a
// deterministic synthetic terrain of the reference's "named shape".
//
//   SvoCoordSpace (world <-> SVO space)  src/systems/worldsvo.rs:438-503, src/world/chunk.rs:251-298
//   LOD by 2-D chunk distance            src/systems/chunkloader.rs:127-134
//   loading disc dx^2+dz^2 <= r^2        src/systems/chunkloader.rs:68-73, y-chunks 0..8 (gamelogic/world.rs:85)
//   column -> blocks rule                src/gamelogic/worldgen.rs:294-316 (grass / 3x dirt / stone)
//   chunk emitted iff column spans it    src/gamelogic/worldgen.rs:172-176
//   splines                              src/gamelogic/world.rs:56-78
// The noise itself is NOT the reference's (`noise` 0.8.2 Perlin is a third-party crate absent from the
// checkout); a hash-based gradient noise stands in, so worlds have the reference's shape, not its
// literal voxels.
#pragma once
#include <array>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdint>
#include <map>
#include <thread>
#include <tuple>
#include <type_traits>
#include <vector>

#include "csvo.hpp"
#include "esvo.hpp"
#include "picker.hpp"

namespace vxh {

struct ChunkPos { int32_t x, y, z; };

// worldsvo.rs:438-503
struct SvoCoordSpace {
    ChunkPos center{0, 0, 0};
    uint32_t dst = 0;

    static int32_t floor_div32(int32_t v) { return v >> 5; }

    // world space -> SVO space (worldsvo.rs:452-462 via BlockPos::from / to_point, chunk.rs:270-297)
    Vec3 cnv_block_pos(Vec3 p) const {
        float in[3] = {p.x, p.y, p.z}, out[3];
        const int32_t c[3] = {center.x, center.y, center.z};
        for (int k = 0; k < 3; ++k) {
            float fl = std::floor(in[k]);
            int32_t b = (int32_t)fl;
            float fr = in[k] - std::trunc(in[k]);          // f32::fract
            if (fr != 0.0f && in[k] < 0.0f) fr += 1.0f;
            int32_t chunk = floor_div32(b);
            float rel = (float)(b & 31) + fr;
            int32_t nchunk = (int32_t)dst + (chunk - c[k]);
            int32_t block = (nchunk * 32) | ((int32_t)rel & 31);
            out[k] = (float)block + (rel - std::trunc(rel));
        }
        return Vec3{out[0], out[1], out[2]};
    }

    // SVO space -> world space (worldsvo.rs:465-476)
    Vec3 cnv_svo_pos(Vec3 p) const {
        float in[3] = {p.x, p.y, p.z}, out[3];
        const int32_t c[3] = {center.x, center.y, center.z};
        for (int k = 0; k < 3; ++k) {
            float fl = std::floor(in[k]);
            int32_t b = (int32_t)fl;
            float fr = in[k] - std::trunc(in[k]);
            if (fr != 0.0f && in[k] < 0.0f) fr += 1.0f;
            int32_t chunk = floor_div32(b);
            float rel = (float)(b & 31) + fr;
            int32_t nchunk = c[k] + (chunk - (int32_t)dst);
            int32_t block = (nchunk * 32) | ((int32_t)rel & 31);
            out[k] = (float)block + (rel - std::trunc(rel));
        }
        return Vec3{out[0], out[1], out[2]};
    }

    // worldsvo.rs:481-502: None (false) when outside the y range or the xz disc
    bool cnv_chunk_pos(ChunkPos pos, Position& out) const {
        float r = (float)dst;
        Vec3 p = cnv_block_pos(Vec3{(float)(pos.x * 32), (float)(pos.y * 32), (float)(pos.z * 32)});
        p.x /= 32.0f; p.y /= 32.0f; p.z /= 32.0f;
        float dcy = p.y - r;
        if (dcy < -r || dcy > r) return false;
        float dcx = p.x - r, dcz = p.z - r;
        if (std::fma(dcx, dcx, dcz * dcz) > r * r) return false;
        out = Position{(uint32_t)p.x, (uint32_t)p.y, (uint32_t)p.z};
        return true;
    }
};

// chunkloader.rs:127-134
inline uint8_t calculate_lod(ChunkPos center, ChunkPos pos) {
    float dx = (float)(pos.x - center.x), dz = (float)(pos.z - center.z);
    int d = (int)std::sqrt(dx * dx + dz * dz);
    if (d <= 6) return 5;
    if (d <= 12) return 4;
    if (d <= 19) return 3;
    return 2;
}

// systems::worldsvo::Svo::shift_chunks (worldsvo.rs:158-196): after the SVO coordinate space moved to a new centre chunk, every
// chunk leaf is moved to its new position in SVO space (replacing whatever sat there) and chunks that fell out of the window are
// removed. Leaves are moved inside the world octree; their serialized records stay where they are in the buffer, so a shift dirties
// only the world-root octants. `leaf_ids` is iterated in key order here (the reference iterates an FxHashMap, i.e. in an unspecified
// order; the bookkeeping of overridden leaves makes the result independent of it).
template <typename SvoT>
void shift_chunks(const SvoCoordSpace& cs, std::map<std::tuple<int32_t, int32_t, int32_t>, LeafId>& leaf_ids, SvoT& svo) {
    using LeafT = typename std::decay<decltype(*svo.get_leaf(Position{0, 0, 0}))>::type;
    std::map<std::pair<uint32_t, uint8_t>, LeafT> overridden;
    std::vector<std::tuple<int32_t, int32_t, int32_t>> removed;
    for (auto& kv : leaf_ids) {
        LeafId& leaf_id = kv.second;
        const auto key = std::make_pair(leaf_id.parent, leaf_id.idx);
        Position np;
        if (!cs.cnv_chunk_pos(ChunkPos{std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first)}, np)) {
            // remove the leaf unless another moved leaf already took its slot
            if (!overridden.count(key)) svo.remove_leaf(leaf_id);
            overridden.erase(key);
            removed.push_back(kv.first);
            continue;
        }
        std::pair<LeafId, std::optional<LeafT>> r = [&]() {
            auto it = overridden.find(key);
            if (it != overridden.end()) {
                LeafT value = std::move(it->second);
                overridden.erase(it);
                return svo.set_leaf(np, std::move(value), false);   // only moved: its record may already be in the buffer
            }
            return svo.move_leaf(leaf_id, np);
        }();
        leaf_id = r.first;
        if (r.second) overridden.insert_or_assign(std::make_pair(r.first.parent, r.first.idx), std::move(*r.second));
    }
    for (auto& k : removed) leaf_ids.erase(k);
}

// `noise::Perlin` of the noise crate 0.8.2 (Cargo.lock:936-943; the crate source is not part of the reference checkout), restated
// from its published algorithm and pinned by the reference's own known-answer test `noise_tests::get`
// (src/gamelogic/worldgen.rs:88-101, tests/test_host_worldgen.py) and end-to-end by `gamelogic_world_end_to_end_expected.png`:
//   PermutationTable::new(seed): XorShiftRng (rand_xorshift 0.2) seeded with the bytes {1,0,0,0, seed, seed, seed} (little endian),
//     Fisher-Yates shuffle of 0..255 from the back with rand 0.7.3's `gen_range(0, i + 1)` (widening-multiply rejection sampler);
//   hash(x, y) = perm[perm[x & 255] ^ (y & 255)];
//   perlin_2d: gradients (+-1, +-1) chosen by hash & 3, quintic fade, bilinear blend, x sqrt(2), clamped to [-1, 1].
struct RefPerlin {
    uint8_t perm[256];

    explicit RefPerlin(uint32_t seed = 0) {
        uint32_t st[4] = {1u, seed, seed, seed};   // from_seed: four little-endian u32 (never all zero)
        auto next = [&]() {
            const uint32_t x = st[0], t = x ^ (x << 11);
            st[0] = st[1]; st[1] = st[2]; st[2] = st[3];
            st[3] = st[3] ^ (st[3] >> 19) ^ (t ^ (t >> 8));
            return st[3];
        };
        for (int i = 0; i < 256; ++i) perm[i] = (uint8_t)i;
        for (uint32_t i = 255; i >= 1; --i) {
            const uint32_t range = i + 1, zone = (range << __builtin_clz(range)) - 1u;
            uint32_t j;
            for (;;) {
                const uint64_t m = (uint64_t)next() * range;
                if ((uint32_t)m <= zone) { j = (uint32_t)(m >> 32); break; }
            }
            std::swap(perm[i], perm[j]);
        }
    }
    uint32_t hash(int64_t x, int64_t y) const { return perm[perm[x & 255] ^ (uint32_t)(y & 255)]; }
    static double corner(uint32_t h, double dx, double dy) {
        switch (h & 3u) {
            case 0: return dx + dy;
            case 1: return -dx + dy;
            case 2: return dx - dy;
            default: return -dx - dy;
        }
    }
    static double quintic(double t) { return t * t * t * (t * (t * 6.0 - 15.0) + 10.0); }
    double get(double x, double y) const {
        const double fx = std::floor(x), fy = std::floor(y);
        const int64_t cx = (int64_t)fx, cy = (int64_t)fy;
        const double dx = x - fx, dy = y - fy;
        const double g00 = corner(hash(cx, cy), dx, dy), g10 = corner(hash(cx + 1, cy), dx - 1.0, dy);
        const double g01 = corner(hash(cx, cy + 1), dx, dy - 1.0), g11 = corner(hash(cx + 1, cy + 1), dx - 1.0, dy - 1.0);
        const double u = quintic(dx), v = quintic(dy);
        const double r = (g00 + (g10 - g00) * u + (g01 - g00) * v + (g00 + g11 - g10 - g01) * u * v) * 1.4142135623730951;
        return r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    }
};

// `worldgen::Noise` (src/gamelogic/worldgen.rs:14-77): octaves of Perlin noise pushed through a piecewise-linear spline, with the
// reference's mixed f32 / f64 arithmetic (f32 frequency and spline points, f64 accumulation, fused multiply-adds).
struct RefNoise {
    float frequency; int octaves; const float (*points)[2]; int n_points;

    double value(const RefPerlin& p, double x, double z) const {   // get_noise_value, worldgen.rs:42-54
        double f = (double)frequency, a = 1.0, v = 0.0;
        for (int i = 0; i < octaves; ++i) { v += p.get(std::fma(x, f, 0.5), std::fma(z, f, 0.5)) * a; f *= 2.0; a *= 0.5; }
        return v;
    }
    double spline(double x) const {   // interpolate_spline_points, worldgen.rs:56-77
        if (n_points == 0) return 0.0;
        int rhs = -1;
        for (int i = 0; i < n_points; ++i) if ((double)points[i][0] > x) { rhs = i; break; }
        if (rhs < 0) return (double)points[n_points - 1][1];
        if (rhs == 0) return (double)points[0][1];
        const float* l = points[rhs - 1]; const float* r = points[rhs];
        const float factor = ((float)x - l[0]) / (r[0] - l[0]);
        return std::fma((double)(r[1] - l[1]), (double)factor, (double)l[1]);
    }
    double get(const RefPerlin& p, double x, double z) const { return spline(value(p, x, z)); }
};

// Terrain height function. kind 1 = the reference's generator (Generator::get_height_at, gamelogic/worldgen.rs:191-199, with the
// Config of gamelogic/world.rs:54-78 and Perlin::new(seed)); kind 0 = a deterministic stand-in of the same shape: two fBm layers
// of hash-gradient noise pushed through the reference's splines.
struct Terrain {
    uint32_t seed = 1;
    int kind = 0;
    RefPerlin perlin{1};

    void set_kind(int k) { kind = k; if (k == 1) perlin = RefPerlin(seed); }

    static uint32_t hash2(uint32_t x, uint32_t y, uint32_t s) {
        uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u + s * 0xC2B2AE3Du);
        h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
        return h;
    }
    static double fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
    double grad(int32_t ix, int32_t iz, double dx, double dz, uint32_t s) const {
        static const double G[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {0.7071067811865476, 0.7071067811865476},
                                       {-0.7071067811865476, 0.7071067811865476}, {0.7071067811865476, -0.7071067811865476},
                                       {-0.7071067811865476, -0.7071067811865476}};
        const double* g = G[hash2((uint32_t)ix, (uint32_t)iz, s) & 7];
        return g[0] * dx + g[1] * dz;
    }
    // gradient noise in about [-1, 1]
    double noise(double x, double z, uint32_t s) const {
        double fx = std::floor(x), fz = std::floor(z);
        int32_t ix = (int32_t)fx, iz = (int32_t)fz;
        double dx = x - fx, dz = z - fz;
        double u = fade(dx), v = fade(dz);
        double n00 = grad(ix, iz, dx, dz, s), n10 = grad(ix + 1, iz, dx - 1, dz, s);
        double n01 = grad(ix, iz + 1, dx, dz - 1, s), n11 = grad(ix + 1, iz + 1, dx - 1, dz - 1, s);
        double a = n00 + u * (n10 - n00), b = n01 + u * (n11 - n01);
        return (a + v * (b - a)) * 1.4142135623730951;
    }
    double fbm(double x, double z, double freq, int octaves, uint32_t s) const {
        double f = freq, a = 1.0, v = 0.0;
        for (int i = 0; i < octaves; ++i) { v += noise(x * f + 0.5, z * f + 0.5, s + (uint32_t)i) * a; f *= 2.0; a *= 0.5; }
        return v;
    }
    static double spline(const double (*pts)[2], int n, double x) {   // gamelogic/worldgen.rs:57-77
        int rhs = -1;
        for (int i = 0; i < n; ++i) if (pts[i][0] > x) { rhs = i; break; }
        if (rhs < 0) return pts[n - 1][1];
        if (rhs == 0) return pts[0][1];
        double f = (x - pts[rhs - 1][0]) / (pts[rhs][0] - pts[rhs - 1][0]);
        return pts[rhs - 1][1] + (pts[rhs][1] - pts[rhs - 1][1]) * f;
    }
    int32_t height_at(int32_t x, int32_t z) const {   // gamelogic/worldgen.rs:191-199
        if (kind == 1) {
            static const float RCONT[6][2] = {{-1.0f, 20.0f}, {0.4f, 50.0f}, {0.6f, 70.0f}, {0.8f, 120.0f}, {0.9f, 190.0f}, {1.0f, 200.0f}};
            static const float RERO[2][2] = {{-1.0f, -10.0f}, {1.0f, 4.0f}};
            const RefNoise continentalness{0.001f, 3, RCONT, 6}, erosion{0.01f, 4, RERO, 2};
            const double h = continentalness.get(perlin, (double)x, (double)z);
            return (int32_t)(h + erosion.get(perlin, (double)x, (double)z));
        }
        static const double CONT[6][2] = {{-1.0, 20.0}, {0.4, 50.0}, {0.6, 70.0}, {0.8, 120.0}, {0.9, 190.0}, {1.0, 200.0}};
        static const double ERO[2][2] = {{-1.0, -10.0}, {1.0, 4.0}};
        // the stand-in noise is stretched (x1.5, +0.25) so that mountains like the reference's appear inside a radius-20 disc
        // (seed + 2: with the reference's seed = 1 the default camera (-24, 80, 174) then hovers over a valley next to
        // a mountain, like the reference's end-to-end image, instead of sitting inside a hill)
        const uint32_t s = seed + 2u;
        double c = fbm((double)x, (double)z, 0.001 * 2.5, 3, s * 7919u) * 1.5 + 0.25;
        double e = fbm((double)x, (double)z, 0.01, 4, s * 104729u + 17u);
        double h = spline(CONT, 6, c) + spline(ERO, 2, e);
        return (int32_t)h;
    }
    // gamelogic/worldgen.rs:294-316
    static BlockId block_at(int32_t y, int32_t height) {
        if (y > height) return 0;
        if (y >= height) return 1;        // GRASS
        if (y >= height - 3) return 2;    // DIRT
        return 3;                         // STONE
    }
};

// The world-level SVO: owns the Esvo of chunks and the world<->SVO mapping (systems::worldsvo::Svo
// without the job system / GPU handle, worldsvo.rs:48-60,90-151,198-218).
enum class SvoFormat : int { Esvo = 0, Csvo = 1 };   // the reference's compile-time SVO_TYPE (Cargo features use-esvo / use-csvo)

class WorldSvo {
public:
    SvoFormat format = SvoFormat::Esvo;
    Esvo<SerializedChunk> esvo;
    Csvo csvo;
    SvoCoordSpace space;
    Terrain terrain;
    std::map<std::tuple<int32_t, int32_t, int32_t>, LeafId> leaf_ids;
    // sparse block edits on top of the terrain function: chunk -> (x,y,z,id)
    std::map<std::tuple<int32_t, int32_t, int32_t>, std::vector<std::array<uint32_t, 4>>> edits;
    bool no_lod = false;

    static uint64_t chunk_uid(ChunkPos p) {
        return ((uint64_t)((uint32_t)p.x & 0x1fffffu) << 42) | ((uint64_t)((uint32_t)p.y & 0x1fffffu) << 21) | (uint64_t)((uint32_t)p.z & 0x1fffffu);
    }

    struct Column { int32_t min_y, max_y; int16_t h[32 * 32]; };   // gamelogic/worldgen.rs:166-176

    Column make_column(int32_t cx, int32_t cz) const {
        Column c; c.min_y = INT32_MAX; c.max_y = INT32_MIN;
        for (int32_t z = 0; z < 32; ++z)
            for (int32_t x = 0; x < 32; ++x) {
                int32_t h = terrain.height_at(cx * 32 + x, cz * 32 + z);
                c.min_y = h < c.min_y ? h : c.min_y; c.max_y = h > c.max_y ? h : c.max_y;
                c.h[z * 32 + x] = (int16_t)h;
            }
        return c;
    }
    static bool column_contains(const Column& c, int32_t chunk_y) { return c.min_y <= (chunk_y + 1) * 32 && c.max_y >= chunk_y * 32; }

    // fills a dense 32^3 array for the chunk from the column heights + edits; returns false if empty
    bool fill_chunk(ChunkPos p, const Column& col, std::vector<BlockId>& blocks) const {
        blocks.assign(32 * 32 * 32, 0);
        bool any = false;
        for (int32_t z = 0; z < 32; ++z)
            for (int32_t x = 0; x < 32; ++x) {
                int32_t h = (int32_t)col.h[z * 32 + x] - p.y * 32;
                int32_t top = h < 31 ? h : 31;
                for (int32_t y = 0; y <= top; ++y) { blocks[(size_t)x + 32 * ((size_t)y + 32 * (size_t)z)] = Terrain::block_at(y, h); any = true; }
            }
        auto it = edits.find({p.x, p.y, p.z});
        if (it != edits.end())
            for (auto& e : it->second) { blocks[(size_t)e[0] + 32 * ((size_t)e[1] + 32 * (size_t)e[2])] = e[3]; any = any || e[3] != 0; }
        return any;
    }

    // systems::worldsvo::Svo::set_chunk + process_serialized_chunks (worldsvo.rs:90-99,198-218)
    bool set_chunk(ChunkPos p, SerializedChunk&& sc) {
        Position sp;
        if (!space.cnv_chunk_pos(p, sp)) return false;
        auto r = esvo.set_leaf(sp, std::move(sc), true);
        leaf_ids[{p.x, p.y, p.z}] = r.first;
        return true;
    }
    bool set_chunk(ChunkPos p, CsvoChunk&& cc) {
        Position sp;
        if (!space.cnv_chunk_pos(p, sp)) return false;
        auto r = csvo.set_leaf(sp, std::move(cc), true);
        leaf_ids[{p.x, p.y, p.z}] = r.first;
        return true;
    }

    // systems::worldsvo::Svo::update (worldsvo.rs:133-137) + on_coord_space_change: the player entered another chunk
    bool set_center(ChunkPos c) {
        if (c.x == space.center.x && c.y == space.center.y && c.z == space.center.z) return false;
        space.center = c;
        if (format == SvoFormat::Csvo) shift_chunks(space, leaf_ids, csvo); else shift_chunks(space, leaf_ids, esvo);
        return true;
    }

    void remove_chunk(ChunkPos p) {   // worldsvo.rs:101-108
        auto it = leaf_ids.find({p.x, p.y, p.z});
        if (it == leaf_ids.end()) return;
        if (format == SvoFormat::Csvo) csvo.remove_leaf(it->second); else esvo.remove_leaf(it->second);
        leaf_ids.erase(it);
    }

    // format dispatch of the WorldSvo<T> trait surface (src/world/hds/common.rs:3-15)
    void serialize() { if (format == SvoFormat::Csvo) csvo.serialize(); else esvo.serialize(); }
    uint8_t depth() const { return format == SvoFormat::Csvo ? csvo.depth() : esvo.depth(); }
    size_t size_in_bytes() const { return format == SvoFormat::Csvo ? csvo.size_in_bytes() : esvo.size_in_bytes(); }
    size_t write_to(uint8_t* dst) const { return format == SvoFormat::Csvo ? csvo.write_to(dst) : esvo.write_to(dst); }
    bool write_changes_to(uint8_t* dst, size_t dst_len, bool reset) {
        return format == SvoFormat::Csvo ? csvo.write_changes_to(dst, dst_len, reset) : esvo.write_changes_to(dst, dst_len, reset);
    }
    RangeBuffer& range_buffer() { return format == SvoFormat::Csvo ? csvo.buffer : esvo.buffer; }
    Range root_range() const { return format == SvoFormat::Csvo ? csvo.root_range() : esvo.root_range(); }
    // bytes in front of the RangeBuffer image inside the GPU buffer: f32 scale + 20-byte preamble (ESVO) / + u32 root offset (CSVO)
    size_t header_bytes() const { return format == SvoFormat::Csvo ? 8 : 24; }

    uint8_t lod_for(ChunkPos p) const { return no_lod ? 5 : calculate_lod(space.center, p); }

    // (re)serialises one terrain chunk (terrain + edits) with the LOD rule and inserts it
    bool regenerate_chunk(ChunkPos p) {
        std::vector<BlockId> blocks;
        Column col = make_column(p.x, p.z);
        if (!fill_chunk(p, col, blocks)) { remove_chunk(p); return false; }
        if (format == SvoFormat::Csvo) return set_chunk(p, CsvoChunk::from_dense(chunk_uid(p), blocks.data(), lod_for(p)));
        return set_chunk(p, SerializedChunk::from_dense(p.x, p.y, p.z, chunk_uid(p), blocks.data(), lod_for(p)));
    }

    // A Load / LodChange event for one chunk as the reference handles it: NopStorage has nothing, the generator takes the chunk only if
    // its column's height span touches it (Generator::is_interested_in, worldgen.rs:270-273), then it is serialized with the LOD rule
    bool load_chunk(ChunkPos p) {
        Column col = make_column(p.x, p.z);
        if (!column_contains(col, p.y) && !edits.count({p.x, p.y, p.z})) { remove_chunk(p); return false; }
        return regenerate_chunk(p);
    }

    // Loads every chunk of the disc of radius `dst` around `center` for world y-chunks [y0, y1]
    // (ChunkLoader::new(radius, 0, 8), gamelogic/world.rs:85), nearest columns first
    // (chunkloader.rs:116-120). Chunk serialisation runs on `threads` workers like the reference's job
    // system (worldsvo.rs:90-99); insertion into the world octree stays on the calling thread.
    size_t generate(int32_t y0, int32_t y1, int threads) {
        struct Job { ChunkPos p; SerializedChunk sc; CsvoChunk cc; bool ok = false; };
        std::vector<std::pair<int64_t, std::pair<int32_t, int32_t>>> cols;
        int32_t r = (int32_t)space.dst;
        for (int32_t dz = -r; dz <= r; ++dz)
            for (int32_t dx = -r; dx <= r; ++dx)
                if (dx * dx + dz * dz <= r * r) cols.push_back({(int64_t)dx * dx + (int64_t)dz * dz, {dx, dz}});
        std::stable_sort(cols.begin(), cols.end(), [](auto& a, auto& b) { return a.first < b.first; });
        std::vector<std::vector<Job>> per_col(cols.size());
        if (threads < 1) threads = 1;
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            std::vector<BlockId> blocks;
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= cols.size()) break;
                int32_t cx = space.center.x + cols[i].second.first, cz = space.center.z + cols[i].second.second;
                Column col = make_column(cx, cz);
                for (int32_t y = y0; y <= y1; ++y) {
                    if (!column_contains(col, y)) continue;
                    ChunkPos p{cx, y, cz};
                    Position sp;
                    if (!space.cnv_chunk_pos(p, sp)) continue;
                    if (!fill_chunk(p, col, blocks)) continue;
                    Job j; j.p = p; j.ok = true;
                    if (format == SvoFormat::Csvo) j.cc = CsvoChunk::from_dense(chunk_uid(p), blocks.data(), lod_for(p));
                    else j.sc = SerializedChunk::from_dense(p.x, p.y, p.z, chunk_uid(p), blocks.data(), lod_for(p));
                    per_col[i].push_back(std::move(j));
                }
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
        size_t loaded = 0;
        for (auto& v : per_col)
            for (auto& j : v)
                if (j.ok && (format == SvoFormat::Csvo ? (j.cc.has_data() && set_chunk(j.p, std::move(j.cc)))
                                                        : (j.sc.has_data() && set_chunk(j.p, std::move(j.sc))))) ++loaded;
        return loaded;
    }
};

}  // namespace vxh
