// esvo.cpp — dense-chunk fast paths of the ESVO and CSVO serializers (see esvo.hpp, csvo.hpp).
#include "csvo.hpp"
#include "esvo.hpp"

namespace vxh {
namespace {

// Occupancy pyramid over a 32^3 block array: occ[k] has (32>>k)^3 cells of edge 2^k voxels, k = 1..5.
struct DenseChunk {
    const BlockId* blocks;
    std::vector<uint8_t> occ[6];

    static size_t cell(uint32_t n, uint32_t x, uint32_t y, uint32_t z) { return (size_t)x + (size_t)n * (y + (size_t)n * z); }
    BlockId at(uint32_t x, uint32_t y, uint32_t z) const { return blocks[cell(32, x, y, z)]; }

    explicit DenseChunk(const BlockId* b) : blocks(b) {
        for (uint32_t k = 1; k <= 5; ++k) {
            uint32_t n = 32u >> k;
            occ[k].assign((size_t)n * n * n, 0);
            for (uint32_t z = 0; z < n; ++z)
                for (uint32_t y = 0; y < n; ++y)
                    for (uint32_t x = 0; x < n; ++x) {
                        uint8_t any = 0;
                        for (uint32_t i = 0; i < 8 && !any; ++i) {
                            uint32_t cx = 2 * x + (i & 1), cy = 2 * y + ((i >> 1) & 1), cz = 2 * z + ((i >> 2) & 1);
                            any = (k == 1) ? (at(cx, cy, cz) != 0) : occ[k - 1][cell(n * 2, cx, cy, cz)];
                        }
                        occ[k][cell(n, x, y, z)] = any;
                    }
        }
    }

    // cell of edge 2^k at cell coordinates (x,y,z) — occupied?
    bool occupied(uint32_t k, uint32_t x, uint32_t y, uint32_t z) const {
        return k == 0 ? at(x, y, z) != 0 : occ[k][cell(32u >> k, x, y, z)] != 0;
    }

    // pick_leaf_for_lod (internal.rs:461-485) on the implicit octree: first non-empty child in the
    // order 2,3,6,7,0,1,4,5, recursively; voxels only exist at the deepest level.
    BlockId pick(uint32_t k, uint32_t x, uint32_t y, uint32_t z) const {
        static const uint8_t ORDER[8] = {2, 3, 6, 7, 0, 1, 4, 5};
        for (uint8_t i : ORDER) {
            uint32_t cx = 2 * x + (i & 1), cy = 2 * y + ((i >> 1) & 1), cz = 2 * z + ((i >> 2) & 1);
            if (!occupied(k - 1, cx, cy, cz)) continue;
            return k == 1 ? at(cx, cy, cz) : pick(k - 1, cx, cy, cz);
        }
        return 0;
    }

    // serialize_octant (esvo.rs:439-512) for the octant of edge 2^k at cell (x,y,z)
    SerializationResult ser(uint32_t k, uint32_t x, uint32_t y, uint32_t z, std::vector<uint32_t>& dst, uint8_t lod) const {
        const size_t start = dst.size();
        dst.resize(start + 12, 0u);
        SerializationResult result;
        for (uint32_t idx = 0; idx < 8; ++idx) {
            uint32_t cx = 2 * x + (idx & 1), cy = 2 * y + ((idx >> 1) & 1), cz = 2 * z + ((idx >> 2) & 1);
            if (!occupied(k - 1, cx, cy, cz)) continue;
            result.child_mask |= (uint8_t)(1u << idx);
            if (k == 1 || lod == 1) {
                BlockId v = (k == 1) ? at(cx, cy, cz) : pick(k - 1, cx, cy, cz);
                result.leaf_mask |= (uint8_t)(1u << idx);
                dst[start + 4 + idx] = v;
                result.depth = 1;
            } else {
                uint8_t child_lod = lod > 0 ? (uint8_t)(lod - 1) : 0;
                uint32_t child_offset = (uint32_t)(dst.size() - start);
                SerializationResult cr = ser(k - 1, cx, cy, cz, dst, child_lod);
                uint32_t mask = ((uint32_t)cr.child_mask << 8) | cr.leaf_mask;
                if (idx & 1) mask <<= 16;
                dst[start + idx / 2] |= mask;
                dst[start + 4 + idx] = (child_offset - 4 - idx) | (1u << 31);
                result.depth = std::max<uint8_t>(result.depth, (uint8_t)(cr.depth + 1));
            }
        }
        return result;
    }
};

// csvo.rs:434-535 for the octant of edge 2^k at cell (x,y,z) with `depth` levels left to serialize.
static std::vector<uint8_t> csvo_ser(const DenseChunk& c, uint32_t k, uint32_t x, uint32_t y, uint32_t z, uint8_t depth, uint16_t material_offset,
                                     std::vector<BlockId>& materials) {
    std::vector<uint8_t> buffer;
    if (depth == 1) {
        uint8_t leaf_mask = 0;
        for (uint32_t idx = 0; idx < 8; ++idx) {
            uint32_t cx = 2 * x + (idx & 1), cy = 2 * y + ((idx >> 1) & 1), cz = 2 * z + ((idx >> 2) & 1);
            if (!c.occupied(k - 1, cx, cy, cz)) continue;
            materials.push_back(k == 1 ? c.at(cx, cy, cz) : c.pick(k - 1, cx, cy, cz));
            leaf_mask |= (uint8_t)(1u << idx);
        }
        buffer.push_back(leaf_mask);
        return buffer;
    }
    std::vector<std::pair<uint8_t, std::vector<uint8_t>>> children;
    for (uint32_t idx = 0; idx < 8; ++idx) {
        uint32_t cx = 2 * x + (idx & 1), cy = 2 * y + ((idx >> 1) & 1), cz = 2 * z + ((idx >> 2) & 1);
        if (!c.occupied(k - 1, cx, cy, cz)) continue;
        children.push_back({(uint8_t)idx, csvo_ser(c, k - 1, cx, cy, cz, (uint8_t)(depth - 1), (uint16_t)materials.size(), materials)});
    }
    if (depth == 2) {
        buffer.push_back(0);
        if (!children.empty()) { buffer.push_back((uint8_t)material_offset); buffer.push_back((uint8_t)(material_offset >> 8)); }
        for (auto& ch : children) { buffer[0] |= (uint8_t)(1u << ch.first); buffer.insert(buffer.end(), ch.second.begin(), ch.second.end()); }
    } else if (depth == 3) {
        buffer.assign(1 + children.size(), 0);
        uint8_t running = 0;
        for (size_t i = 0; i < children.size(); ++i) {
            buffer[0] |= (uint8_t)(1u << children[i].first);
            buffer[1 + i] = running;
            running = (uint8_t)(running + children[i].second.size());
        }
        for (auto& ch : children) buffer.insert(buffer.end(), ch.second.begin(), ch.second.end());
    } else {
        csvo_emit_internal(buffer, children);
    }
    return buffer;
}

}  // namespace

bool csvo_serialize_dense_chunk(const BlockId* blocks, uint8_t depth, std::vector<uint8_t>& nodes, std::vector<BlockId>& materials) {
    DenseChunk c(blocks);
    if (!c.occupied(5, 0, 0, 0)) return false;
    nodes = csvo_ser(c, 5, 0, 0, 0, depth, 0, materials);
    return true;
}

SerializationResult serialize_dense_chunk(const BlockId* blocks, std::vector<uint32_t>& dst, uint8_t lod) {
    DenseChunk c(blocks);
    if (!c.occupied(5, 0, 0, 0)) return SerializationResult{};
    return c.ser(5, 0, 0, 0, dst, lod);
}

}  // namespace vxh
