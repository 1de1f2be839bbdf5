// esvo.hpp — host-side ESVO serializer: the producer of the byte buffer the CUDA ray caster reads.
//
// Mirrors voxel-rs `world::hds::esvo` + `world::hds::internal` (paths relative to the reference):
//   RangeBuffer        src/world/hds/internal.rs:163-277   first-fit free list, merged dirty list
//   pick_leaf_for_lod  src/world/hds/internal.rs:461-485   representative leaf, order 2,3,6,7,0,1,4,5
//   serialize_octant   src/world/hds/esvo.rs:439-512       12-u32 records, depth-first, relative ptrs
//   SerializedChunk    src/world/hds/esvo.rs:343-413       chunk octree -> records with leaf values
//   Esvo               src/world/hds/esvo.rs:102-340       world octree of chunks, preamble, write_to,
//                                                          write_changes_to
// Record format (esvo.rs:74-101): hdr[k] low16 = child 2k, high16 = child 2k+1, each
// (child_mask<<8)|leaf_mask OF THAT CHILD; body[i] = relative ptr|1<<31, absolute ptr (world-octree
// leaves, +5 preamble words) or the leaf value.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <optional>
#include <unordered_map>
#include <vector>

#include "octree.hpp"

namespace vxh {

using BlockId = uint32_t;

struct Range {
    size_t start, length;
    bool operator==(const Range& o) const { return start == o.start && length == o.length; }
};

// internal.rs:163-277
class RangeBuffer {
public:
    std::vector<uint8_t> bytes;
    std::vector<Range> free_ranges;
    std::vector<Range> updated_ranges;
    std::unordered_map<uint64_t, Range> id_to_range;

    RangeBuffer() = default;
    // with_capacity_in (internal.rs:180-194): zero-filled, one free range over all of it
    explicit RangeBuffer(size_t initial_capacity) : bytes(initial_capacity, 0) {
        if (initial_capacity > 0) free_ranges.push_back(Range{0, initial_capacity});
    }

    size_t insert(uint64_t id, const uint8_t* buf, size_t length) {
        remove(id);
        size_t ptr = bytes.size();
        size_t hit = free_ranges.size();
        for (size_t i = 0; i < free_ranges.size(); ++i)
            if (length <= free_ranges[i].length) { hit = i; break; }
        if (hit < free_ranges.size()) {
            Range& r = free_ranges[hit];
            ptr = r.start;
            if (length < r.length) { r.start += length; r.length -= length; }
            else free_ranges.erase(free_ranges.begin() + (long)hit);
            std::memcpy(bytes.data() + ptr, buf, length);
        } else {
            bytes.insert(bytes.end(), buf, buf + length);
        }
        id_to_range[id] = Range{ptr, length};
        updated_ranges.push_back(Range{ptr, length});
        merge(updated_ranges);
        return ptr;
    }

    void remove(uint64_t id) {
        auto it = id_to_range.find(id);
        if (it == id_to_range.end()) return;
        free_ranges.push_back(it->second);
        id_to_range.erase(it);
        merge(free_ranges);
    }

    void clear() {
        free_ranges.clear();
        free_ranges.push_back(Range{0, bytes.capacity()});
        updated_ranges.clear();
        id_to_range.clear();
    }

    size_t size_in_bytes() const { return bytes.size(); }

    // internal.rs:252-272: sort by start, fold overlapping/adjacent ranges
    static void merge(std::vector<Range>& rs) {
        std::sort(rs.begin(), rs.end(), [](const Range& a, const Range& b) { return a.start < b.start; });
        size_t i = 1;
        while (i < rs.size()) {
            Range rhs = rs[i];
            Range& lhs = rs[i - 1];
            if (rhs.start <= lhs.start + lhs.length) {
                size_t diff = lhs.start + lhs.length - rhs.start;
                if (rhs.length > diff) lhs.length += rhs.length - diff;
                rs.erase(rs.begin() + (long)i);
            } else {
                ++i;
            }
        }
    }
};

// esvo.rs:32-44
struct SerializationResult {
    uint8_t child_mask = 0, leaf_mask = 0, depth = 0;
    bool operator==(const SerializationResult& o) const { return child_mask == o.child_mask && leaf_mask == o.leaf_mask && depth == o.depth; }
};

// internal.rs:461-485
template <typename T>
const T* pick_leaf_for_lod(const Octree<T>& tree, uint32_t octant_id) {
    static const uint8_t ORDER[8] = {2, 3, 6, 7, 0, 1, 4, 5};
    const auto& o = tree.octants[octant_id];
    for (uint8_t i : ORDER)
        if (o.kind[i] == ChildKind::Leaf) return tree.leaf_value(octant_id, i);
    for (uint8_t i : ORDER) {
        if (o.kind[i] != ChildKind::Octant) continue;
        if (const T* r = pick_leaf_for_lod(tree, o.ref[i])) return r;
    }
    return nullptr;
}

// Argument bundle of the per-leaf encoder (esvo.rs:415-426).
template <typename T>
struct ChildEncode {
    uint8_t idx;
    SerializationResult* result;
    uint32_t* rec;  // the 12 words of the parent record
    const T* content;
};

// esvo.rs:439-512
template <typename T, typename Enc>
SerializationResult serialize_octant(const Octree<T>& tree, uint32_t octant_id, std::vector<uint32_t>& dst, uint8_t lod, const Enc& enc) {
    const size_t start = dst.size();
    dst.resize(start + 12, 0u);
    SerializationResult result;
    for (uint8_t idx = 0; idx < 8; ++idx) {
        const auto& o = tree.octants[octant_id];
        if (o.kind[idx] == ChildKind::None) continue;
        result.child_mask |= (uint8_t)(1u << idx);
        if (o.kind[idx] == ChildKind::Leaf || lod == 1) {
            const T* content = tree.leaf_value(octant_id, idx);
            if (!content && o.kind[idx] == ChildKind::Octant) content = pick_leaf_for_lod(tree, o.ref[idx]);
            if (!content) continue;
            enc(ChildEncode<T>{idx, &result, dst.data() + start, content});
        } else {
            uint32_t child_id = o.ref[idx];
            uint8_t child_lod = lod > 0 ? (uint8_t)(lod - 1) : 0;
            uint32_t child_offset = (uint32_t)(dst.size() - start);
            SerializationResult cr = serialize_octant(tree, child_id, dst, child_lod, enc);
            uint32_t mask = ((uint32_t)cr.child_mask << 8) | cr.leaf_mask;
            if (idx & 1) mask <<= 16;
            dst[start + idx / 2] |= mask;
            uint32_t rel = child_offset - 4 - idx;
            dst[start + 4 + idx] = rel | (1u << 31);
            result.depth = std::max<uint8_t>(result.depth, (uint8_t)(cr.depth + 1));
        }
    }
    return result;
}

// esvo.rs:369-383: chunk octree of BlockIds -> records whose leaf slots hold the block id.
inline SerializationResult serialize_block_octree(const Octree<BlockId>& tree, std::vector<uint32_t>& dst, uint8_t lod) {
    if (!tree.root) return SerializationResult{};
    return serialize_octant<BlockId>(tree, *tree.root, dst, lod, [](const ChildEncode<BlockId>& p) {
        p.result->leaf_mask |= (uint8_t)(1u << p.idx);
        p.rec[4 + p.idx] = *p.content;
        p.result->depth = 1;
    });
}

// Fast path with the same output as Chunk::fill_with (chunk.rs:125-131: construct_octants_with(5,..))
// followed by serialize_block_octree: serialises a dense 32^3 block array (index = x + 32*(y + 32*z),
// 0 = air) without materialising the pointer octree. tests/test_host_esvo.py checks equality with the
// generic path.
SerializationResult serialize_dense_chunk(const BlockId* blocks, std::vector<uint32_t>& dst, uint8_t lod);

// esvo.rs:343-413
struct SerializedChunk {
    int32_t cx = 0, cy = 0, cz = 0;
    uint64_t uid = 0;
    uint8_t lod = 0;
    std::vector<uint32_t> buffer;
    bool has_buffer = false;
    SerializationResult result;

    uint64_t unique_id() const { return uid; }
    bool has_data() const { return has_buffer; }
    SerializationResult serialize(std::vector<uint32_t>& dst, uint8_t /*lod*/) {
        if (has_buffer) {
            dst.insert(dst.end(), buffer.begin(), buffer.end());
            buffer.clear(); buffer.shrink_to_fit();
            has_buffer = false;
        }
        return result;
    }
    static SerializedChunk from_octree(int32_t cx, int32_t cy, int32_t cz, uint64_t uid, const Octree<BlockId>& tree, uint8_t lod) {
        SerializedChunk c; c.cx = cx; c.cy = cy; c.cz = cz; c.uid = uid; c.lod = lod;
        c.result = serialize_block_octree(tree, c.buffer, lod);
        c.has_buffer = c.result.depth > 0;
        return c;
    }
    static SerializedChunk from_dense(int32_t cx, int32_t cy, int32_t cz, uint64_t uid, const BlockId* blocks, uint8_t lod) {
        SerializedChunk c; c.cx = cx; c.cy = cy; c.cz = cz; c.uid = uid; c.lod = lod;
        c.result = serialize_dense_chunk(blocks, c.buffer, lod);
        c.has_buffer = c.result.depth > 0;
        return c;
    }
};

// The reference's test fake `impl Serializable for u32` (src/systems/worldsvo.rs:236-245).
struct U32Leaf {
    uint32_t v = 0;
    uint64_t unique_id() const { return v; }
    SerializationResult serialize(std::vector<uint32_t>& dst, uint8_t) { dst.push_back(v); return SerializationResult{1, 1, 1}; }
};

// esvo.rs:102-340
template <typename T>
class Esvo {
public:
    static constexpr uint32_t PREAMBLE_LENGTH_IN_U32 = 5;  // esvo.rs:134
    static constexpr uint64_t ROOT_ID = ~0ull;             // u64::MAX, esvo.rs:270

    struct LeafInfo { size_t buf_offset; SerializationResult serialization; };

    Octree<T> octree;
    RangeBuffer buffer;
    std::unordered_map<uint64_t, LeafInfo> leaf_info;
    std::optional<LeafInfo> root_info;

    void clear() { octree.reset(); changes_.clear(); buffer.clear(); leaf_info.clear(); root_info.reset(); }

    // esvo.rs:203-212
    std::pair<LeafId, std::optional<T>> set_leaf(Position pos, T leaf, bool serialize) {
        uint64_t uid = leaf.unique_id();
        auto r = octree.set_leaf(pos, std::move(leaf));
        if (serialize || !leaf_info.count(uid)) add_change(Change{true, uid, r.first});
        return r;
    }
    std::pair<LeafId, std::optional<T>> move_leaf(LeafId leaf, Position to) { return octree.move_leaf(leaf, to); }
    // esvo.rs:221-228
    std::optional<T> remove_leaf(LeafId leaf) {
        std::optional<T> v = octree.remove_leaf_by_id(leaf);
        if (v) add_change(Change{false, v->unique_id(), LeafId{0, 0}});
        return v;
    }
    const T* get_leaf(Position pos) const { return octree.get_leaf(pos); }

    // esvo.rs:237-276. The reference drains an FxHashSet (unspecified order); this drains in
    // insertion order, which is one of the orders the reference can produce.
    void serialize() {
        if (!octree.root) return;
        std::vector<uint32_t> tmp;
        std::vector<Change> changes;
        changes.swap(changes_);
        for (const Change& ch : changes) {
            if (ch.add) {
                T* content = octree.leaf_value(ch.leaf.parent, ch.leaf.idx);
                if (!content) continue;
                SerializationResult r = content->serialize(tmp, 0);
                if (r.depth > 0) {
                    size_t off = buffer.insert(ch.uid, (const uint8_t*)tmp.data(), tmp.size() * 4);
                    tmp.clear();
                    leaf_info[ch.uid] = LeafInfo{off / 4, r};
                }
            } else {
                buffer.remove(ch.uid);
                leaf_info.erase(ch.uid);
            }
        }
        tmp.clear();
        SerializationResult r = serialize_root(tmp);
        size_t off = buffer.insert(ROOT_ID, (const uint8_t*)tmp.data(), tmp.size() * 4);
        root_info = LeafInfo{off / 4, r};
    }

    uint8_t depth() const { return root_info ? root_info->serialization.depth : 0; }
    size_t size_in_bytes() const { return buffer.size_in_bytes(); }
    // byte range of the world-root octree inside the RangeBuffer (for vx_svo_set_hot_range)
    Range root_range() const { auto it = buffer.id_to_range.find(ROOT_ID); return it == buffer.id_to_range.end() ? Range{0, 0} : it->second; }

    // esvo.rs:291-305
    size_t write_to(uint8_t* dst) const {
        if (!root_info) return 0;
        uint8_t* p = write_preamble(*root_info, dst);
        std::memcpy(p, buffer.bytes.data(), buffer.bytes.size());
        return (size_t)(p - dst) + buffer.bytes.size();
    }

    // esvo.rs:310-339. Returns false where the reference panics (range does not fit dst_len).
    bool write_changes_to(uint8_t* dst, size_t dst_len, bool reset) {
        if (!root_info) return true;
        if (buffer.updated_ranges.empty()) return true;
        uint8_t* p = write_preamble(*root_info, dst);
        for (const Range& r : buffer.updated_ranges) {
            if (!(r.start + r.length < dst_len)) return false;
            std::memcpy(p + r.start, buffer.bytes.data() + r.start, r.length);
        }
        if (reset) buffer.updated_ranges.clear();
        return true;
    }

private:
    struct Change { bool add; uint64_t uid; LeafId leaf; };
    std::vector<Change> changes_;

    void add_change(const Change& c) {
        for (const Change& e : changes_)
            if (e.add == c.add && e.uid == c.uid && (!c.add || e.leaf == c.leaf)) return;  // set semantics
        changes_.push_back(c);
    }

    // esvo.rs:151-175
    SerializationResult serialize_root(std::vector<uint32_t>& dst) const {
        return serialize_octant<T>(octree, *octree.root, dst, 0, [this](const ChildEncode<T>& p) {
            auto it = leaf_info.find(p.content->unique_id());
            if (it == leaf_info.end()) return;
            const LeafInfo& info = it->second;
            uint32_t mask = ((uint32_t)info.serialization.child_mask << 8) | info.serialization.leaf_mask;
            if (p.idx & 1) mask <<= 16;
            p.rec[p.idx / 2] |= mask;
            p.rec[4 + p.idx] = (uint32_t)info.buf_offset + PREAMBLE_LENGTH_IN_U32;
            p.result->depth = std::max<uint8_t>(p.result->depth, (uint8_t)(info.serialization.depth + 1));
        });
    }

    // esvo.rs:179-188
    static uint8_t* write_preamble(const LeafInfo& info, uint8_t* dst) {
        uint32_t w[5] = {(uint32_t)info.serialization.child_mask << 8, 0, 0, 0, (uint32_t)info.buf_offset + PREAMBLE_LENGTH_IN_U32};
        std::memcpy(dst, w, sizeof(w));
        return dst + sizeof(w);
    }
};

}  // namespace vxh
