// csvo.hpp — host-side CSVO ("clustered SVO") serializer: the reference's DEFAULT world format (Cargo.toml:39-45).
//
// Mirrors voxel-rs `world::hds::csvo` (paths relative to the reference):
//   SerializedChunk::serialize_octant   src/world/hds/csvo.rs:434-535   byte-packed chunk octree, four node kinds
//   SerializedChunk::new                src/world/hds/csvo.rs:403-432   LOD -> depth, material list
//   Csvo::serialize_root                src/world/hds/csvo.rs:70-141    world octree, absolute 32-bit chunk pointers
//   Csvo (WorldSvo impl)                src/world/hds/csvo.rs:143-326   change set, RangeBuffer, write_to / write_changes_to
//
// Chunk record in the RangeBuffer (csvo.rs:213-233): [u8 lod][u32 material_bytes][material_bytes of u32 BlockIds][nodes].
// Node kinds by remaining depth d (svo.csvo.glsl:54-133):
//   d > 3  internal   u16 header, 2 bits per child: 0 = absent, 1/2/3 = child offset stored in 1/2/4 bytes; offsets; children
//   d = 3  pre-leaf   u8 child mask; one u8 offset per present child; children
//   d = 2  leaf       u8 child mask; u16 index of this node's first material in the chunk's material list; one u8 per
//                     present child = that child's voxel mask
//   d = 1  (LOD)      a single u8 voxel mask (a chunk serialized at lod 1)
// Voxel values are not stored in the nodes: the k-th set voxel bit of a leaf node (in child order) is material
// list entry [node's first material + k].
// World octree: internal nodes only; at the level above the chunks every present child has tag 3 and its 4 bytes are
// the chunk record's RangeBuffer offset | 1 << 31 ("crossed boundary", svo.csvo.glsl:103-107).
// GPU buffer = [f32 2^-depth][u32 root offset][RangeBuffer bytes] (svo.csvo.glsl:1-5, csvo.rs:267-283).
#pragma once
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "esvo.hpp"

namespace vxh {

// ilog2(max(x, 1)) / 8 + 1: bytes tag of an offset (csvo.rs:116-118, 511-513)
inline uint16_t csvo_offset_tag(uint32_t offset) {
    uint32_t v = offset < 1 ? 1 : offset, bits = 0;
    while (v >>= 1) ++bits;
    return (uint16_t)(bits / 8 + 1);
}

// Appends an internal node (header + offsets + children) — shared by chunk octants (csvo.rs:500-533) and the world
// octree above the chunk level (csvo.rs:106-133).
inline void csvo_emit_internal(std::vector<uint8_t>& out, const std::vector<std::pair<uint8_t, std::vector<uint8_t>>>& children) {
    const size_t start = out.size();
    out.push_back(0); out.push_back(0);
    uint16_t header = 0;
    uint32_t running = 0;
    std::vector<uint32_t> offsets;
    for (auto& c : children) { offsets.push_back(running); running += (uint32_t)c.second.size(); }
    for (size_t i = 0; i < children.size(); ++i) {
        const uint16_t tag = csvo_offset_tag(offsets[i]);
        header |= (uint16_t)(tag << (children[i].first * 2));
        if (tag == 1) out.push_back((uint8_t)offsets[i]);
        else if (tag == 2) { out.push_back((uint8_t)offsets[i]); out.push_back((uint8_t)(offsets[i] >> 8)); }
        else { for (int k = 0; k < 4; ++k) out.push_back((uint8_t)(offsets[i] >> (8 * k))); }
    }
    for (auto& c : children) out.insert(out.end(), c.second.begin(), c.second.end());
    out[start] = (uint8_t)header; out[start + 1] = (uint8_t)(header >> 8);
}

// csvo.rs:434-535 over the pointer octree.
inline std::vector<uint8_t> csvo_serialize_octant(const Octree<BlockId>& tree, uint32_t octant_id, uint8_t depth, uint16_t material_offset,
                                                  std::vector<BlockId>& materials) {
    const auto& o = tree.octants[octant_id];
    std::vector<uint8_t> buffer;
    if (depth == 1) {
        uint8_t leaf_mask = 0;
        for (uint8_t idx = 0; idx < 8; ++idx) {
            if (o.kind[idx] == ChildKind::None) continue;
            const BlockId* content = tree.leaf_value(octant_id, idx);
            if (!content && o.kind[idx] == ChildKind::Octant) content = pick_leaf_for_lod(tree, o.ref[idx]);
            if (!content) continue;
            materials.push_back(*content);
            leaf_mask |= (uint8_t)(1u << idx);
        }
        buffer.push_back(leaf_mask);
        return buffer;
    }
    std::vector<std::pair<uint8_t, std::vector<uint8_t>>> children;
    for (uint8_t idx = 0; idx < 8; ++idx) {
        if (o.kind[idx] == ChildKind::None) continue;
        if (o.kind[idx] == ChildKind::Leaf) throw std::runtime_error("octree leaves must be at a uniform level");   // csvo.rs:471
        children.push_back({idx, csvo_serialize_octant(tree, o.ref[idx], (uint8_t)(depth - 1), (uint16_t)materials.size(), materials)});
    }
    if (depth == 2) {                                            // leaf nodes, csvo.rs:481-492
        buffer.push_back(0);
        if (!children.empty()) { buffer.push_back((uint8_t)material_offset); buffer.push_back((uint8_t)(material_offset >> 8)); }
        for (auto& c : children) { buffer[0] |= (uint8_t)(1u << c.first); buffer.insert(buffer.end(), c.second.begin(), c.second.end()); }
    } else if (depth == 3) {                                     // pre-leaf nodes, csvo.rs:493-506
        buffer.assign(1 + children.size(), 0);
        uint8_t running = 0;
        for (size_t i = 0; i < children.size(); ++i) {
            buffer[0] |= (uint8_t)(1u << children[i].first);
            buffer[1 + i] = running;
            running = (uint8_t)(running + children[i].second.size());
        }
        for (auto& c : children) buffer.insert(buffer.end(), c.second.begin(), c.second.end());
    } else {                                                     // internal nodes, csvo.rs:507-533
        csvo_emit_internal(buffer, children);
    }
    return buffer;
}

// Same bytes from a dense 32^3 block array (index x + 32*(y + 32*z), 0 = air) without materialising the pointer octree
// (the octree of Chunk::fill_with has exactly the non-empty octants). Defined in csvo.cpp.
bool csvo_serialize_dense_chunk(const BlockId* blocks, uint8_t depth, std::vector<uint8_t>& nodes, std::vector<BlockId>& materials);

// csvo.rs:393-432
struct CsvoChunk {
    uint64_t uid = 0;
    uint8_t lod = 0;                 // what the chunk record's first byte says: chunk.lod, or the storage depth when lod == 0
    std::vector<uint8_t> buffer;
    std::vector<BlockId> materials;
    bool has_buffer = false;

    uint64_t unique_id() const { return uid; }
    bool has_data() const { return has_buffer; }

    static CsvoChunk from_octree(uint64_t uid, const Octree<BlockId>& tree, uint8_t chunk_lod) {
        CsvoChunk c; c.uid = uid;
        uint8_t depth = tree.depth();
        if (chunk_lod != 0 && chunk_lod < depth) depth = chunk_lod;
        c.lod = chunk_lod != 0 ? chunk_lod : tree.depth();
        if (tree.root) { c.buffer = csvo_serialize_octant(tree, *tree.root, depth, 0, c.materials); c.has_buffer = true; }
        return c;
    }
    static CsvoChunk from_dense(uint64_t uid, const BlockId* blocks, uint8_t chunk_lod) {
        CsvoChunk c; c.uid = uid;
        uint8_t depth = 5;
        if (chunk_lod != 0 && chunk_lod < depth) depth = chunk_lod;
        c.lod = chunk_lod != 0 ? chunk_lod : 5;
        c.has_buffer = csvo_serialize_dense_chunk(blocks, depth, c.buffer, c.materials);
        return c;
    }
};

// csvo.rs:28-326
class Csvo {
public:
    static constexpr uint64_t ROOT_ID = ~0ull;

    Octree<CsvoChunk> octree;
    RangeBuffer buffer;
    std::unordered_map<uint64_t, size_t> leaf_info;   // uid -> RangeBuffer offset of the chunk record
    std::optional<size_t> root_info;
    uint8_t child_depth = 0;

    void clear() { octree.reset(); changes_.clear(); child_depth = 0; buffer.clear(); leaf_info.clear(); root_info.reset(); }

    std::pair<LeafId, std::optional<CsvoChunk>> set_leaf(Position pos, CsvoChunk leaf, bool serialize) {   // csvo.rs:158-167
        const uint64_t uid = leaf.uid;
        auto r = octree.set_leaf(pos, std::move(leaf));
        if (serialize || !leaf_info.count(uid)) add_change(Change{true, uid, r.first});
        return r;
    }
    std::pair<LeafId, std::optional<CsvoChunk>> move_leaf(LeafId leaf, Position to) { return octree.move_leaf(leaf, to); }
    const CsvoChunk* get_leaf(Position pos) const { return octree.get_leaf(pos); }
    std::optional<CsvoChunk> remove_leaf(LeafId leaf) {                                                     // csvo.rs:176-183
        std::optional<CsvoChunk> v = octree.remove_leaf_by_id(leaf);
        if (v) add_change(Change{false, v->uid, LeafId{0, 0}});
        return v;
    }

    // csvo.rs:192-251. Changes are drained in insertion order (the reference drains an FxHashSet: unspecified order).
    void serialize() {
        if (!octree.root) return;
        std::vector<Change> changes;
        changes.swap(changes_);
        std::vector<uint8_t> tmp;
        for (const Change& ch : changes) {
            if (ch.add) {
                CsvoChunk* content = octree.leaf_value(ch.leaf.parent, ch.leaf.idx);
                if (!content) continue;
                child_depth = std::max(child_depth, content->lod);
                if (content->has_buffer) {
                    const uint32_t material_bytes = (uint32_t)(content->materials.size() * sizeof(BlockId));
                    tmp.clear();
                    tmp.push_back(content->lod);
                    for (int k = 0; k < 4; ++k) tmp.push_back((uint8_t)(material_bytes >> (8 * k)));
                    for (BlockId m : content->materials) for (int k = 0; k < 4; ++k) tmp.push_back((uint8_t)(m >> (8 * k)));
                    tmp.insert(tmp.end(), content->buffer.begin(), content->buffer.end());
                    leaf_info[ch.uid] = buffer.insert(ch.uid, tmp.data(), tmp.size());
                    content->buffer.clear(); content->buffer.shrink_to_fit();
                    content->materials.clear(); content->materials.shrink_to_fit();
                    content->has_buffer = false;
                }
            } else {
                buffer.remove(ch.uid);
                leaf_info.erase(ch.uid);
            }
        }
        std::vector<uint8_t> root = serialize_root(*octree.root, octree.depth());
        root_info = buffer.insert(ROOT_ID, root.data(), root.size());
    }

    uint8_t depth() const { return (uint8_t)(octree.depth() + child_depth); }   // csvo.rs:253-255
    size_t size_in_bytes() const { return buffer.size_in_bytes(); }
    Range root_range() const { auto it = buffer.id_to_range.find(ROOT_ID); return it == buffer.id_to_range.end() ? Range{0, 0} : it->second; }

    // csvo.rs:263-283: [u32 root offset][bytes]
    size_t write_to(uint8_t* dst) const {
        if (!root_info) return 0;
        const uint32_t off = (uint32_t)*root_info;
        std::memcpy(dst, &off, 4);
        std::memcpy(dst + 4, buffer.bytes.data(), buffer.bytes.size());
        return 4 + buffer.bytes.size();
    }
    // csvo.rs:288-325. Returns false where the reference panics.
    bool write_changes_to(uint8_t* dst, size_t dst_len, bool reset) {
        if (!root_info) return true;
        if (buffer.updated_ranges.empty()) return true;
        const uint32_t off = (uint32_t)*root_info;
        std::memcpy(dst, &off, 4);
        for (const Range& r : buffer.updated_ranges) {
            if (!(r.start + r.length < dst_len)) return false;
            std::memcpy(dst + 4 + r.start, buffer.bytes.data() + r.start, r.length);
        }
        if (reset) buffer.updated_ranges.clear();
        return true;
    }

private:
    struct Change { bool add; uint64_t uid; LeafId leaf; };
    std::vector<Change> changes_;
    void add_change(const Change& c) {
        for (const Change& e : changes_)
            if (e.add == c.add && e.uid == c.uid && (!c.add || e.leaf == c.leaf)) return;
        changes_.push_back(c);
    }

    // csvo.rs:70-141
    std::vector<uint8_t> serialize_root(uint32_t octant_id, uint8_t depth) const {
        const auto& o = octree.octants[octant_id];
        std::vector<std::pair<uint8_t, std::vector<uint8_t>>> children;
        for (uint8_t idx = 0; idx < 8; ++idx) {
            if (o.kind[idx] == ChildKind::None) continue;
            if (depth == 1) {
                if (const CsvoChunk* content = octree.leaf_value(octant_id, idx)) {
                    auto it = leaf_info.find(content->uid);
                    if (it != leaf_info.end()) {
                        const uint32_t pointer = (uint32_t)it->second | (1u << 31);
                        children.push_back({idx, {(uint8_t)pointer, (uint8_t)(pointer >> 8), (uint8_t)(pointer >> 16), (uint8_t)(pointer >> 24)}});
                    }
                }
                continue;
            }
            if (o.kind[idx] == ChildKind::Leaf) throw std::runtime_error("octree leaves must be at a uniform level");   // csvo.rs:89
            children.push_back({idx, serialize_root(o.ref[idx], (uint8_t)(depth - 1))});
        }
        std::vector<uint8_t> out;
        if (depth == 1) {
            out.push_back(0); out.push_back(0);
            uint16_t header = 0;
            for (auto& c : children) { header |= (uint16_t)(3u << (c.first * 2)); out.insert(out.end(), c.second.begin(), c.second.end()); }
            out[0] = (uint8_t)header; out[1] = (uint8_t)(header >> 8);
        } else {
            csvo_emit_internal(out, children);
        }
        return out;
    }
};

}  // namespace vxh
