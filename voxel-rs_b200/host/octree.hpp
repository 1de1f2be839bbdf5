// octree.hpp — host-side pointer octree, the structure the ESVO serializer walks.
//
// Mirrors the observable behaviour of voxel-rs `world::hds::octree::Octree<T>`
// (src/world/hds/octree.rs): child index = x + 2y + 4z (:21-23), set_leaf grows the tree with
// expand_to(required_depth) (:101-122), expand() wraps the root as child 0 of a new root and, on an
// EMPTY tree, leaves a chain of empty octants under child 0 (:311-324) which the
// serializer emits and the ray caster traverses, compact() prunes empty octants depth-first
// (:341-376), construct_octants_with builds bottom-up and skips empty branches (:127-172).
// Storage is index-based (octant table + leaf pool + free lists) rather than enum children.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <optional>
#include <utility>
#include <vector>

namespace vxh {

struct Position {
    uint32_t x, y, z;
    uint8_t child_index() const { return (uint8_t)(x + y * 2 + z * 4); }
    // octree.rs:25-28
    uint8_t required_depth() const {
        uint32_t m = std::max<uint32_t>(1, std::max(x, std::max(y, z)));
        return (uint8_t)((int)std::floor(std::log2((float)m)) + 1);
    }
};

struct LeafId {
    uint32_t parent;
    uint8_t idx;
    bool operator==(const LeafId& o) const { return parent == o.parent && idx == o.idx; }
};

enum class ChildKind : uint8_t { None = 0, Octant = 1, Leaf = 2 };

template <typename T>
class Octree {
public:
    struct Octant {
        int64_t parent = -1;
        uint8_t children_count = 0;
        ChildKind kind[8] = {};
        uint32_t ref[8] = {};  // octant id or leaf-pool slot
    };

    std::optional<uint32_t> root;
    std::vector<Octant> octants;

    uint8_t depth() const { return depth_; }
    const std::vector<uint32_t>& free_list() const { return free_octants_; }   // octree.rs `free_list` (ids of recycled octants)

    void reset() {
        root.reset();
        octants.clear();
        free_octants_.clear();
        leaves_.clear();
        free_leaves_.clear();
        depth_ = 0;
    }

    const T* leaf_value(uint32_t octant, uint8_t idx) const {
        const Octant& o = octants[octant];
        return o.kind[idx] == ChildKind::Leaf ? &leaves_[o.ref[idx]] : nullptr;
    }
    T* leaf_value(uint32_t octant, uint8_t idx) {
        Octant& o = octants[octant];
        return o.kind[idx] == ChildKind::Leaf ? &leaves_[o.ref[idx]] : nullptr;
    }

    // octree.rs:101-122
    std::pair<LeafId, std::optional<T>> set_leaf(Position pos, T leaf) {
        expand_to(pos.required_depth());
        uint32_t it = *root;
        uint32_t size = 1u << depth_;
        while (size >= 1) {
            size /= 2;
            uint8_t idx = Position{pos.x / size, pos.y / size, pos.z / size}.child_index();
            pos.x %= size; pos.y %= size; pos.z %= size;
            if (size == 1) {
                std::optional<T> prev = clear_child(it, idx);
                put_leaf(it, idx, std::move(leaf));
                return {LeafId{it, idx}, std::move(prev)};
            }
            it = step_into_or_create(it, idx);
        }
        return {LeafId{0, 0}, std::nullopt};  // unreachable
    }

    // octree.rs:177-218
    std::pair<LeafId, std::optional<T>> move_leaf(LeafId leaf_id, Position to) {
        expand_to(to.required_depth());
        uint32_t it = *root;
        uint32_t size = 1u << depth_;
        while (size >= 1) {
            size /= 2;
            uint8_t idx = Position{to.x / size, to.y / size, to.z / size}.child_index();
            to.x %= size; to.y %= size; to.z %= size;
            if (size == 1) {
                if (it == leaf_id.parent && idx == leaf_id.idx) return {leaf_id, std::nullopt};
                std::optional<T> old_leaf = clear_child(it, idx);
                std::optional<T> moving = clear_child(leaf_id.parent, leaf_id.idx);
                if (moving) put_leaf(it, idx, std::move(*moving));
                return {LeafId{it, idx}, std::move(old_leaf)};
            }
            it = step_into_or_create(it, idx);
        }
        return {LeafId{0, 0}, std::nullopt};  // unreachable
    }

    // octree.rs:239-267
    std::pair<std::optional<T>, std::optional<LeafId>> remove_leaf(Position pos) {
        if (pos.required_depth() > depth_ || !root) return {std::nullopt, std::nullopt};
        uint32_t it = *root;
        uint32_t size = 1u << depth_;
        while (size >= 1) {
            size /= 2;
            if (size == 0) break;
            uint8_t idx = Position{pos.x / size, pos.y / size, pos.z / size}.child_index();
            pos.x %= size; pos.y %= size; pos.z %= size;
            const Octant& o = octants[it];
            if (o.kind[idx] == ChildKind::None) break;
            if (o.kind[idx] == ChildKind::Octant) { it = o.ref[idx]; continue; }
            std::optional<T> v = clear_child(it, idx);
            return {std::move(v), LeafId{it, idx}};
        }
        return {std::nullopt, std::nullopt};
    }

    // octree.rs:270-281
    std::optional<T> remove_leaf_by_id(LeafId id) {
        if (id.parent >= octants.size() || octants[id.parent].kind[id.idx] != ChildKind::Leaf) return std::nullopt;
        return clear_child(id.parent, id.idx);
    }

    // octree.rs:284-307
    const T* get_leaf(Position pos) const {
        if (!root) return nullptr;
        uint32_t it = *root;
        uint32_t size = 1u << depth_;
        while (size > 1) {
            size /= 2;
            uint8_t idx = Position{pos.x / size, pos.y / size, pos.z / size}.child_index();
            pos.x %= size; pos.y %= size; pos.z %= size;
            const Octant& o = octants[it];
            if (o.kind[idx] == ChildKind::None) return nullptr;
            if (o.kind[idx] == ChildKind::Octant) { it = o.ref[idx]; continue; }
            return &leaves_[o.ref[idx]];
        }
        return nullptr;
    }

    // octree.rs:311-324
    void expand(uint8_t by) {
        for (uint8_t i = 0; i < by; ++i) {
            uint32_t new_root = new_octant(-1);
            if (root) {
                octants[*root].parent = new_root;
                link_octant(new_root, 0, *root);
            }
            root = new_root;
        }
        depth_ = (uint8_t)(depth_ + by);
    }

    // octree.rs:328-336
    void expand_to(uint8_t to) {
        if (depth_ > to) return;
        uint8_t diff = (uint8_t)(to - depth_);
        if (diff > 0) expand(diff);
    }

    // octree.rs:341-353
    void compact() {
        if (!root) return;
        compact_octant(*root);
        if (octants[*root].children_count != 0) return;
        reset();
    }

    // octree.rs:127-172
    void construct_octants_with(uint8_t depth, const std::function<std::optional<T>(Position)>& f) {
        reset();
        int64_t r = construct_impl(1u << depth, Position{0, 0, 0}, f);
        if (r >= 0) { root = (uint32_t)r; depth_ = depth; }
    }

private:
    std::vector<uint32_t> free_octants_;
    std::vector<T> leaves_;
    std::vector<uint32_t> free_leaves_;
    uint8_t depth_ = 0;

    uint32_t new_octant(int64_t parent) {
        if (!free_octants_.empty()) {
            uint32_t id = free_octants_.back();
            free_octants_.pop_back();
            octants[id].parent = parent;
            return id;
        }
        octants.push_back(Octant{});
        octants.back().parent = parent;
        return (uint32_t)octants.size() - 1;
    }

    void link_octant(uint32_t parent, uint8_t idx, uint32_t child) {
        Octant& o = octants[parent];
        if (o.kind[idx] == ChildKind::None) o.children_count++;
        o.kind[idx] = ChildKind::Octant;
        o.ref[idx] = child;
    }

    void put_leaf(uint32_t parent, uint8_t idx, T value) {
        uint32_t slot;
        if (!free_leaves_.empty()) { slot = free_leaves_.back(); free_leaves_.pop_back(); leaves_[slot] = std::move(value); }
        else { slot = (uint32_t)leaves_.size(); leaves_.push_back(std::move(value)); }
        Octant& o = octants[parent];
        if (o.kind[idx] == ChildKind::None) o.children_count++;
        o.kind[idx] = ChildKind::Leaf;
        o.ref[idx] = slot;
    }

    // sets the child to None; returns the leaf value if it held one (octants are just unlinked)
    std::optional<T> clear_child(uint32_t parent, uint8_t idx) {
        Octant& o = octants[parent];
        std::optional<T> out;
        if (o.kind[idx] == ChildKind::Leaf) {
            out = std::move(leaves_[o.ref[idx]]);
            free_leaves_.push_back(o.ref[idx]);
        }
        if (o.kind[idx] != ChildKind::None) o.children_count--;
        o.kind[idx] = ChildKind::None;
        o.ref[idx] = 0;
        return out;
    }

    // octree.rs:220-234
    uint32_t step_into_or_create(uint32_t it, uint8_t idx) {
        if (octants[it].kind[idx] == ChildKind::Octant) return octants[it].ref[idx];
        uint32_t next = new_octant(it);
        link_octant(it, idx, next);
        return next;
    }

    // octree.rs:355-376, 395-412
    void compact_octant(uint32_t id) {
        for (uint8_t i = 0; i < 8; ++i) {
            if (octants[id].kind[i] != ChildKind::Octant) continue;
            uint32_t child = octants[id].ref[i];
            compact_octant(child);
            if (octants[child].children_count == 0) {
                octants[child] = Octant{};
                free_octants_.push_back(child);
                clear_child(id, i);
            }
        }
    }

    int64_t construct_impl(uint32_t size, Position pos, const std::function<std::optional<T>(Position)>& f) {
        size /= 2;
        int64_t parent = -1;
        for (uint8_t i = 0; i < 8; ++i) {
            Position cp{pos.x + size * (i & 1u), pos.y + size * ((i >> 1) & 1u), pos.z + size * ((i >> 2) & 1u)};
            if (size > 1) {
                int64_t child = construct_impl(size, cp, f);
                if (child < 0) continue;
                if (parent < 0) parent = new_octant(-1);
                link_octant((uint32_t)parent, i, (uint32_t)child);
                octants[(uint32_t)child].parent = parent;
                continue;
            }
            std::optional<T> v = f(cp);
            if (v) {
                if (parent < 0) parent = new_octant(-1);
                put_leaf((uint32_t)parent, i, std::move(*v));
            }
        }
        return parent;
    }
};

}  // namespace vxh
