// chunkloader.hpp — which chunks are loaded around the player, and at which level of detail.
//
// Mirrors systems::chunkloader::ChunkLoader (src/systems/chunkloader.rs:8-143): a disc of `radius` chunk columns around the chunk
// the position is in, world heights [start_y, end_y) clipped to the radius, LOD by 2-D distance (:127-134). update() reports what
// changed since the last call as Load / Unload / LodChange events sorted by distance to the new position (:56-124). The reference
// keeps the loaded set in an FxHashMap and therefore emits Unload events in an unspecified order; this keeps a std::map (key order).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <optional>
#include <tuple>
#include <vector>

#include "world.hpp"

namespace vxh {

struct ChunkEvent {
    enum Kind : uint8_t { Load = 0, Unload = 1, LodChange = 2 };   // declaration order of the reference's enum (its derived Ord)
    Kind kind;
    ChunkPos pos;
    uint8_t lod;   // 0 for Unload
};

class ChunkLoader {
public:
    ChunkLoader(uint32_t radius, int32_t start_y, int32_t end_y) : radius_(radius), start_y_(start_y), end_y_(end_y) {}

    uint32_t radius() const { return radius_; }
    void set_radius(uint32_t r) { radius_ = r; last_.reset(); }   // :49-53: recheck everything on the next update
    bool is_loaded(ChunkPos p) const { return loaded_.count(key(p)) != 0; }
    void add_loaded_chunk(ChunkPos p, uint8_t lod) { loaded_[key(p)] = lod; }
    size_t loaded_count() const { return loaded_.size(); }

    // :56-124. `pos` is a world-space block position; `as i32` truncates toward zero before the >> 5 (chunk.rs:150-152).
    std::vector<ChunkEvent> update(float px, float py, float pz) {
        std::vector<ChunkEvent> events;
        const ChunkPos cur{(int32_t)px >> 5, (int32_t)py >> 5, (int32_t)pz >> 5};
        if (last_ && last_->x == cur.x && last_->y == cur.y && last_->z == cur.z) return events;
        // The reference never assigns last_pos (it stays None, :13,40,52,61), so every call re-evaluates the disc; a call that changes
        // nothing returns no events either way. Kept like that: add_loaded_chunk() between two calls is then seen, as in the reference.
        const int32_t r = (int32_t)radius_;
        for (int32_t dx = -r; dx <= r; ++dx)
            for (int32_t dz = -r; dz <= r; ++dz) {
                if (dx * dx + dz * dz > r * r) continue;
                ChunkPos p{cur.x + dx, 0, cur.z + dz};
                const uint8_t lod = calculate_lod(cur, p);
                for (int32_t y = start_y_; y < end_y_; ++y) {
                    const int32_t dy = y - cur.y;
                    if (dy < -r || dy > r) continue;
                    p.y = y;
                    auto it = loaded_.find(key(p));
                    if (it != loaded_.end()) {
                        if (it->second != lod) { events.push_back(ChunkEvent{ChunkEvent::LodChange, p, lod}); it->second = lod; }
                    } else {
                        events.push_back(ChunkEvent{ChunkEvent::Load, p, lod});
                        loaded_[key(p)] = lod;
                    }
                }
            }
        std::vector<std::tuple<int32_t, int32_t, int32_t>> gone;
        for (auto& kv : loaded_) {
            const int32_t dx = std::abs(std::get<0>(kv.first) - cur.x), dy = std::abs(std::get<1>(kv.first) - cur.y),
                          dz = std::abs(std::get<2>(kv.first) - cur.z);
            if (dy > r || dx * dx + dz * dz > r * r) {
                gone.push_back(kv.first);
                events.push_back(ChunkEvent{ChunkEvent::Unload, ChunkPos{std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first)}, 0});
            }
        }
        for (auto& k : gone) loaded_.erase(k);
        auto dst_sq = [&](const ChunkPos& p) {   // chunk.rs:155-160
            const float dx = (float)(cur.x - p.x), dy = (float)(cur.y - p.y), dz = (float)(cur.z - p.z);
            return std::fma(dz, dz, std::fma(dx, dx, dy * dy));
        };
        std::stable_sort(events.begin(), events.end(), [&](const ChunkEvent& a, const ChunkEvent& b) { return dst_sq(a.pos) < dst_sq(b.pos); });
        return events;
    }

private:
    static std::tuple<int32_t, int32_t, int32_t> key(ChunkPos p) { return {p.x, p.y, p.z}; }
    uint32_t radius_;
    int32_t start_y_, end_y_;
    std::optional<ChunkPos> last_;
    std::map<std::tuple<int32_t, int32_t, int32_t>, uint8_t> loaded_;
};

}  // namespace vxh
