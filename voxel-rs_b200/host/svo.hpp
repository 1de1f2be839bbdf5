// svo.hpp — host-side mirror of voxel-rs `graphics::Svo` on top of the libvoxelrt C ABI.
//
//   graphics::svo_registry::{Material, VoxelRegistry}  src/graphics/svo_registry.rs:17-165
//   graphics::texture_array::TextureArrayBuilder       src/graphics/texture_array.rs:42-176
//   graphics::svo::{RenderParams, Stats, Svo}          src/graphics/svo.rs:75-255
//   cgmath 0.18 Matrix4::look_to_rh(..).invert()       src/graphics/svo.rs:197 (third-party; restated)
// Same method names, argument meaning and failure behaviour (errors surface as C++ exceptions where
// the reference panics). This layer only calls vx_* — it contains no device code.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/voxelrt.h"
#include "esvo.hpp"
#include "picker.hpp"

namespace vxh {

// ------------------------------------------------------------- registry --

// svo_registry.rs:17-97
struct Material {
    float specular_pow = 0, specular_strength = 0;
    std::optional<std::string> tex_top, tex_side, tex_bottom, tex_top_normal, tex_side_normal, tex_bottom_normal;

    Material& specular(float pow, float strength) { specular_pow = pow; specular_strength = strength; return *this; }
    Material& all_sides(const std::string& n) { return top(n).side(n).bottom(n); }
    Material& top(const std::string& n) { tex_top = n; return *this; }
    Material& side(const std::string& n) { tex_side = n; return *this; }
    Material& bottom(const std::string& n) { tex_bottom = n; return *this; }
    Material& with_normals() {
        if (tex_top) tex_top_normal = *tex_top + "_normal";
        if (tex_side) tex_side_normal = *tex_side + "_normal";
        if (tex_bottom) tex_bottom_normal = *tex_bottom + "_normal";
        return *this;
    }
};

// svo_registry.rs:99-165 + texture_array.rs:42-176. Images are handed over decoded (RGBA8, row 0 =
// top row, as the `image` crate yields them); the v-flip of texture_array.rs:92,126 happens here.
class VoxelRegistry {
public:
    VoxelRegistry& add_texture(const std::string& name, uint32_t w, uint32_t h, const uint8_t* rgba8_top_down) {
        if (tex_index_.count(name)) throw std::runtime_error("name '" + name + "' is already registered");   // texture_array.rs:75-81
        if (textures_.empty()) { width_ = w; height_ = h; }
        if (w != width_ || h != height_) throw std::runtime_error("image does not match base dimensions");   // texture_array.rs:137
        tex_index_[name] = (uint32_t)textures_.size();
        std::vector<uint8_t> flipped((size_t)w * h * 4);
        for (uint32_t y = 0; y < h; ++y)
            std::memcpy(&flipped[(size_t)y * w * 4], rgba8_top_down + (size_t)(h - 1 - y) * w * 4, (size_t)w * 4);
        textures_.push_back(std::move(flipped));
        return *this;
    }
    VoxelRegistry& add_material(BlockId block, const Material& m) { materials_.push_back({block, m}); return *this; }
    void set_mip_levels(uint8_t levels) { mip_levels_ = levels; }   // TextureArrayBuilder::new(6, 4.0), svo_registry.rs:126

    std::optional<uint32_t> lookup(const std::string& name) const {
        auto it = tex_index_.find(name);
        if (it == tex_index_.end()) return std::nullopt;
        return it->second;
    }

    // svo_registry.rs:135-165
    std::vector<VxMaterial> build_material_buffer() const {
        if (materials_.empty()) throw std::runtime_error("registry has no materials");
        BlockId max_id = 0;
        for (auto& e : materials_) max_id = std::max(max_id, e.first);
        VxMaterial zero{};
        std::vector<VxMaterial> out((size_t)max_id + 1, zero);
        auto look = [this](const std::optional<std::string>& n) -> int32_t {
            if (!n) return -1;
            auto id = lookup(*n);
            return id ? (int32_t)*id : 0;
        };
        for (auto& e : materials_) {
            const Material& m = e.second;
            out[e.first] = VxMaterial{m.specular_pow, m.specular_strength, look(m.tex_top), look(m.tex_side), look(m.tex_bottom),
                                      look(m.tex_top_normal), look(m.tex_side_normal), look(m.tex_bottom_normal)};
        }
        return out;
    }

    uint32_t width() const { return width_; }
    uint32_t height() const { return height_; }
    uint32_t layers() const { return (uint32_t)textures_.size(); }
    uint8_t mip_levels() const { return mip_levels_; }
    std::vector<uint8_t> packed_level0() const {
        std::vector<uint8_t> out;
        for (auto& t : textures_) out.insert(out.end(), t.begin(), t.end());
        return out;
    }

private:
    std::vector<std::vector<uint8_t>> textures_;
    std::map<std::string, uint32_t> tex_index_;
    std::vector<std::pair<BlockId, Material>> materials_;
    uint32_t width_ = 0, height_ = 0;
    uint8_t mip_levels_ = 6;
};

// --------------------------------------------------------------- camera --

// cgmath 0.18.0 (Cargo.lock:246-247) Matrix4::look_to_rh followed by SquareMatrix::invert, in f32.
// Column-major 16 floats. Restated from the crate's published source; pinned only through the
// reference's expected PNGs.
inline void look_to_rh_inverted(const float eye[3], const float dir[3], const float up[3], float out[16]) {
    auto norm = [](float v[3]) { float l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); v[0] /= l; v[1] /= l; v[2] /= l; };
    auto cross = [](const float a[3], const float b[3], float o[3]) {
        o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
    };
    auto dot3 = [](const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    float f[3] = {dir[0], dir[1], dir[2]};
    norm(f);
    float s[3];
    cross(f, up, s);
    norm(s);
    float u[3];
    cross(s, f, u);
    // m[c][r]
    float m[4][4] = {{s[0], u[0], -f[0], 0}, {s[1], u[1], -f[1], 0}, {s[2], u[2], -f[2], 0}, {-dot3(eye, s), -dot3(eye, u), dot3(eye, f), 1}};

    // cofactor C(c,r) of element m[c][r]; inverse[c][r] = C(r,c) / det
    auto cofactor = [&](int c0, int r0) {
        float a[3][3];
        int cc = 0;
        for (int c = 0; c < 4; ++c) {
            if (c == c0) continue;
            int rr = 0;
            for (int r = 0; r < 4; ++r) {
                if (r == r0) continue;
                a[cc][rr++] = m[c][r];
            }
            ++cc;
        }
        float d = a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[1][0] * (a[0][1] * a[2][2] - a[2][1] * a[0][2]) +
                  a[2][0] * (a[0][1] * a[1][2] - a[1][1] * a[0][2]);
        return ((c0 + r0) & 1) ? -d : d;
    };
    float det = 0;
    for (int r = 0; r < 4; ++r) det += m[0][r] * cofactor(0, r);
    float inv_det = 1.0f / det;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out[c * 4 + r] = cofactor(r, c) * inv_det;
}

// ------------------------------------------------------------------ Svo --

// graphics::svo::RenderParams, svo.rs:85-106
struct RenderParams {
    float ambient_intensity = 0.3f;
    float light_dir[3] = {-0.57735026f, -0.57735026f, -0.57735026f};
    float cam_pos[3] = {0, 0, 0};
    float cam_fwd[3] = {0, 0, -1};
    float cam_up[3] = {0, 1, 0};
    float fov_y_rad = 1.2566371f;
    float aspect_ratio = 1.0f;
    std::optional<Vec3> selected_voxel;
    bool render_shadows = true;
    float shadow_distance = 500.0f;
};

// graphics::svo::Stats, svo.rs:75-83
struct Stats { size_t used_bytes = 0, capacity_bytes = 0; uint8_t depth = 0; };

// Framebuffer (framebuffer.rs:20-54,97-111): just the size; pixels live in the VxCtx.
struct Framebuffer { uint32_t width, height; };

class Svo {
public:
    // graphics::Svo::new, svo.rs:109-149. size_mb*1e6 bytes of world buffer; max_rays replaces the
    // fixed 100-entry picker buffers (svo.rs:138-139).
    Svo(const VoxelRegistry& registry, size_t size_mb, uint32_t max_width, uint32_t max_height, uint64_t max_rays, int device = 0,
        uint32_t flags = 0) {
        VxConfig cfg{};
        cfg.device = device; cfg.flags = flags;
        cfg.svo_capacity_bytes = (uint64_t)size_mb * 1000 * 1000;
        cfg.max_width = max_width; cfg.max_height = max_height; cfg.max_rays = max_rays;
        check(vx_create(&cfg, &ctx_), "vx_create");
        std::vector<uint8_t> tex = registry.packed_level0();
        check(vx_set_textures(ctx_, tex.data(), registry.width(), registry.height(), registry.layers(), registry.mip_levels()), "vx_set_textures");
        std::vector<VxMaterial> mats = registry.build_material_buffer();
        check(vx_set_materials(ctx_, mats.data(), (uint32_t)mats.size()), "vx_set_materials");
    }
    ~Svo() { if (ctx_) vx_destroy(ctx_); }
    Svo(const Svo&) = delete;
    Svo& operator=(const Svo&) = delete;

    VxCtx* ctx() const { return ctx_; }

    // graphics::Svo::update, svo.rs:171-189: octree_scale = 2^-depth at byte 0, wait for the frame in
    // flight, write_changes_to(mirror + 4, len - 1), refresh stats. The dirty list is read BEFORE
    // write_changes_to resets it (the accessor SURVEY §8b asks the Rust shim to add).
    // W = Esvo<T> or Csvo (the reference's `T: WorldSvo<U>`); the context must have been created with the matching
    // SVO type (VX_FLAG_SVO_CSVO in `flags` = the reference's SvoType argument of graphics::Svo::new, svo.rs:109).
    template <typename W>
    void update(W& svo) {
        VxStats st{};
        vx_stats(ctx_, &st);
        uint8_t* mirror = vx_svo_host_mirror(ctx_);
        std::vector<VxRange> dirty;
        for (const Range& r : svo.buffer.updated_ranges) dirty.push_back(VxRange{r.start, r.length});
        float scale = std::exp2(-(float)svo.depth());
        check(vx_render_wait(ctx_), "vx_render_wait");
        if (!svo.write_changes_to(mirror + 4, (size_t)st.capacity_bytes - 1, true))
            throw std::runtime_error("dst is not large enough");   // esvo.rs:328-331
        Range rr = svo.root_range();
        vx_svo_set_hot_range(ctx_, rr.start, rr.length);
        check(vx_svo_commit(ctx_, scale, dirty.data(), (uint32_t)dirty.size(), svo.size_in_bytes(), svo.depth()), "vx_svo_commit");
        stats_ = Stats{svo.size_in_bytes(), (size_t)st.capacity_bytes, svo.depth()};
    }

    Stats get_stats() const { return stats_; }   // svo.rs:191-193

    // graphics::Svo::render, svo.rs:196-229
    void render(const RenderParams& p, const Framebuffer& target, const VxShard* shard = nullptr) {
        VxRenderParams rp{};
        look_to_rh_inverted(p.cam_pos, p.cam_fwd, p.cam_up, rp.view);
        rp.fov_y_rad = p.fov_y_rad; rp.aspect_ratio = p.aspect_ratio; rp.ambient_intensity = p.ambient_intensity;
        for (int k = 0; k < 3; ++k) { rp.light_dir[k] = p.light_dir[k]; rp.cam_pos[k] = p.cam_pos[k]; }
        const float nan = std::numeric_limits<float>::quiet_NaN();   // svo.rs:211-215
        rp.highlight_pos[0] = p.selected_voxel ? p.selected_voxel->x : nan;
        rp.highlight_pos[1] = p.selected_voxel ? p.selected_voxel->y : nan;
        rp.highlight_pos[2] = p.selected_voxel ? p.selected_voxel->z : nan;
        rp.render_shadows = p.render_shadows ? 1u : 0u;
        rp.shadow_distance = p.shadow_distance;
        check(vx_render(ctx_, &rp, target.width, target.height, shard, nullptr), "vx_render");
    }

    // graphics::Svo::raycast, svo.rs:233-255
    void raycast(const PickerBatch& batch, PickerBatchResult& result) {
        size_t n = batch.serialize_tasks(tasks_);
        results_.resize(n);
        if (n) check(vx_raycast(ctx_, tasks_.data(), n, results_.data()), "vx_raycast");
        batch.deserialize_results(results_.data(), n, result);
    }

private:
    VxCtx* ctx_ = nullptr;
    Stats stats_;
    std::vector<VxPickerTask> tasks_;
    std::vector<VxPickerResult> results_;

    void check(int rc, const char* what) const {
        if (rc != VX_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + vx_last_error(ctx_));
    }
};

}  // namespace vxh
