// picker.hpp — ray / AABB probe batches above the C ABI.
//
// Mirrors voxel-rs `graphics::svo_picker` (src/graphics/svo_picker.rs): PickerBatch::serialize_tasks
// (:63-80), deserialize_results (:84-104), Aabb::generate_picker_tasks (:183-243: every lattice point
// of the box casts one axis ray per axis on which it lies on the box surface, max_dst fixed at 10)
// and Aabb::parse_picker_results (:245-299: per-axis/per-sign minimum distance, -1 = no hit).
// Unlike the reference there is no 100-task cap (svo_picker.rs:5); batches grow as needed.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/voxelrt.h"

namespace vxh {

struct Vec3 { float x, y, z; };

struct Ray { Vec3 pos, dir; float max_dst; };
struct RayResult {
    float dst; bool inside_voxel; Vec3 pos, normal;
    bool did_hit() const { return dst != -1.0f; }   // svo_picker.rs:151-153
};
struct Aabb { Vec3 pos, offset, extents; };
struct AabbResult { Vec3 neg{-1, -1, -1}, pos{-1, -1, -1}; };   // svo_picker.rs:169-176

struct PickerBatchResult {
    std::vector<RayResult> rays;
    std::vector<AabbResult> aabbs;
    void reset() { rays.clear(); aabbs.clear(); }
};

class PickerBatch {
public:
    std::vector<Ray> rays;
    std::vector<Aabb> aabbs;

    void reset() { rays.clear(); aabbs.clear(); }
    void add_ray(Vec3 pos, Vec3 dir, float max_dst) { rays.push_back(Ray{pos, dir, max_dst}); }
    void add_aabb(const Aabb& a) { aabbs.push_back(a); }

    // svo_picker.rs:63-80
    size_t serialize_tasks(std::vector<VxPickerTask>& tasks) const {
        tasks.clear();
        for (const Ray& r : rays) tasks.push_back(make_task(r.max_dst, r.pos, r.dir));
        for (const Aabb& a : aabbs) generate_aabb_tasks(a, tasks);
        return tasks.size();
    }

    // svo_picker.rs:84-104
    void deserialize_results(const VxPickerResult* results, size_t n, PickerBatchResult& dst) const {
        size_t off = 0;
        for (size_t i = 0; i < rays.size() && off < n; ++i, ++off) {
            const VxPickerResult& r = results[off];
            dst.rays.push_back(RayResult{r.dst, (r.inside_voxel & 0xffu) != 0, Vec3{r.pos[0], r.pos[1], r.pos[2]},
                                         Vec3{r.normal[0], r.normal[1], r.normal[2]}});
        }
        for (const Aabb& a : aabbs) {
            AabbResult res;
            off += parse_aabb_results(a, results + off, res);
            dst.aabbs.push_back(res);
        }
    }

private:
    static VxPickerTask make_task(float max_dst, Vec3 p, Vec3 d) {
        VxPickerTask t{};
        t.max_dst = max_dst;
        t.pos[0] = p.x; t.pos[1] = p.y; t.pos[2] = p.z;
        t.dir[0] = d.x; t.dir[1] = d.y; t.dir[2] = d.z;
        return t;
    }

    // Visits every (lattice point, axis) pair of the box surface in the reference's order and calls
    // f(x, y, z, axis, at_min_side).
    template <typename F>
    static void for_each_probe(const Aabb& a, const int n[3], F&& f) {
        int p[3];
        for (p[0] = 0; p[0] <= n[0]; ++p[0])
            for (p[1] = 0; p[1] <= n[1]; ++p[1])
                for (p[2] = 0; p[2] <= n[2]; ++p[2])
                    for (int axis = 0; axis < 3; ++axis) {
                        int v = p[axis];
                        if (v != 0 && v != n[axis]) continue;
                        f(p, axis, v == 0);
                    }
    }

    static void blocks_per_axis(const Aabb& a, int n[3]) {
        n[0] = (int)std::ceil(a.extents.x); n[1] = (int)std::ceil(a.extents.y); n[2] = (int)std::ceil(a.extents.z);
    }

    // svo_picker.rs:183-243
    static void generate_aabb_tasks(const Aabb& a, std::vector<VxPickerTask>& tasks) {
        int n[3];
        blocks_per_axis(a, n);
        const float step[3] = {a.extents.x / (float)n[0], a.extents.y / (float)n[1], a.extents.z / (float)n[2]};
        for_each_probe(a, n, [&](const int p[3], int axis, bool at_min) {
            float d[3] = {0, 0, 0};
            d[axis] = at_min ? -1.0f : 1.0f;
            Vec3 point{(float)p[0] * step[0], (float)p[1] * step[1], (float)p[2] * step[2]};
            Vec3 pos{(a.pos.x + a.offset.x) + point.x, (a.pos.y + a.offset.y) + point.y, (a.pos.z + a.offset.z) + point.z};
            tasks.push_back(make_task(10.0f, pos, Vec3{d[0], d[1], d[2]}));
        });
    }

    // svo_picker.rs:245-299; returns the number of results consumed
    static size_t parse_aabb_results(const Aabb& a, const VxPickerResult* data, AabbResult& out) {
        int n[3];
        blocks_per_axis(a, n);
        float* refs[6] = {&out.pos.x, &out.neg.x, &out.pos.y, &out.neg.y, &out.pos.z, &out.neg.z};
        size_t consumed = 0;
        for_each_probe(a, n, [&](const int*, int axis, bool at_min) {
            float dst = data[consumed++].dst;
            if (dst == -1.0f) return;
            float* r = refs[axis * 2 + (at_min ? 1 : 0)];
            *r = (*r == -1.0f) ? dst : std::fmin(*r, dst);
        });
        return consumed;
    }
};

}  // namespace vxh
