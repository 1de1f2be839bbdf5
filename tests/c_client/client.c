/* client.c — a plain C99 client of include/voxelrt.h.
 *
 * Two purposes: (1) the header is a C header (gcc -std=c99 -pedantic -Wall -Werror, no C++): what a `bindgen` run or the
 * hand-written extern "C" block of rust/src/graphics/voxelrt_sys.rs binds; (2) the call sequence of graphics::Svo (new -> update ->
 * render -> read_pixels -> raycast, src/graphics/svo.rs:109-255) from C, end to end on a GPU.
 * Without a GPU vx_create must fail with VX_E_CUDA and a message — there is no CPU fallback — and the program exits 3.
 *
 * The scene: one octant record with a single voxel (block id 1) at SVO (0,0,0) of a depth-1 octree, built by hand from the
 * buffer layout of SURVEY Appendix A (f32 scale | 5-word preamble | records).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "voxelrt.h"

static int check(VxCtx* ctx, int rc, const char* what) {
    if (rc != VX_OK) {
        fprintf(stderr, "%s failed (%d): %s\n", what, rc, vx_last_error(ctx));
        return 1;
    }
    return 0;
}

int main(void) {
    VxConfig cfg;
    VxCtx* ctx = NULL;
    VxMaterial mat[2];
    VxRenderParams rp;
    VxRange dirty;
    VxPickerTask task;
    VxPickerResult res;
    VxStats st;
    uint8_t texels[4 * 4 * 4];
    uint8_t* mirror;
    uint8_t* rgba8;
    uint32_t* w32;
    unsigned hit = 0, sky = 0, i;
    const uint32_t W = 64, H = 48;
    int rc;

    memset(&cfg, 0, sizeof cfg);
    cfg.device = 0;
    cfg.svo_capacity_bytes = 1 << 16;
    cfg.max_width = W; cfg.max_height = H; cfg.max_rays = 16;
    rc = vx_create(&cfg, &ctx);
    if (rc != VX_OK) {
        printf("vx_create: rc=%d (%s)\n", rc, vx_last_error(NULL));
        return rc == VX_E_CUDA ? 3 : 1;
    }

    /* Svo::new: one opaque white 4x4 texture, materials 0 and 1 use it on every face, no normal maps */
    memset(texels, 0xff, sizeof texels);
    if (check(ctx, vx_set_textures(ctx, texels, 4, 4, 1, 1), "vx_set_textures")) return 1;
    memset(mat, 0, sizeof mat);
    for (i = 0; i < 2; ++i) {
        mat[i].specular_pow = 1.0f; mat[i].specular_strength = 0.0f;
        mat[i].tex_top = mat[i].tex_side = mat[i].tex_bottom = 0;
        mat[i].tex_top_normal = mat[i].tex_side_normal = mat[i].tex_bottom_normal = -1;
    }
    if (check(ctx, vx_set_materials(ctx, mat, 2), "vx_set_materials")) return 1;

    /* Svo::update: Esvo::write_changes_to writes into the pinned mirror (here: by hand), vx_svo_commit uploads the dirty range */
    mirror = vx_svo_host_mirror(ctx);
    w32 = (uint32_t*)(void*)mirror;
    {
        const float scale = 0.5f;                 /* 2^-depth, depth 1 */
        memcpy(mirror, &scale, 4);
    }
    w32[1] = 0x01u << 8;                          /* preamble: fake parent whose child 0 (the root) has child_mask 0x01 ... */
    w32[2] = w32[3] = w32[4] = 0;
    w32[5] = 5;                                   /* ... and lives at descriptors[5] (absolute pointer) */
    memset(w32 + 6, 0, 48);                       /* the root record (12 words at descriptors[5]) */
    w32[1] |= 0x01u;                              /* its child 0 is a LEAF: leaf_mask bit 0 in the parent's field */
    w32[6 + 4] = 1;                               /* body[0] = block id 1 */
    dirty.offset = 0; dirty.length = 48;
    if (check(ctx, vx_svo_commit(ctx, 0.5f, &dirty, 1, 48, 1), "vx_svo_commit")) return 1;
    if (check(ctx, vx_stats(ctx, &st), "vx_stats")) return 1;

    /* Svo::render from (0.5, 0.5, 3) looking down -z: view = inverse look_to_rh = identity rotation + translation, column-major */
    memset(&rp, 0, sizeof rp);
    rp.view[0] = rp.view[5] = rp.view[10] = rp.view[15] = 1.0f;
    rp.view[12] = 0.5f; rp.view[13] = 0.5f; rp.view[14] = 3.0f;
    rp.fov_y_rad = 1.0f; rp.aspect_ratio = (float)W / (float)H; rp.ambient_intensity = 0.5f;
    rp.light_dir[0] = -0.5f; rp.light_dir[1] = -0.7f; rp.light_dir[2] = -0.5f;
    rp.cam_pos[0] = 0.5f; rp.cam_pos[1] = 0.5f; rp.cam_pos[2] = 3.0f;
    rp.highlight_pos[0] = rp.highlight_pos[1] = rp.highlight_pos[2] = -1000.0f;
    rp.render_shadows = 1; rp.shadow_distance = 100.0f;
    if (check(ctx, vx_render(ctx, &rp, W, H, NULL, NULL), "vx_render")) return 1;
    rgba8 = (uint8_t*)malloc((size_t)W * H * 4);
    if (!rgba8) return 1;
    if (check(ctx, vx_read_frame_rgba8(ctx, rgba8), "vx_read_frame_rgba8")) return 1;
    for (i = 0; i < W * H; ++i) {
        /* sky pixels are the blue-ish gradient of get_sky_color (b > r); the voxel is grey (r == g == b) */
        if (rgba8[4 * i] == rgba8[4 * i + 2]) ++hit; else ++sky;
    }

    /* Svo::raycast: one picker ray straight at the voxel's +z face */
    memset(&task, 0, sizeof task);
    task.max_dst = 10.0f;
    task.pos[0] = 0.5f; task.pos[1] = 0.5f; task.pos[2] = 3.0f;
    task.dir[2] = -1.0f;
    if (check(ctx, vx_raycast(ctx, &task, 1, &res), "vx_raycast")) return 1;
    printf("depth=%u used=%llu hit_pixels=%u sky_pixels=%u picker dst=%.3f normal=(%.0f,%.0f,%.0f)\n", st.depth,
           (unsigned long long)st.used_bytes, hit, sky, (double)res.dst, (double)res.normal[0], (double)res.normal[1], (double)res.normal[2]);
    free(rgba8);
    vx_destroy(ctx);
    if (!hit || !sky) return 1;
    if (res.dst != 2.0f || res.normal[2] != 1.0f) return 1;
    return 0;
}
