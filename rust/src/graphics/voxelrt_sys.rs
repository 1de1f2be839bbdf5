//! Raw FFI of libvoxelrt (include/voxelrt.h). One declaration per C entry point `graphics::Svo` needs; layouts are
//! `#[repr(C)]` twins of the C structs (sizes asserted at the bottom, as `src/graphics/svo_picker.rs:310-418` does for the
//! picker records).
//!
//! Not compiled in the libvoxelrt repository (no rustc in that image). `tests/c_client/client.c` there is the same list of
//! calls from plain C and proves the header is a C header; `voxel-rs_b200/__init__.py` binds the same symbols with ctypes.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

use crate::graphics::svo_picker::{PickerResult, PickerTask};
use crate::graphics::svo_registry::MaterialInstance;

pub const VX_OK: c_int = 0;
pub const VX_E_ARG: c_int = -1;
pub const VX_E_CAPACITY: c_int = -2;
pub const VX_E_CUDA: c_int = -3;
pub const VX_E_NCCL: c_int = -4;
pub const VX_E_STATE: c_int = -5;

pub const VX_FLAG_NO_L2_WINDOW: u32 = 1;
pub const VX_FLAG_SVO_CSVO: u32 = 4;

#[repr(C)]
pub struct VxCtx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct VxGroup {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct VxConfig {
    pub device: i32,
    pub flags: u32,
    pub svo_capacity_bytes: u64,
    pub max_width: u32,
    pub max_height: u32,
    pub max_rays: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct VxRange {
    pub offset: u64,
    pub length: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct VxShard {
    pub rank: u32,
    pub world_size: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct VxStats {
    pub used_bytes: u64,
    pub capacity_bytes: u64,
    pub depth: u32,
}

/// = the uniforms of `assets/shaders/world.glsl:12-25`, with `u_view` already inverted on the host (`svo.rs:197`).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct VxRenderParams {
    pub view: [f32; 16],
    pub fov_y_rad: f32,
    pub aspect_ratio: f32,
    pub ambient_intensity: f32,
    pub light_dir: [f32; 3],
    pub cam_pos: [f32; 3],
    pub highlight_pos: [f32; 3],
    pub render_shadows: u32,
    pub shadow_distance: f32,
}

extern "C" {
    pub fn vx_create(cfg: *const VxConfig, out: *mut *mut VxCtx) -> c_int;
    pub fn vx_destroy(ctx: *mut VxCtx);
    pub fn vx_last_error(ctx: *const VxCtx) -> *const c_char;

    // Svo::new (svo.rs:109-149)
    pub fn vx_set_materials(ctx: *mut VxCtx, materials: *const MaterialInstance, count: u32) -> c_int;
    pub fn vx_set_textures(ctx: *mut VxCtx, rgba8: *const u8, width: u32, height: u32, layers: u32, mip_levels: u32) -> c_int;

    // Svo::update (svo.rs:171-189)
    pub fn vx_svo_host_mirror(ctx: *mut VxCtx) -> *mut u8;
    pub fn vx_svo_set_hot_range(ctx: *mut VxCtx, offset: u64, length: u64) -> c_int;
    pub fn vx_svo_commit(ctx: *mut VxCtx, octree_scale: f32, dirty: *const VxRange, n_dirty: u32, used_bytes: u64, depth: u32) -> c_int;
    pub fn vx_stats(ctx: *const VxCtx, out: *mut VxStats) -> c_int;

    // Svo::render (svo.rs:196-229) + Framebuffer::read_pixels (framebuffer.rs:97-105)
    pub fn vx_render(ctx: *mut VxCtx, params: *const VxRenderParams, width: u32, height: u32, shard: *const VxShard, rgba32f_out: *mut f32) -> c_int;
    pub fn vx_render_wait(ctx: *mut VxCtx) -> c_int;
    pub fn vx_read_frame_rgba8(ctx: *mut VxCtx, rgba8_out: *mut u8) -> c_int;
    pub fn vx_read_frame_rgba32f(ctx: *mut VxCtx, rgba32f_out: *mut f32) -> c_int;
    pub fn vx_render_read_rgba8(ctx: *mut VxCtx, params: *const VxRenderParams, width: u32, height: u32, shard: *const VxShard,
                                rgba8_out: *mut u8, bands: u32) -> c_int;
    // the same in two halves; up to two frames in flight (frame k+1 renders while frame k is read back)
    pub fn vx_render_read_rgba8_begin(ctx: *mut VxCtx, params: *const VxRenderParams, width: u32, height: u32, shard: *const VxShard,
                                      rgba8_out: *mut u8, bands: u32) -> c_int;
    pub fn vx_render_read_rgba8_end(ctx: *mut VxCtx) -> c_int;
    pub fn vx_frame_device_ptr(ctx: *mut VxCtx, out_ptr: *mut *mut c_void, width: *mut u32, height: *mut u32) -> c_int;

    // Svo::raycast (svo.rs:233-255)
    pub fn vx_raycast(ctx: *mut VxCtx, tasks: *const PickerTask, n: u64, results: *mut PickerResult) -> c_int;

    // single process, several GPUs (the engine is one process: src/gamelogic/game.rs:102-160)
    pub fn vx_group_create(cfg: *const VxConfig, devices: *const c_int, n_devices: u32, out: *mut *mut VxGroup) -> c_int;
    pub fn vx_group_destroy(group: *mut VxGroup);
    pub fn vx_group_last_error(group: *const VxGroup) -> *const c_char;
    pub fn vx_group_ctx(group: *mut VxGroup, index: u32) -> *mut VxCtx;
    pub fn vx_group_set_materials(group: *mut VxGroup, materials: *const MaterialInstance, count: u32) -> c_int;
    pub fn vx_group_set_textures(group: *mut VxGroup, rgba8: *const u8, width: u32, height: u32, layers: u32, mip_levels: u32) -> c_int;
    pub fn vx_group_svo_commit(group: *mut VxGroup, octree_scale: f32, dirty: *const VxRange, n_dirty: u32, used_bytes: u64, depth: u32) -> c_int;
    pub fn vx_group_render(group: *mut VxGroup, params: *const VxRenderParams, width: u32, height: u32) -> c_int;
    pub fn vx_group_render_read_rgba8(group: *mut VxGroup, params: *const VxRenderParams, width: u32, height: u32, rgba8_out: *mut u8, bands: u32) -> c_int;
    pub fn vx_group_svo_host_mirror(group: *mut VxGroup) -> *mut u8;
    pub fn vx_group_host_frame(group: *mut VxGroup, bytes: u64) -> *mut u8;
    pub fn vx_group_raycast(group: *mut VxGroup, tasks: *const PickerTask, n: u64, results: *mut PickerResult) -> c_int;
}

#[cfg(test)]
mod tests {
    use std::mem::size_of;

    use super::*;

    /// The C side asserts the same numbers (voxel-rs_b200/csrc/voxelrt.cu static_asserts, tests/test_host_picker_abi.py).
    #[test]
    fn abi_sizes() {
        assert_eq!(size_of::<VxConfig>(), 32);
        assert_eq!(size_of::<VxRange>(), 16);
        assert_eq!(size_of::<VxShard>(), 8);
        assert_eq!(size_of::<VxStats>(), 24);
        assert_eq!(size_of::<VxRenderParams>(), 120);
        assert_eq!(size_of::<MaterialInstance>(), 32);
        assert_eq!(size_of::<PickerTask>(), 48);
        assert_eq!(size_of::<PickerResult>(), 48);
    }
}
