//! graphics::Svo on libvoxelrt — drop-in for `src/graphics/svo.rs` of tim-oster/voxel-rs.
//!
//! Same public surface, same argument meaning: `Svo::new`, `reload_resources`, `update`, `get_stats`, `render`, `raycast`,
//! `SvoType`, `Stats`, `RenderParams`. What changes is what sits behind them: the two GLSL programs, the persistently mapped GL
//! buffers and the two fences become one `VxCtx` (hand-written CUDA for sm_100a behind `include/voxelrt.h`).
//!
//!   reference (svo.rs)                                    here
//!   Svo::new :109-149    shaders, SSBOs, MappedBuffer      vx_create + vx_set_textures + vx_set_materials
//!   update   :171-189    fence wait + write into mapping   write_changes_to into the pinned mirror + vx_svo_commit (async H2D)
//!   render   :196-229    uniforms + DispatchCompute        VxRenderParams + vx_render (+ vx_read_frame_* / CUDA-GL interop)
//!   raycast  :233-255    ≤100 tasks, blocking fence        vx_raycast (any n)
//!
//! The one change outside this file: `WorldSvo` gains `updated_ranges()` (patches/worldsvo_updated_ranges.patch) because the
//! dirty list is private to the serializers today (internal.rs:166, esvo.rs:321-338).
//!
//! NOT compiled in the libvoxelrt repository (no rustc there). `voxel-rs_b200/host/svo.hpp` is the same call sequence in C++,
//! compiled and tested against the CPU oracle.
use std::cell::RefCell;
use std::ffi::CStr;
use std::mem;
use std::ops::Deref;
use std::ptr;

use cgmath::{EuclideanSpace, Matrix4, Point3, SquareMatrix, Vector3};

use crate::graphics::framebuffer::Framebuffer;
use crate::graphics::svo_picker::{PickerBatch, PickerBatchResult, PickerResult, PickerTask};
use crate::graphics::svo_registry::VoxelRegistry;
use crate::graphics::voxelrt_sys::*;
use crate::world::hds::WorldSvo;

#[derive(Debug, Copy, Clone)]
pub enum SvoType {
    Esvo,
    Csvo,
}

#[derive(Debug, Copy, Clone)]
pub struct SvoTypeProperties {
    pub name: &'static str,
    /// was the value of `#define SVO_TYPE` for the shaders (svo.rs:115-128); now selects the node decode of the CUDA kernels
    pub vx_flags: u32,
}

impl Deref for SvoType {
    type Target = SvoTypeProperties;

    fn deref(&self) -> &Self::Target {
        match self {
            Self::Esvo => &SvoTypeProperties { name: "ESVO", vx_flags: 0 },
            Self::Csvo => &SvoTypeProperties { name: "CSVO", vx_flags: VX_FLAG_SVO_CSVO },
        }
    }
}

/// Largest frame / ray batch a context reserves device memory for. The reference sizes nothing up front (GL allocates
/// per framebuffer); 8K and 16 Mi rays cost 0.5 GB + 1.5 GB of the 180 GB.
const MAX_WIDTH: u32 = 7680;
const MAX_HEIGHT: u32 = 4320;
const MAX_RAYS: u64 = 1 << 24;

pub struct Svo {
    ctx: *mut VxCtx,
    capacity_bytes: usize,
    // host staging of the picker records; RefCell because `raycast` takes `&self` like the reference's (svo.rs:233)
    picker_tasks: RefCell<Vec<PickerTask>>,
    picker_results: RefCell<Vec<PickerResult>>,
    stats: Stats,
}

#[derive(Clone, Copy, Debug)]
pub struct Stats {
    pub used_bytes: usize,
    pub capacity_bytes: usize,
    pub depth: u8,
}

pub struct RenderParams {
    pub ambient_intensity: f32,
    pub light_dir: Vector3<f32>,
    pub cam_pos: Point3<f32>,
    pub cam_fwd: Vector3<f32>,
    pub cam_up: Vector3<f32>,
    pub fov_y_rad: f32,
    pub aspect_ratio: f32,
    pub selected_voxel: Option<Point3<f32>>,
    pub render_shadows: bool,
    pub shadow_distance: f32,
}

impl Svo {
    pub fn new(registry: &VoxelRegistry, typ: SvoType, size_mb: usize) -> Self {
        let cfg = VxConfig {
            device: 0,
            flags: typ.vx_flags,
            svo_capacity_bytes: (size_mb * 1000 * 1000) as u64,   // MappedBuffer::new(size_mb * 1000 * 1000), svo.rs:137
            max_width: MAX_WIDTH,
            max_height: MAX_HEIGHT,
            max_rays: MAX_RAYS,
        };
        let mut ctx: *mut VxCtx = ptr::null_mut();
        unsafe {
            Self::check(ptr::null(), vx_create(&cfg, &mut ctx), "vx_create");

            // registry.build_texture_array() (svo_registry.rs:121-132) minus the GL upload: level-0 RGBA8 images, already
            // flipped vertically like texture_array.rs:92,126, all layers back to back. The library builds the mip chain
            // (texture_array.rs:258-260) and fixes the sampler state of texture_array.rs:200-203.
            let tex = registry.build_texture_pixels().unwrap();   // patches/registry_texture_pixels.patch
            Self::check(ctx, vx_set_textures(ctx, tex.rgba8.as_ptr(), tex.width, tex.height, tex.layers, tex.mip_levels as u32), "vx_set_textures");

            // registry.build_material_buffer() minus the GL buffer: the same Vec<MaterialInstance> (svo_registry.rs:134-165)
            let materials = registry.build_material_instances(&tex.names);
            Self::check(ctx, vx_set_materials(ctx, materials.as_ptr(), materials.len() as u32), "vx_set_materials");
        }

        Self {
            ctx,
            capacity_bytes: size_mb * 1000 * 1000,
            picker_tasks: RefCell::new(Vec::new()),
            picker_results: RefCell::new(Vec::new()),
            stats: Stats { used_bytes: 0, capacity_bytes: 0, depth: 0 },
        }
    }

    /// GL buffer bindings are gone; kept so that callers (`gamelogic::World::new`, world.rs:81) compile unchanged.
    pub fn bind_buffers_globally(&self) {}

    /// Shader hot-reload has no counterpart: the kernels are compiled by build.rs.
    pub fn reload_resources(&mut self) {}

    /// Writes all changes from the given `svo` to the GPU buffer.
    pub fn update<T: WorldSvo<U> + ?Sized, U>(&mut self, svo: &mut T) {
        unsafe {
            let mirror = vx_svo_host_mirror(self.ctx);   // pinned host memory, capacity bytes: stands in for the GL persistent mapping
            let max_depth_exp = (-(svo.depth() as f32)).exp2();   // svo.rs:173

            // the merged dirty list, BEFORE write_changes_to(reset = true) clears it
            let dirty: Vec<VxRange> = svo.updated_ranges().iter()
                .map(|r| VxRange { offset: r.start as u64, length: r.length as u64 })
                .collect();

            // no `render_fence.wait()` (svo.rs:178): the library orders the upload behind the frame / ray batch in flight on the
            // GPU timeline (events between its upload, render and picker streams); the CPU does not stall
            let len = self.capacity_bytes - 1;
            svo.write_changes_to(mirror.add(4), len, true);   // unchanged call, svo.rs:180-181

            Self::check(self.ctx,
                        vx_svo_commit(self.ctx, max_depth_exp, dirty.as_ptr(), dirty.len() as u32, svo.size_in_bytes() as u64, svo.depth() as u32),
                        "vx_svo_commit");   // VX_E_CAPACITY where esvo.rs:328-331 asserts

            self.stats = Stats {
                used_bytes: svo.size_in_bytes(),
                capacity_bytes: self.capacity_bytes,
                depth: svo.depth(),
            };
        }
    }

    pub fn get_stats(&self) -> Stats {
        self.stats
    }

    /// Casts one primary (+ shadow) ray per pixel of `target` (world.glsl main()). The frame stays in device memory
    /// (`vx_frame_device_ptr` for CUDA-GL interop); `Framebuffer::read_pixels` becomes `vx_read_frame_rgba8`.
    pub fn render(&self, params: &RenderParams, target: &Framebuffer) {
        let view_mat = Matrix4::look_to_rh(params.cam_pos, params.cam_fwd, params.cam_up).invert().unwrap();   // svo.rs:197
        let view: &[f32; 16] = view_mat.as_ref();   // column-major, as uploaded by set_f32mat4 (svo.rs:204)

        let mut selected_block = Vector3::new(f32::NAN, f32::NAN, f32::NAN);   // svo.rs:211-215
        if let Some(pos) = params.selected_voxel {
            selected_block = pos.to_vec();
        }

        let p = VxRenderParams {
            view: *view,
            fov_y_rad: params.fov_y_rad,
            aspect_ratio: params.aspect_ratio,
            ambient_intensity: params.ambient_intensity,
            light_dir: params.light_dir.into(),
            cam_pos: params.cam_pos.to_vec().into(),
            highlight_pos: selected_block.into(),
            render_shadows: params.render_shadows as u32,
            shadow_distance: params.shadow_distance,
        };
        unsafe {
            Self::check(self.ctx,
                        vx_render(self.ctx, &p, target.width() as u32, target.height() as u32, ptr::null(), ptr::null_mut()),
                        "vx_render");
        }
    }

    /// `Framebuffer::read_pixels` (framebuffer.rs:97-105) for a frame rendered by [`Svo::render`].
    pub fn read_pixels(&self, width: u32, height: u32) -> Vec<u8> {
        let mut out = vec![0u8; (width * height * 4) as usize];
        unsafe { Self::check(self.ctx, vx_read_frame_rgba8(self.ctx, out.as_mut_ptr()), "vx_read_frame_rgba8"); }
        out
    }

    /// Runs the picker (picker.glsl main()) over `batch`. Synchronous like the reference's fence place + wait (svo.rs:248-249);
    /// the 100-task cap of the mapped buffers (svo.rs:141-142) is gone.
    pub fn raycast(&self, batch: &PickerBatch, result: &mut PickerBatchResult) {
        // upper bound of what serialize_tasks writes: one task per ray, and per AABB up to 3 rays at each of the
        // (ceil(ex)+1)(ceil(ey)+1)(ceil(ez)+1) grid points (svo_picker.rs:183-245)
        let cap = batch.rays.len() + batch.aabbs.iter().map(|a| {
            3 * (a.extents.x.ceil() as usize + 1) * (a.extents.y.ceil() as usize + 1) * (a.extents.z.ceil() as usize + 1)
        }).sum::<usize>();
        let mut tasks = self.picker_tasks.borrow_mut();
        let mut results = self.picker_results.borrow_mut();
        // PickerTask / PickerResult are plain #[repr(C)] records of f32 / bool: all-zero is a valid value of both
        tasks.resize(cap, unsafe { mem::zeroed() });
        results.resize(cap, unsafe { mem::zeroed() });
        let task_count = batch.serialize_tasks(&mut tasks);
        unsafe {
            Self::check(self.ctx, vx_raycast(self.ctx, tasks.as_ptr(), task_count as u64, results.as_mut_ptr()), "vx_raycast");
        }
        batch.deserialize_results(&results[..task_count], result);
    }

    /// The FFI never unwinds; the reference panics on every failure (`unwrap`, `assert!`, `gl_assert_no_error!`), so this does too.
    unsafe fn check(ctx: *const VxCtx, rc: i32, what: &str) {
        if rc != VX_OK {
            let msg = CStr::from_ptr(vx_last_error(ctx)).to_string_lossy().into_owned();
            panic!("{what} failed ({rc}): {msg}");
        }
    }
}

impl Drop for Svo {
    fn drop(&mut self) {
        unsafe { vx_destroy(self.ctx) };
    }
}
