// build.rs for voxel-rs with the libvoxelrt ray-cast path (drop this file over the reference's build.rs).
//
// Two jobs:
//   1. what the reference's build.rs does today (build.rs:12-33 of tim-oster/voxel-rs): generate `asset_bundle.rs`;
//      that code is unchanged and only summarised here by the call to `bundle_assets()` — keep the reference's
//      `find_assets` / `generate_asset_bundle` functions below this file's `main` exactly as they are;
//   2. NEW: compile the CUDA library for sm_100a with nvcc and link it (replaces the GLSL world / picker shaders that
//      `graphics::Svo` used to compile at run time, src/graphics/svo.rs:115-128).
//
// Layout expected in the crate root:   voxelrt/include/voxelrt.h   voxelrt/csrc/{voxelrt.cu,traverse.cuh,kernels.cuh,chunks.cuh}
// (= include/ and voxel-rs_b200/csrc/ of the libvoxelrt repository).
//
// NOT compiled in the repository this file ships in: that image has no rustc / cargo. The C++ host mirror
// (voxel-rs_b200/host/svo.hpp) issues the same call sequence and IS compiled and tested there.
use std::env;
use std::path::{Path, PathBuf};
use std::process::Command;

fn main() {
    bundle_assets();
    build_voxelrt();
}

fn build_voxelrt() {
    let manifest_dir = PathBuf::from(env::var_os("CARGO_MANIFEST_DIR").unwrap());
    let out_dir = PathBuf::from(env::var_os("OUT_DIR").unwrap());
    let csrc = manifest_dir.join("voxelrt").join("csrc");
    let header = manifest_dir.join("voxelrt").join("include").join("voxelrt.h");

    for f in ["voxelrt.cu", "traverse.cuh", "kernels.cuh", "chunks.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", header.display());
    println!("cargo:rerun-if-env-changed=NVCC");
    println!("cargo:rerun-if-env-changed=CUDA_HOME");

    let nvcc = env::var_os("NVCC").map(PathBuf::from).unwrap_or_else(|| {
        let home = env::var_os("CUDA_HOME").map(PathBuf::from).unwrap_or_else(|| PathBuf::from("/usr/local/cuda"));
        home.join("bin").join("nvcc")
    });
    let lib = out_dir.join("libvoxelrt.so");
    let status = Command::new(&nvcc)
        .current_dir(&csrc)
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a",   // B200 only: no other architectures, no PTX fallback
            "-O3", "-lineinfo", "-std=c++17",
            "--fmad=false",                                  // numeric contract of the ray path: no FMA contraction
            "-Xcompiler", "-fPIC", "-shared", "-o",
        ])
        .arg(&lib)
        .arg("voxelrt.cu")
        .status()
        .unwrap_or_else(|e| panic!("could not run {}: {e} (set NVCC or CUDA_HOME)", nvcc.display()));
    assert!(status.success(), "nvcc failed to build libvoxelrt.so");

    println!("cargo:rustc-link-search=native={}", out_dir.display());
    println!("cargo:rustc-link-lib=dylib=voxelrt");          // cudart is linked statically into the .so by nvcc
    // multi-GPU (vx_group_*): NCCL is opened with dlopen("libnccl.so.2") at run time, nothing to link here
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out_dir.display());
}

/// The reference's asset bundling (build.rs:12-33), unchanged.
fn bundle_assets() {
    let env_cargo_manifest_dir = env::var_os("CARGO_MANIFEST_DIR").unwrap();
    let env_out_dir = env::var_os("OUT_DIR").unwrap();
    println!("cargo:rerun-if-changed=assets");
    let manifest_dir = Path::new(&env_cargo_manifest_dir);
    let bundle_path = Path::new(&env_out_dir).join("asset_bundle.rs");
    let mut asset_list = Vec::new();
    if env::var_os("CARGO_FEATURE_BUNDLE_ASSETS") == Some(std::ffi::OsString::from("1")) {
        let skip_dirs = rustc_hash::FxHashSet::from_iter(["tests".to_string()]);
        asset_list = find_assets("assets", skip_dirs);
    }
    generate_asset_bundle(manifest_dir, &bundle_path, asset_list).unwrap();
}

// fn find_assets(..) and fn generate_asset_bundle(..): keep the reference's definitions (build.rs:35-118) here.
include!("build_assets.rs");   // = those two functions and `struct Asset`, moved verbatim into their own file
