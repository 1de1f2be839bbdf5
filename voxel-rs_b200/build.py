"""In-tree build of the native libraries (no JIT cache, the .so files travel with the repo snapshot).

  libvoxelrt.so        CUDA kernels + C ABI (include/voxelrt.h), nvcc, sm_100a only
  libvoxelrs_host.so   C++ host mirror of the reference's Rust interfaces (links libvoxelrt)
  libvoxelrs_world.so  the world-producer half of it alone (octree / ESVO / CSVO serializers, terrain generator): no libvoxelrt

The CPU oracle under oracle/ is test infrastructure with its own Makefile; nothing here builds or loads it.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",  # numeric contract: no FMA contraction anywhere in the ray path (DESIGN.md "Numerics")
    "-Xcompiler", "-fPIC", "-shared",
]


def _cxx():
    # the image exports CXX=/opt/gcc/bin/g++ (no libgomp/older libstdc++); prefer the distro compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_variant(name, defines, verbose=False):
    """A/B build of libvoxelrt with extra -D defines into voxel-rs_b200/variants/<name>/libvoxelrt.so (tools/ab_kernels.py swaps
    it in for one measurement). Not loaded by anything else."""
    src = os.path.join(PKG, "csrc", "voxelrt.cu")
    out_dir = os.path.join(PKG, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libvoxelrt.so")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, src]
    subprocess.run(cmd, check=True, cwd=os.path.join(PKG, "csrc"))
    return out


def build_cuda(force=False, verbose=False):
    src = os.path.join(PKG, "csrc", "voxelrt.cu")
    deps = [src] + [os.path.join(PKG, "csrc", f) for f in ("traverse.cuh", "kernels.cuh", "chunks.cuh", "group.inl")] + [os.path.join(ROOT, "include", "voxelrt.h")]
    out = os.path.join(PKG, "libvoxelrt.so")
    if not force and not _newer(out, deps):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, src, "-ldl"]
    subprocess.run(cmd, check=True, cwd=os.path.join(PKG, "csrc"))
    return out


def build_host(force=False):
    hdir = os.path.join(PKG, "host")
    srcs = [os.path.join(hdir, f) for f in ("capi.cpp", "esvo.cpp")]
    deps = srcs + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".hpp")] + [os.path.join(ROOT, "include", "voxelrt.h")]
    out = os.path.join(PKG, "libvoxelrs_host.so")
    if not force and not _newer(out, deps):
        return out
    cmd = [_cxx(), "-O2", "-std=c++17", "-fPIC", "-Wall", "-shared", "-o", out] + srcs + [
        "-L" + PKG, "-lvoxelrt", "-Wl,-rpath,$ORIGIN", "-lpthread"]
    subprocess.run(cmd, check=True, cwd=hdir)
    # the same sources without the graphics::Svo mirror: world producers only, NO dependency on libvoxelrt. bench.py's reference arm
    # builds its world with this one, so that the CPU baseline process never maps the product library.
    cmd = [_cxx(), "-O2", "-std=c++17", "-fPIC", "-Wall", "-shared", "-DVXH_WORLD_ONLY", "-Wl,--no-undefined", "-o",
           os.path.join(PKG, "libvoxelrs_world.so")] + srcs + ["-lpthread"]
    subprocess.run(cmd, check=True, cwd=hdir)
    return out


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_host(force)


if __name__ == "__main__":
    import sys
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
