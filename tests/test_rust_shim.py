"""SURVEY §8f n1 — the voxel-rs side of the drop-in (rust/), checked without rustc (none in the image):

  * every `pub fn vx_*` in rust/src/graphics/voxelrt_sys.rs is exported by libvoxelrt.so and declared in include/voxelrt.h with
    the same number of parameters; every struct size its `abi_sizes` test asserts equals the C struct's;
  * include/voxelrt.h is a C99 header: tests/c_client/client.c (the call sequence of graphics::Svo from plain C) compiles with
    gcc -std=c99 -pedantic -Wall -Werror and links against the library; without a GPU it must stop at vx_create with VX_E_CUDA
    (no CPU fallback); on a GPU it renders its one-voxel world and casts a picker ray (the `-m gpu` half).
"""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYS_RS = os.path.join(ROOT, "rust", "src", "graphics", "voxelrt_sys.rs")
HEADER = os.path.join(ROOT, "include", "voxelrt.h")


def _c_prototypes():
    text = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(vx_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def _rust_externs():
    text = re.sub(r"//.*", "", open(SYS_RS).read())
    block = text[text.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    out = {}
    for m in re.finditer(r"pub fn (vx_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return out


def test_rust_externs_match_the_c_header(pkg):
    protos, externs = _c_prototypes(), _rust_externs()
    assert len(externs) >= 25
    L = pkg.lib()
    for name, n_args in externs.items():
        assert name in protos, f"{name}: declared in voxelrt_sys.rs but not in voxelrt.h"
        assert protos[name] == n_args, f"{name}: {n_args} parameters in Rust, {protos[name]} in C"
        getattr(L, name)   # exported by the built library


def test_rust_struct_sizes_match(pkg):
    text = open(SYS_RS).read()
    sizes = dict(re.findall(r"assert_eq!\(size_of::<(\w+)>\(\), (\d+)\);", text))
    py = {"VxConfig": pkg.VxConfig, "VxRange": pkg.VxRange, "VxShard": pkg.VxShard, "VxStats": pkg.VxStats,
          "VxRenderParams": pkg.VxRenderParams, "MaterialInstance": pkg.VxMaterial}
    for name, cls in py.items():
        assert int(sizes[name]) == C.sizeof(cls), name
    assert int(sizes["PickerTask"]) == pkg.TASK_DTYPE.itemsize == 48 and int(sizes["PickerResult"]) == pkg.RESULT_DTYPE.itemsize == 48
    # field order of the #[repr(C)] structs = field order of the ctypes twins
    for name, cls in (("VxConfig", pkg.VxConfig), ("VxRenderParams", pkg.VxRenderParams), ("VxStats", pkg.VxStats)):
        body = text[text.index(f"pub struct {name} {{"):]
        body = body[:body.index("}")]
        assert re.findall(r"pub (\w+):", body) == [f for f, _ in cls._fields_], name


def _build_client(tmp_path):
    exe = str(tmp_path / "vx_c_client")
    pkg_dir = os.path.join(ROOT, "voxel-rs_b200")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_client", "client.c"), "-o", exe, "-L", pkg_dir, "-lvoxelrt", "-Wl,-rpath," + pkg_dir]
    subprocess.run(cmd, check=True)
    return exe


def test_header_is_c99_and_there_is_no_cpu_fallback(pkg, tmp_path):
    import torch
    exe = _build_client(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the gpu-marked half runs the client")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)          # vx_create -> VX_E_CUDA
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_client_renders_and_picks(pkg, tmp_path):
    exe = _build_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "picker dst=2.000 normal=(0,0,1)" in r.stdout, r.stdout
