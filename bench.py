#!/usr/bin/env python
"""bench.py — headline benchmark of the voxel-rs ray-cast hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): Mrays/s of primary + shadow rays actually cast, and ms/frame, at 4K (3840x2160) on the
reference's generated-terrain world (its own generator restated: Perlin seed 1 + splines; default radius 20 chunks, LOD rule,
camera (-24,80,174) looking along -z, fov 72 deg, sun (-1,-1,-1)/sqrt3, shadows on) = BASELINE.json configs[2]. A "step" is one frame.

  value        rays/s with everything resident in HBM: per step = L2 flush + vx_render (+ for N>1: the frame's dirty SVO ranges —
               NCCL broadcast one frame ahead on a side stream — applied by the scatter kernel, shard render, finished pixels
               stored by the kernels into GPU 0's RGBA8 frame over NVLink peer memory, frame flags as the barrier)
  e2e          the same frame through the C ABI with HOST buffers, every step: dirty ranges copied into the pinned mirror and
               uploaded (vx_svo_commit), vx_render_read_rgba8_begin(frame k), vx_render_read_rgba8_end(frame k-1): one frame's
               RGBA8 read-back is in flight under the next frame's tracing, the last frame is drained inside the timed region.
               e2e.blocking_call = the same step through the one blocking vx_render_read_rgba8; e2e.pcie_d2h = a frame-sized
               D2H copy alone. N>1: every rank DMAs its stripes into one of two host frames shared by the ranks.
  config       names the workload and is identical in both arms (shared_config); how an arm ran is in `details`
  roofline     algorithmic node/texel/frame bytes per launch (SURVEY §8d formula, counts from a counting launch of the
               same frame) / mean kernel time (CUDA events on the launching stream) vs measured HBM copy bandwidth
  cpu_baseline the CPU oracle (a port of the GLSL, NOT the reference itself) on the host cores over a bounded sample

--impl reference times that same CPU port (the reference is Rust + GLSL and cannot run here: no rustc, no GL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "Mrays/s (primary+shadow)"
FLUSH_BYTES = 144 << 20   # L2 flush between steps: a device fill larger than the 126 MB L2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--radius", type=int, default=20)
    ap.add_argument("--terrain", choices=["reference", "standin"], default="reference",
                    help="reference = the reference's own generator (Generator::new(1, cfg): noise 0.8.2 Perlin + splines, pinned by its KATs and "
                         "end-to-end image); standin = hash-gradient noise of the same shape (the world of the profiles up to r01_v5)")
    ap.add_argument("--no-lod", action="store_true")
    ap.add_argument("--format", default="esvo", choices=["esvo", "csvo"], help="SVO type of the world buffer: esvo = the north_star path "
                    "(default); csvo = the reference's default feature (world::hds::csvo + svo.csvo.glsl)")
    ap.add_argument("--no-shadows", action="store_true")
    ap.add_argument("--refill", type=int, default=0, help="refill threshold of the persistent kernel (lanes still walking)")
    ap.add_argument("--refill-shadow", type=int, default=0, help="refill threshold of the shadow-ray kernel alone")
    ap.add_argument("--no-l2-window", action="store_true")
    ap.add_argument("--tma", action="store_true", help="shade kernel writes whole framebuffer strips with TMA bulk copies instead of per-thread 16-byte stores (A/B)")
    ap.add_argument("--gather", default="p2p8", choices=["p2p8", "p2p", "nccl"], help="N>1: tiles to GPU 0 by peer stores from the render "
                    "kernels as RGBA8 pixels (default) or RGBA32F pixels (p2p), or by pack + NCCL send/recv + unpack (nccl)")
    ap.add_argument("--no-overlap", action="store_true", help="serialized wavefront: shade_kernel starts only after trace_primary_kernel has finished (A/B)")
    ap.add_argument("--overlap", action="store_true", help="overlapped wavefront whatever the launch size (default: only for small launches)")
    ap.add_argument("--morton", action="store_true", help="A/B: Z-order enumeration of the macro blocks instead of row-major")
    ap.add_argument("--lifo", type=int, default=-1, help="A/B: vx_set_option 14 value (1 order + store policy, +2 shade discards records, +4 shadow kernel discards its list)")
    ap.add_argument("--no-lifo", action="store_true", help="A/B: streaming stores + producer order for the wavefront buffers instead of the LIFO hand-over (vx_set_option 14)")
    ap.add_argument("--no-clip", action="store_true", help="A/B: rays are not clipped against the occupied box of the world")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush between steps (diagnostic; not a bench line)")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--bands", type=int, default=3, help="e2e: bands of the blocking vx_render_read_rgba8 (render/read-back overlap inside one frame)")
    ap.add_argument("--bands-pipelined", type=int, default=1, help="e2e: bands per frame of the begin/end loop (the read-back overlaps the NEXT frame there)")
    ap.add_argument("--one-stream", action="store_true", help="A/B: uploads on the render stream too (no overlap of the dirty-range scatter with the L2 flush)")
    ap.add_argument("--e2e-sync", action="store_true", help="e2e through the blocking vx_render_read_rgba8 only (no frame in flight across steps)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--workload", default="frame", choices=["frame", "picker", "serialize"], help="frame = BASELINE configs[2] (the metric's config, "
                    "default); picker = configs[3], 16 Mi incoherent picker rays against an r=40 no-LOD world; serialize = SURVEY §8f n3, ESVO "
                    "serialization of every chunk of the r=20 world on the GPU")
    ap.add_argument("--sim-shard", type=int, default=0, help="diagnostic, 1 GPU: render only shard 0 of N of every frame (what one rank of an "
                    "N-GPU run does, without the collectives) — for tuning the small-frame regime; not a bench line")
    ap.add_argument("--group", action="store_true", help="--gpus N from ONE process through vx_group_* (the reference's shape: a single process) "
                    "instead of one rank per GPU under torchrun")
    ap.add_argument("--rays", type=int, default=1 << 24, help="picker workload: number of rays")
    ap.add_argument("--bin", type=int, default=-1, help="picker workload: ray binning (vx_set_option 15): bits per axis of the origin cell's Z-order code, "
                    "+16 = direction octant below it, 0 = trace in task order; default: the library's (0 — binning measured a loss, profiles/r02_picker_binning.md)")
    ap.add_argument("--max-dst", type=float, default=-1.0, help="picker workload: max_dst of every task (-1 = unlimited)")
    return ap.parse_args()


def terrain_name(args):
    return "the reference's generator, seed 1" if args.terrain == "reference" else "stand-in noise"


def build_world(pkg, args):
    t = time.time()
    world = pkg.World(radius=args.radius, center=(-1, 2, 5), seed=1, no_lod=args.no_lod, terrain=args.terrain,
                      fmt=pkg.FORMAT_CSVO if getattr(args, "format", "esvo") == "csvo" else pkg.FORMAT_ESVO)
    world.generate(0, 8)
    world.serialize()
    return world, time.time() - t


def frame_params(pkg, world, args):
    # default player rotation "0 -90 0" (src/main.rs:82-87) through Entity::get_forward (physics.rs:21-27) in f32 = (-4.4e-8, 0, -1): the
    # view of the reference's end-to-end image. The stand-in world keeps the +x view its profiles were taken with.
    yaw = np.float32(np.deg2rad(np.float32(-90.0)))
    fwd = (float(np.cos(yaw)), 0.0, float(np.sin(yaw))) if args.terrain == "reference" else (1.0, 0.0, 0.0)
    p = pkg.render_params(cam_pos=(-24.0, 80.0, 174.0), cam_fwd=fwd, fov_y_deg=72.0, aspect=args.width / args.height,
                          render_shadows=not args.no_shadows)
    import ctypes as C
    q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
    q.cam_pos = (C.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))   # systems::worldsvo::Svo::render, worldsvo.rs:397-409
    return pkg.to_vx_render_params(q)


def workload_name(args):
    return (f"generated-terrain ({terrain_name(args)}) r={args.radius} chunks{' no-LOD' if args.no_lod else ' LOD'}{' (CSVO format)' if getattr(args, 'format', 'esvo') == 'csvo' else ''}, {args.width}x{args.height}, "
            f"primary{'' if args.no_shadows else '+shadow'} rays, camera (-24,80,174) fov72 (BASELINE configs[2])")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _run(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(st, pixels):
    """B = 4*N_iter + 4*N_push + 40*N_leaf + T*N_tex + 16*pixels (node words, child pointers, leaf ptr+value+material,
    4 B per texel read, RGBA32F store)."""
    return 4 * st["steps"] + 4 * st["pushes"] + 40 * st["leaf_tests"] + 4 * st["tex_fetches"] + 16 * pixels


def cpu_sample_bands(height, frac=0.1, bands=27):
    """Row bands spread evenly over the frame (sky, horizon and ground all represented)."""
    rows = max(1, int(height * frac / bands))
    return [(int(i * height / bands), min(height, int(i * height / bands) + rows)) for i in range(bands)]


def cpu_render_sample(scene, vxp, args, threads, budget_s, frac=0.1, min_s=0.0):
    """Times the CPU oracle on row bands of the SAME frame until the budget is used; with min_s the pass is repeated until that much
    wall time has been measured (a whole 4K frame is a fraction of a second on a many-core host: too short a sample otherwise).
    Returns (rays/s, description, rays, seconds)."""
    bands = cpu_sample_bands(args.height, frac) if frac < 1.0 else [(0, args.height)]   # whole frame: one call, no per-band overhead
    out = np.zeros((args.height, args.width, 4), np.float32)
    rays, t_used, rows, passes = 0, 0.0, 0, 0
    t_start = time.time()
    while True:
        for y0, y1 in bands:
            t0 = time.time()
            _, cnt = scene.render(vxp, args.width, args.height, y0, y1, threads=threads, out=out)
            t_used += time.time() - t0
            rays += cnt["primary_rays"] + cnt["shadow_rays"]
            rows += y1 - y0
            if time.time() - t_start > budget_s:
                break
        passes += 1
        if t_used >= min_s or time.time() - t_start > budget_s:
            break
    what = f"{rows} of {args.height} rows in bands spread over the frame" if passes == 1 else f"{passes} passes over {rows // passes} of {args.height} rows"
    return rays / t_used, f"{what} ({rays} rays, {t_used:.1f} s on {threads} threads)", rays, t_used


def host_threads():
    """Threads the CPU arms run on: every core this process may use. NOT omp_get_max_threads(): torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which made the round-1 reference arm run single-threaded at N>1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def shared_config(args):
    """The `config` object of the bench line: what names the workload, IDENTICAL in both arms (`--impl ours` / `--impl reference`) so that
    the driver compares like with like; everything that describes how an arm ran goes to `details`."""
    return {"workload": workload_name(args), "frame": [args.width, args.height], "svo_format": args.format, "shadow_rays": not args.no_shadows,
            "l2": "GPU arm: " + ("not flushed (--no-flush)" if args.no_flush else "flushed between steps by a 144 MiB device fill (> 126 MB L2) inside the timed region") +
                  "; CPU arm: host caches as they are (the 33 MB SVO exceeds any private cache of the box)"}


def run_reference(args):
    """--impl reference: the CPU port of the reference shaders (oracle/), all host threads, bounded sample per step.
    The process never maps the product library: the world is built with libvoxelrs_world.so (VOXELRS_WORLD_ONLY)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["VOXELRS_WORLD_ONLY"] = "1"
    pkg, ora = graft.load_pkg(), graft.load_oracle()
    world, _ = build_world(pkg, args)
    reg = pkg.content_registry(pkg.load_atlas())
    tex, mips = reg.textures()
    scene = ora.Scene(world.gpu_buffer(), reg.materials().tobytes(), tex, mips, fmt=world.fmt)
    vxp = frame_params(pkg, world, args)
    threads = host_threads()
    with open("/proc/self/maps") as f:
        assert "libvoxelrt" not in f.read(), "the reference arm must not map the product library"
    # one step = the whole frame on all host threads (about a third of a second at 4K on 16 threads); if the box is so slow
    # that K+W frames would take more than ~3 minutes, fall back to a 10 % row sample per step
    t0 = time.time()
    cpu_render_sample(scene, vxp, args, threads, 1e9, frac=1.0)
    frac = 1.0 if (time.time() - t0) * (args.steps + args.warmup) < 180.0 else 0.1
    for _ in range(max(0, args.warmup - 1)):
        cpu_render_sample(scene, vxp, args, threads, 1e9, frac=frac)
    tot_rays, tot_t, desc = 0, 0.0, ""
    for _ in range(args.steps):
        _, desc, rays, t = cpu_render_sample(scene, vxp, args, threads, 1e9, frac=frac)
        tot_rays += rays; tot_t += t
    value = tot_rays / tot_t / 1e6
    full_frame_rays = None
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": shared_config(args),
        "details": {"note": "CPU port of the reference GLSL (oracle/), not the Rust+OpenGL reference itself: no rustc / Mesa in this image "
                            "(llvmpipe: not measurable on this box)"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": "per step: " + desc},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "picker":
        return run_picker(args)
    if args.workload == "serialize":
        return run_serialize(args)
    if args.group:
        return run_group(args)

    import torch
    import torch.distributed as dist

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the ray-cast path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world_size
    dev = torch.device("cuda", local_rank)

    graft.build()
    pkg = graft.load_pkg()
    world, gen_s = build_world(pkg, args)
    reg = pkg.content_registry(pkg.load_atlas())
    W, H = args.width, args.height
    size_mb = int(world.size_bytes // 1_000_000 + 64)
    svo = pkg.Svo(reg, size_mb=size_mb, max_width=W, max_height=H, max_rays=1024, device=local_rank,
                  flags=(pkg.VX_FLAG_NO_L2_WINDOW if args.no_l2_window else 0) | world.svo_flags)
    # a real (non-legacy) torch stream shared with the library: torch ops, NCCL collectives and vx_* kernels order on it without
    # host syncs, and torch.cuda.Event timings on it see the library's kernels (stream 0 would mean "keep the library's own streams")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    # render (and picker) work on the bench's stream, where its events are; the upload stream stays the library's own, so that the
    # dirty-range scatter + bounds kernels of a step (ordered behind the previous frame by an event, in front of this frame by another)
    # run next to the L2 flush instead of after it (--one-stream: everything on one stream, the round-1 arrangement)
    svo.set_streams(stream.cuda_stream, stream.cuda_stream if args.one_stream else None, stream.cuda_stream)
    if args.refill:
        svo.set_option(pkg.OPT_REFILL, args.refill)
    if args.refill_shadow:
        svo.set_option(pkg.OPT_REFILL_SHADOW, args.refill_shadow)
    if args.ctas_per_sm:
        svo.set_option(pkg.OPT_CTAS_PER_SM, args.ctas_per_sm)
    if args.tma:
        svo.set_option(pkg.OPT_TMA, 1)
    if args.morton:
        svo.set_option(pkg.OPT_MORTON, 1)
    if args.no_lifo:
        svo.set_option(14, 0)
    elif args.lifo >= 0:
        svo.set_option(14, args.lifo)
    if args.no_clip:
        svo.set_option(pkg.OPT_CLIP, 0)
    if args.no_overlap:
        svo.set_option(pkg.OPT_OVERLAP, 0)
    if args.overlap:
        svo.set_option(pkg.OPT_OVERLAP, 1)
    svo.update(world)
    vxp = frame_params(pkg, world, args)
    shard = (rank, n_gpus) if not args.sim_shard else (0, args.sim_shard)
    pixels = W * H

    # scripted per-frame dirty set: 4 chunk-sized ranges (half-full LOD-5 chunk ~ 112 KB) + the world root
    root_off, root_len = world.root_range()
    chunk_len = 2336 * 48
    stride = max(chunk_len, ((world.size_bytes - chunk_len) // 4) // 48 * 48)
    dirty = [(i * stride, min(chunk_len, world.size_bytes - i * stride)) for i in range(4) if i * stride < world.size_bytes]
    dirty.append((root_off, root_len))
    HB = world.header_bytes   # 24 (ESVO: scale + preamble) or 8 (CSVO: scale + root offset)
    dirty_bytes = sum(l for _, l in dirty) + HB
    octree_scale = float(np.float32(2.0 ** -world.depth))
    packed_n = svo.pack_dirty(dirty, None)
    packed_host = torch.empty(packed_n, dtype=torch.uint8, pin_memory=True)
    packed_hosts = [packed_host, torch.empty(packed_n, dtype=torch.uint8, pin_memory=True)]
    flush_buf = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)   # > the 126 MB L2

    sf = pkg.sharded.ShardedFrame(svo, rank, n_gpus, dist=dist if n_gpus > 1 else None, torch=torch, device=dev, gather=args.gather)
    sf.configure(W, H, max_dirty_bytes=packed_n)
    if rank == 0:
        svo.pack_dirty(dirty, packed_host.numpy())
        if n_gpus > 1:
            for b in sf.dirty_bufs:
                b[:packed_n].copy_(packed_host)
    # the dirty ranges of frame i+1 are broadcast (side stream) while frame i renders: prime the pipeline with frame 1's
    sf.prefetch_dirty(len(dirty), dirty_bytes, world.size_bytes, world.depth)

    def flush():
        if not args.no_flush:
            flush_buf.fill_(1)

    def step_resident():
        """One frame with every input already in HBM (rank 0 holds the packed dirty set on the device)."""
        flush()
        # per-frame changed-chunk ranges: rank 0's packed dirty set -> all GPUs over NVLink (NCCL broadcast, issued one frame
        # ahead on a side stream), applied here by a scatter kernel; then the next frame's set starts travelling
        sf.apply_dirty()
        sf.prefetch_dirty(len(dirty), dirty_bytes, world.size_bytes, world.depth)
        if args.sim_shard:
            svo.render_raw(vxp, W, H, shard=shard)
            return
        sf.render(vxp)
        sf.finish()   # tiles in GPU 0's framebuffer (p2p: stored there by the render kernels; nccl: pack/send/recv/unpack)
        sf.release()

    shm = None
    if n_gpus > 1:
        # e2e at N > 1: host frames shared by the ranks (POSIX shared memory, page-locked in every process; TWO of them, used
        # alternately, so that frame k can be rendered while frame k-1 is still being read back), every rank DMAs the stripes it
        # rendered (VX_SHARD_ROWS) into the frame over its own PCIe link; a word per rank behind the frames says which step's
        # stripes are in (host-side release / acquire: a plain store after the rank's copies of that frame finished, rank 0 polls)
        from multiprocessing import shared_memory
        name = "vxframe_%s" % os.environ.get("MASTER_PORT", "0")
        nbytes = 2 * W * H * 4 + 4096
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except FileNotFoundError:
                pass
            shm = shared_memory.SharedMemory(create=True, size=nbytes, name=name)
        dist.barrier()
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name)
            try:   # rank 0 owns (unlinks) the segment: keep this process' resource tracker from complaining about it at exit
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        shm_np = np.ndarray((nbytes,), dtype=np.uint8, buffer=shm.buf)
        rc = torch.cuda.cudart().cudaHostRegister(shm_np.ctypes.data, nbytes, 1)   # cudaHostRegisterPortable
        assert int(rc) == 0, f"cudaHostRegister: {rc}"
        frame8_nps = [shm_np[i * W * H * 4:(i + 1) * W * H * 4].reshape(H, W, 4) for i in range(2)]
        frame8_ptrs = [shm_np.ctypes.data + i * W * H * 4 for i in range(2)]
        step_words = shm_np[2 * W * H * 4:2 * W * H * 4 + 256].view(np.uint32)   # [r] = last e2e step whose stripes of rank r are in; [63] = rank 0's ack
        if rank == 0:
            step_words[:] = 0
        dist.barrier()
    else:
        frame8 = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True)
        frame8_pair = [frame8, torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True)]   # double-buffered read-back (one frame in flight per buffer)
    e2e_step = [0]
    e2e_pub = [0]                         # N > 1: last frame this rank has published (its stripes are in the host frame)
    e2e_sync = [args.e2e_sync]

    def spin_until(cond, what):
        t_end = time.time() + 30.0
        while not cond():
            if time.time() > t_end:
                raise SystemExit(f"bench.py: rank {rank} gave up waiting for {what}")

    def publish(j, tr=None):
        """N > 1: frame j (the oldest in flight) is waited for and announced; rank 0 returns when every rank's stripes of it are in."""
        svo.render_read_rgba8_end()
        if tr is not None: tr.append(time.perf_counter())
        step_words[rank] = j
        if rank == 0:
            spin_until(lambda: int(step_words[:n_gpus].min()) >= j, f"the stripes of frame {j}")
            step_words[63] = j
        e2e_pub[0] = j
        if tr is not None: tr.append(time.perf_counter())
    trace_on = bool(os.environ.get("VX_BENCH_TRACE"))   # diagnostic: host-side phase times of the N > 1 e2e step on stderr
    traces = []
    gpu_marks = []
    mirror = svo.host_mirror(HB + world.size_bytes)
    staged = [bytes(mirror[HB + o:HB + o + l]) for o, l in dirty]

    def step_e2e():
        """The same frame through the C ABI with host buffers: dirty bytes -> pinned mirror -> H2D, render, RGBA8 -> host."""
        if trace_on: t_s = time.perf_counter()
        flush()
        if rank == 0 and n_gpus == 1:
            for (o, l), b in zip(dirty, staged):   # the host-side serializer writing its changes (write_changes_to)
                mirror[HB + o:HB + o + l] = np.frombuffer(b, np.uint8)
        if n_gpus == 1:
            if not os.environ.get("VX_BENCH_SKIP_COMMIT"):   # diagnostic only (never a bench line): the loop without its upload
                svo.commit(octree_scale, dirty, world.size_bytes, world.depth)
            if e2e_sync[0]:
                # render + read-back in ONE blocking call (the reference's render + glReadPixels): finished bands are copied to the
                # host while the next band is traced; returns when the whole frame is in host memory
                svo.render_read_rgba8(vxp, W, H, frame8.data_ptr(), bands=args.bands)
                return
            # the render loop of a streaming consumer: frame k is begun (its kernels and band copies enqueued), THEN frame k-1 is
            # waited for — its read-back ran under this frame's tracing (two device frames, two host frames). Every step still
            # uploads its own dirty set and reads one whole frame back; the last frame is drained inside the timed region.
            k = e2e_step[0]
            e2e_step[0] += 1
            if trace_on: t_b = time.perf_counter()
            svo.render_read_rgba8_begin(vxp, W, H, frame8_pair[k & 1].data_ptr(), bands=args.bands_pipelined)
            if trace_on: t_e = time.perf_counter()
            if k > 0:
                svo.render_read_rgba8_end()
            if trace_on:
                traces.append([(t_b - t_s) * 1e3, (t_e - t_b) * 1e3, (time.perf_counter() - t_e) * 1e3])
            return
        # N > 1, software-pipelined like the resident loop AND like the 1-GPU loop: this frame's dirty set was packed, copied to GPU 0
        # and broadcast during the previous step, the NEXT frame's set is prepared on the host and sent while this frame renders; and
        # frame k is begun before frame k-1 is waited for (two shared host frames), so a rank's read-back and the cross-process
        # hand-shake of frame k-1 run under the tracing of frame k. Every step still copies one dirty set host -> device and reads
        # one frame device -> host; the last frame is drained inside the timed region.
        e2e_step[0] += 1
        k = e2e_step[0]
        tr = [time.perf_counter()] if trace_on else None
        if trace_on:                           # GPU-side phase marks on the render stream (diagnostic)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            gpu_marks.append(ev)
            ev[0].record(stream)
        sf.apply_dirty()                       # scatter of the set sent last step (stream-ordered behind the previous frame)
        if trace_on: ev[1].record(stream)
        if rank != 0 and k > 2:
            spin_until(lambda: step_words[63] >= k - 2, f"rank 0's ack of frame {k - 2}")   # the host frame of parity k & 1 is free again
        if tr: tr.append(time.perf_counter())
        if rank == 0:                          # the next frame's inputs (host side)
            for (o, l), b in zip(dirty, staged):
                mirror[HB + o:HB + o + l] = np.frombuffer(b, np.uint8)
            svo.pack_dirty(dirty, packed_hosts[k & 1].numpy())
        if tr: tr.append(time.perf_counter())
        # the broadcast of the NEXT frame's set is enqueued BEFORE this frame's kernels: launched behind them, the NCCL kernel found
        # every SM taken by the persistent tracing kernels and ran only in their tails, and the next frame's scatter waited for it
        sf.prefetch_dirty(len(dirty), dirty_bytes, world.size_bytes, world.depth, packed_host=packed_hosts[k & 1])
        if tr: tr.append(time.perf_counter())
        svo.render_read_rgba8_begin(vxp, W, H, frame8_ptrs[k & 1], bands=args.bands_pipelined, shard=(rank, n_gpus | pkg.VX_SHARD_ROWS))
        if trace_on: ev[2].record(stream)
        if tr: tr.append(time.perf_counter())
        if e2e_pub[0] < k - 1:
            publish(k - 1, tr)                 # frame k-1: this rank's stripes are in; rank 0: the frame is whole
        elif tr:
            tr += [time.perf_counter()] * 2
        if tr:
            traces.append([(b - a) * 1e3 for a, b in zip(tr[:-1], tr[1:])])

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    issue_ms = []

    def drain_e2e():
        if n_gpus > 1:
            if e2e_pub[0] < e2e_step[0]:
                publish(e2e_step[0])
            return
        if n_gpus == 1 and not e2e_sync[0]:
            while e2e_step[0] > 0:             # whatever is still in flight (at most the last frame): wait until it is in host memory
                svo.render_read_rgba8_end()
                e2e_step[0] = 0

    def timed(step_fn, steps, warmup, per_step_events=False, drain=None):
        for _ in range(warmup):
            step_fn()
        if drain:
            drain()
        barrier()
        l0 = svo.launch_count()
        evs = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            step_fn()
        if drain:
            drain()                             # inside the timed region: the last frame's read-back counts
        e1.record(stream)
        issue_ms.append((time.time() - t0) * 1e3 / steps)   # host time to ISSUE a step (no sync): close to ms/step = launch-bound
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if n_gpus > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, svo.launch_count() - l0, t0, t1

    # ---- counting launch (same frame, counters on): rays, steps, pushes, leaf tests, texels of THIS rank's shard
    svo.set_option(pkg.OPT_COUNT, 1)
    svo.render_raw(vxp, W, H, shard=shard)
    st = svo.frame_stats(0)
    # the same frame without shadow rays: the primary-ray kernel's own step / push / leaf counts (the shadow kernel's = the difference)
    import ctypes as C
    vxp_ns = pkg.VxRenderParams.from_buffer_copy(bytes(vxp))
    vxp_ns.render_shadows = 0
    svo.render_raw(vxp_ns, W, H, shard=shard)
    st_p = svo.frame_stats(0)
    svo.set_option(pkg.OPT_COUNT, 0)
    rays_local = st["primary_rays"] + st["shadow_rays"]
    if n_gpus > 1:
        t = torch.tensor([rays_local, st["primary_rays"], st["shadow_rays"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        rays_total, prim_total, shad_total = (int(v) for v in t.tolist())
    else:
        rays_total, prim_total, shad_total = rays_local, st["primary_rays"], st["shadow_rays"]

    # ---- kernel-only timing of the dominant kernel (events on its stream, L2 flushed before each launch)
    kms, parts = [], []
    for i in range(args.warmup + min(args.steps, 10)):
        flush()
        svo.render_raw(vxp, W, H, shard=shard)
        if i >= args.warmup:
            fs = svo.frame_stats(0)
            kms.append(fs["kernel_ms"])
            parts.append((fs["trace_ms"], fs["shade_ms"], fs["shadow_ms"]))
    kernel_ms = float(np.mean(kms))
    split_ms = dict(zip(("trace_primary", "shade", "trace_shadow"), (round(float(v), 4) for v in np.mean(np.array(parts), axis=0))))

    # ---- N > 1: the gathered frame must be the frame. One sharded step, then rank 0 renders the same frame alone and compares
    # byte for byte (RGBA8 in p2p8 mode, RGBA32F otherwise). A mismatch voids the line: the run stops.
    parity_check = None
    if n_gpus > 1:
        read = svo.read_rgba8 if args.gather == "p2p8" else svo.read_rgba32f
        step_resident()
        barrier()
        got = read() if rank == 0 else None
        barrier()
        if rank == 0:
            svo.render_raw(vxp, W, H, shard=None)
            want = read()
            same = got.tobytes() == want.tobytes()
            parity_check = (f"frame gathered from {n_gpus} GPUs == rank 0's unsharded render, byte for byte ({got.nbytes} bytes, "
                            f"{'RGBA8' if args.gather == 'p2p8' else 'RGBA32F'})") if same else "MISMATCH"
        flag = torch.tensor([1 if (rank != 0 or parity_check != "MISMATCH") else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            raise SystemExit("bench.py: the gathered multi-GPU frame differs from the single-GPU frame — no number reported")

    # ---- the timed region
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, launches, t0, t1 = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_per_step = ms_total / args.steps
    value = rays_total / (ms_per_step * 1e-3) / 1e6

    e2e = None
    if not args.skip_e2e:
        ms_e2e, _, _, _ = timed(step_e2e, args.steps, args.warmup, drain=drain_e2e)
        sync_path = ("host dirty ranges -> pinned mirror -> vx_svo_commit (H2D) -> vx_render_read_rgba8 (%d bands: trace/shade overlapped with the "
                     "RGBA8 D2H copy into pinned host memory), one blocking call per frame" % args.bands)
        e2e = {"value": rays_total / (ms_e2e / args.steps * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_e2e / args.steps,
               "h2d_bytes_per_step": int(dirty_bytes + len(dirty) * 16), "d2h_bytes_per_step": int(W * H * 4),
               "path": (sync_path if e2e_sync[0] else
                        "host dirty ranges -> pinned mirror -> vx_svo_commit (H2D) -> vx_render_read_rgba8_begin(frame k) -> vx_render_read_rgba8_end(frame k-1): "
                        "two frames in flight (two device RGBA8 frames, two pinned host frames), so frame k-1's D2H copy runs under frame k's tracing; "
                        "every step uploads its dirty set and reads one whole frame back, the last frame is drained inside the timed region "
                        "(%d band(s) per frame)" % args.bands_pipelined) if n_gpus == 1 else
                       "host dirty ranges -> pack -> H2D -> NCCL broadcast (sent one step ahead, while the previous frame renders) -> scatter -> every rank renders whole 16-pixel stripes "
                       "(VX_SHARD_ROWS) and DMAs them itself into a page-locked host frame shared by the ranks (N PCIe links; two such frames, used alternately: "
                       "frame k is begun before frame k-1 is waited for); rank 0 returns from a step when all stripes of frame k-1 are in host memory; "
                       "the last frame is drained inside the timed region"}
        if n_gpus == 1:
            # what the PCIe link of this box gives a frame-sized device -> pinned-host copy on its own (nothing else running): the floor of
            # any end-to-end loop that returns a whole RGBA8 frame per step
            src8 = torch.empty(W * H * 4, dtype=torch.uint8, device=dev)
            dst8 = frame8_pair[1].view(-1)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dst8.copy_(src8, non_blocking=True)
            torch.cuda.synchronize()
            c0.record(stream)
            for _ in range(5):
                dst8.copy_(src8, non_blocking=True)
            c1.record(stream)
            torch.cuda.synchronize()
            copy_ms = c0.elapsed_time(c1) / 5
            e2e["pcie_d2h"] = {"frame_bytes": int(W * H * 4), "copy_ms": round(copy_ms, 4), "gb_per_s": round(W * H * 4 / (copy_ms * 1e-3) / 1e9, 2),
                               "note": "a frame-sized D2H copy alone on this box; the pipelined loop cannot go below max(device frame time, this)"}
            del src8
        if n_gpus == 1 and not e2e_sync[0]:
            # the same step through the one blocking call (frame k is in host memory when the call returns): reported beside the pipelined loop
            e2e_sync[0] = True
            ms_sync, _, _, _ = timed(step_e2e, args.steps, args.warmup)
            e2e_sync[0] = False
            e2e["blocking_call"] = {"value": rays_total / (ms_sync / args.steps * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_sync / args.steps, "path": sync_path}
        if trace_on and traces and n_gpus == 1:
            m = np.mean(np.array(traces[-args.steps:]), axis=0)
            print(f"[trace e2e N=1] flush+mirror+commit {m[0]:.3f} begin {m[1]:.3f} end(k-1) {m[2]:.3f} ms (host time per step)", file=sys.stderr, flush=True)
        elif trace_on and traces and rank in (0, 1):
            torch.cuda.synchronize()
            gm = gpu_marks[-args.steps:]
            flush_scatter = np.mean([a[0].elapsed_time(a[1]) for a in gm])
            render = np.mean([a[1].elapsed_time(a[2]) for a in gm])
            idle = np.mean([a[2].elapsed_time(b[0]) for a, b in zip(gm[:-1], gm[1:])])
            print(f"[trace rank {rank}] GPU render stream per frame: flush+scatter(+wait for the broadcast) {flush_scatter:.3f}  render {render:.3f}  idle until the next frame's first launch {idle:.3f} ms",
                  file=sys.stderr, flush=True)
            m = np.mean(np.array(traces[-args.steps:]), axis=0)
            print(f"[trace rank {rank}] apply+ack-wait {m[0]:.3f} host-prep {m[1]:.3f} prefetch {m[2]:.3f} begin {m[3]:.3f} end {m[4]:.3f} poll {m[5]:.3f} ms",
                  file=sys.stderr, flush=True)
        if n_gpus > 1 and rank == 0:
            # the host frame of the last e2e step against rank 0's own unsharded render: the shared frame is the frame
            svo.render_raw(vxp, W, H, shard=None)
            ref8 = svo.read_rgba8()
            if ref8.tobytes() != frame8_nps[e2e_step[0] & 1].tobytes():
                raise SystemExit("bench.py: the host frame assembled from the ranks' stripes differs from the single-GPU frame")
            e2e["parity_check"] = "shared host frame == rank 0's unsharded RGBA8 frame, byte for byte"

    if n_gpus > 1:
        # a frame-flag wait that timed out (lost / slow peer) means torn frames: the numbers above would be of garbage
        errs = torch.tensor([svo.frame_sync_errors()], device=dev)
        dist.all_reduce(errs)
        if int(errs.item()):
            raise SystemExit(f"bench.py: {int(errs.item())} frame-flag waits timed out during the run — no number reported")
    def release_shm():
        if shm is not None:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(shm_np.ctypes.data)
            dist.barrier()
            shm.close()
            if rank == 0:
                shm.unlink()

    if rank != 0:
        if n_gpus > 1:
            dist.barrier()
            release_shm()
            sf.close()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # measured read bandwidth of THIS GPU at the two levels the node fetches can be served from (report against both)
    l2_gbs = svo.probe_read_bandwidth(32 << 20, 64)
    hbm_read_gbs = svo.probe_read_bandwidth(2 << 30, 1)
    alg_bytes = algorithmic_bytes(st, pixels // n_gpus if n_gpus > 1 else pixels)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    # per kernel: node words + child pointers + leaf words, plus the wavefront records each kernel reads / writes
    hits = st_p["leaf_tests"]                       # opaque worlds: one accepted leaf per hit pixel (upper bound otherwise)
    px_local = st["primary_rays"]
    shadow_entries = st["shadow_rays"]
    k_alg = {
        "trace_primary": 4 * st_p["steps"] + 4 * st_p["pushes"] + 4 * st_p["leaf_tests"] + 32 * min(hits, px_local) + 16 * max(px_local - hits, 0),
        "shade": 32 * min(hits, px_local) + 16 * max(px_local - hits, 0) + 32 * min(hits, px_local) + 4 * st_p["tex_fetches"] +
                 36 * shadow_entries + 16 * (px_local - shadow_entries),
        "trace_shadow": 36 * shadow_entries + 4 * (st["steps"] - st_p["steps"]) + 4 * (st["pushes"] - st_p["pushes"]) +
                        4 * (st["leaf_tests"] - st_p["leaf_tests"]) + 16 * shadow_entries,
    }
    per_kernel = {k: {"algorithmic_bytes_per_launch": int(v), "ms": split_ms[k], "achieved_gbs": round(v / (split_ms[k] * 1e-3) / 1e9, 1),
                      "frac": round(v / (split_ms[k] * 1e-3) / 1e9 / peak, 4)} for k, v in k_alg.items() if split_ms[k] > 0}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"]) if per_kernel else None
    traffic, issue = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if n_gpus == 1 and (W, H) == (3840, 2160) and not args.no_shadows and args.terrain == "reference" and args.format == "esvo" \
                and args.radius == 20 and not args.no_lod:   # the ncu capture is of this exact frame
            traffic = tj.get("render_kernel_dram_bytes_per_launch")
            wi = tj.get("warp_instructions_per_frame")
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = (clocks or {}).get("sm_mhz") or 1965.0
            bound_ms = wi / (sms * 4 * mhz * 1e6) * 1e3
            issue = {"bound": "instruction issue (the binding resource, DESIGN.md §6)", "warp_instructions_per_frame": wi,
                     "issue_slots_per_s": sms * 4 * mhz * 1e6, "bound_ms": bound_ms, "frac": bound_ms / kernel_ms, "source": tj.get("source")}
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(shared_config(args), **({"parity_check": parity_check} if n_gpus > 1 else {})),
        "details": {
            "svo_bytes": int(world.size_bytes), "svo_depth": int(world.depth),
            "chunks": int(world.chunk_count), "rays_per_frame": rays_total, "primary_rays": prim_total, "shadow_rays": shad_total,
            "parallelism": f"image tiles (32x16 px macro blocks, interleaved) over {n_gpus} GPU(s), SVO replicated",
            "kernels": "wavefront: trace_primary (persistent) -> shade -> trace_shadow (persistent)" + ("" if args.no_overlap else
                       "; for small launches (shards, small frames) shade runs next to trace_primary on its own stream, its CTAs wait per 32x4-pixel strip"), "ctas_per_sm": args.ctas_per_sm or 8,
            "refill_threshold": args.refill or 1,
            "sim_shard": args.sim_shard or None, "l2_window": not args.no_l2_window, "tma_tile_writeback": args.tma, "lifo_wavefront_buffers": (args.lifo if args.lifo > 0 and not args.no_lifo else 0), "world_gen_s": round(gen_s, 2), "parity_check": parity_check,
            "multi_gpu_step": (None if n_gpus == 1 else "NCCL broadcast of packed dirty ranges + scatter kernel, shard render, " +
                               ("finished pixels stored by the shade/shadow kernels straight into GPU 0's %s framebuffer over NVLink peer memory, "
                                "frame flags in GPU 0's memory as the barrier; the broadcast of frame i+1 overlaps frame i on a side stream"
                                % ("RGBA8" if args.gather == "p2p8" else "RGBA32F")
                                if args.gather != "nccl" else "pack, NCCL send/recv to GPU 0, unpack")),
        },
        "frame_ms": ms_per_step,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": "trace_primary_kernel + shade_kernel + trace_shadow_kernel (one frame; SURVEY §8d formula)",
                     "kernel_ms": kernel_ms, "kernel_ms_split": split_ms, "issue": issue, "algorithmic_bytes_per_launch": int(alg_bytes),
                     "l2": {"bound": "l2", "achieved": achieved, "peak": round(l2_gbs, 1), "unit": "GB/s", "frac": achieved / l2_gbs if l2_gbs else None,
                            "peak_source": "measured here: vx_probe_read_bandwidth, 32 MiB buffer x 64 passes, ld.global.cg (L2-resident)",
                            "hbm_read_peak_measured_here": round(hbm_read_gbs, 1),
                            "note": "the same algorithmic bytes against the L2: the frame touches a few MB of SVO, so node fetches are L1 / L2 hits"},
                     "dominant_kernel": (dict(per_kernel[dom], kernel=dom + "_kernel") if dom else None), "per_kernel": per_kernel,
                     "counts": {k: int(st[k]) for k in ("primary_rays", "shadow_rays", "steps", "pushes", "leaf_tests", "tex_fetches")},
                     "note": "latency/divergence-bound pointer chasing: the SVO is L2-resident after first touch, so the HBM fraction is small "
                             "by construction (SURVEY §8d); see profiles/ for L2 hit rate and warp execution efficiency"},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "host_issue_ms_per_step": round(issue_ms[0], 4) if issue_ms else None,
    }
    if e2e:
        line["e2e"] = e2e
    if not args.skip_cpu and n_gpus == 1:
        ora = graft.load_oracle()
        tex, mips = reg.textures()
        scene = ora.Scene(world.gpu_buffer(), reg.materials().tobytes(), tex, mips, fmt=world.fmt)
        threads = host_threads()
        rps, desc, _, _ = cpu_render_sample(scene, vxp, args, threads, args.cpu_seconds, frac=1.0, min_s=min(1.5, args.cpu_seconds))
        line["cpu_baseline"] = {"value": rps / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": desc}
    print(json.dumps(line), flush=True)

    if n_gpus > 1:
        dist.barrier()
        release_shm()
        sf.close()
        dist.destroy_process_group()


def run_group(args):
    """--group: the same frame on --gpus N devices driven from ONE process through vx_group_* (the reference engine is a single
    process; this is the drop-in's own multi-GPU mode). value = resident frames (vx_group_svo_commit of the frame's dirty ranges:
    packed H2D -> ncclBroadcast -> scatter; vx_group_render: peer-store gather into device 0's RGBA32F frame); e2e = the same
    dirty ranges + vx_group_render_read_rgba8 into the group's page-locked host frame (every device DMAs its own stripes).
    Timed on the host around K steps (the work spans N devices; every step ends with all devices idle)."""
    import ctypes as C
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — no CPU fallback")
    n = args.gpus
    graft.build()
    pkg = graft.load_pkg()
    world, gen_s = build_world(pkg, args)
    reg = pkg.content_registry(pkg.load_atlas())
    W, H = args.width, args.height
    grp = pkg.SvoGroup(reg, list(range(n)), size_mb=int(world.size_bytes // 1_000_000 + 64), max_width=W, max_height=H, max_rays=1024, flags=world.svo_flags)
    one = pkg.Svo(reg, size_mb=int(world.size_bytes // 1_000_000 + 64), max_width=W, max_height=H, max_rays=1024, flags=world.svo_flags)
    world.mark_all_dirty()
    one.update(world)
    world.mark_all_dirty()
    grp.update(world)
    vxp = frame_params(pkg, world, args)
    one.set_option(pkg.OPT_COUNT, 1)
    one.render_raw(vxp, W, H)
    st = one.frame_stats(0)
    one.set_option(pkg.OPT_COUNT, 0)
    want32, want8 = one.read_rgba32f(), one.read_rgba8()
    rays = st["primary_rays"] + st["shadow_rays"]
    # parity before timing: both group paths against the single-GPU context
    grp.render_raw(vxp, W, H)
    ok32 = grp.read_rgba32f().tobytes() == want32.tobytes()
    host = grp.host_frame(W, H)
    grp.render_read_rgba8(vxp, W, H, host.ctypes.data, bands=min(args.bands, 2))
    ok8 = host.tobytes() == want8.tobytes()
    if not (ok32 and ok8):
        raise SystemExit(f"bench.py --group: group frame differs from the single-GPU frame (rgba32f ok={ok32}, rgba8 ok={ok8})")
    one.close()

    root_off, root_len = world.root_range()
    chunk_len = 2336 * 48
    stride = max(chunk_len, ((world.size_bytes - chunk_len) // 4) // 48 * 48)
    dirty = [(i * stride, min(chunk_len, world.size_bytes - i * stride)) for i in range(4) if i * stride < world.size_bytes] + [(root_off, root_len)]
    HB = world.header_bytes
    dirty_bytes = sum(l for _, l in dirty) + HB
    arr = (pkg.VxRange * len(dirty))(*[pkg.VxRange(o, l) for o, l in dirty])
    scale = float(np.float32(2.0 ** -world.depth))
    L = pkg.lib()
    flush = [torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=torch.device("cuda", i)) for i in range(n)]

    def sync():
        for i in range(n):
            torch.cuda.synchronize(i)

    def step(e2e):
        for b in flush:
            b.fill_(1)
        grp._check(L.vx_group_svo_commit(grp.g, scale, arr, len(dirty), world.size_bytes, world.depth))
        if e2e:
            grp.render_read_rgba8(vxp, W, H, host.ctypes.data, bands=min(args.bands, 2))
        else:
            grp.render_raw(vxp, W, H)

    def timed(e2e):
        for _ in range(args.warmup):
            step(e2e)
        sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(e2e)
        if not e2e:
            grp.wait()
        sync()
        return (time.perf_counter() - t0) * 1e3 / args.steps, t0

    sampler = ClockSampler(0)
    t_wall0 = time.time()
    ms, _ = timed(False)
    clocks = sampler.stop(t_wall0, time.time())
    ms_e2e, _ = timed(True)
    line = {
        "metric": METRIC, "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "frame": [W, H], "svo_bytes": int(world.size_bytes), "rays_per_frame": rays,
                   "parallelism": f"ONE process, vx_group_* over {n} GPU(s): SVO replicated, dirty ranges by ncclBroadcast, interleaved macro blocks "
                                  "stored into device 0's frame over NVLink peer memory (value) / whole stripes DMA'd by every device into one "
                                  "page-locked host frame (e2e)",
                   "l2": "flushed between steps: 144 MiB device fill per GPU inside the timed region", "timing": "host clock around K steps, all devices synchronised",
                   "parity_check": "vx_group_render == single-GPU RGBA32F frame and vx_group_render_read_rgba8 == single-GPU RGBA8 frame, byte for byte",
                   "world_gen_s": round(gen_s, 2)},
        "e2e": {"value": rays / (ms_e2e * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(dirty_bytes + len(dirty) * 16),
                "d2h_bytes_per_step": int(W * H * 4), "path": "host dirty ranges -> vx_group_svo_commit -> vx_group_render_read_rgba8 (host frame)"},
        "clocks": clocks, "gpu_launches": int(sum(L.vx_launch_count(grp.ctx(i)) for i in range(n))),
    }
    print(json.dumps(line), flush=True)
    grp.close()


def run_serialize(args):
    """--workload serialize = SURVEY §8f n3: every chunk of the generated r=20 world (dense 32^3 BlockId arrays + the LOD rule)
    through vx_serialize_chunks_esvo. value = chunks/s with the block arrays resident in HBM; e2e = host block arrays -> H2D ->
    kernel -> chunk infos back (records stay on the GPU, where the ray caster reads them); cpu_baseline = the host serializer
    (the C++ restatement of SerializedChunk::new, pinned on esvo.rs:561-1228) on all host threads. Single GPU."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — no CPU fallback")
    graft.build()
    pkg = graft.load_pkg()
    world, gen_s = build_world(pkg, args)
    chunks = world.chunks()
    cb = [world.chunk_blocks(c) for c in chunks]
    blocks = np.stack([b for b, _ in cb])
    lods = np.array([l for _, l in cb], dtype=np.uint8)
    n = len(chunks)
    reg = pkg.content_registry(pkg.load_atlas())
    svo = pkg.Svo(reg, size_mb=1, max_width=8, max_height=8, max_rays=8)
    b_host = torch.from_numpy(blocks.view(np.uint8).reshape(-1)).pin_memory()
    b_dev = b_host.cuda()
    flush_buf = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
    infos, rec, _ = svo.serialize_chunks(blocks, lods)
    out_bytes = len(rec)
    # parity of the whole batch against the host's RangeBuffer image
    image = world.range_bytes()
    ok = all(rec[int(i["offset_bytes"]):int(i["offset_bytes"] + i["length_bytes"])].tobytes() ==
             image[world.chunk_range(c)[0]:world.chunk_range(c)[0] + world.chunk_range(c)[1]].tobytes() for c, i in zip(chunks, infos))

    def run(ptr, steps, warmup):
        kms = []
        for i in range(warmup + steps):
            flush_buf.fill_(1)
            torch.cuda.synchronize()
            t0 = time.time()
            _, _, ms = svo.serialize_chunks(None, lods, want_records=False, blocks_ptr=ptr, n_chunks=n)
            t = (time.time() - t0) * 1e3
            if i >= warmup:
                kms.append((ms, t))
        return float(np.mean([k for k, _ in kms])), float(np.mean([t for _, t in kms]))

    sampler = ClockSampler(0)
    t0 = time.time()
    kernel_ms, _ = run(b_dev.data_ptr(), args.steps, args.warmup)
    clocks = sampler.stop(t0, time.time())
    _, e2e_ms = run(b_host.data_ptr(), max(3, args.steps // 3), 2)
    peak, peak_src = measured_peaks()
    alg = blocks.nbytes + out_bytes                       # every block id read once, every record written once
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    threads = os.cpu_count() or 1
    tc = time.time(); reps = 0
    while time.time() - tc < min(args.cpu_seconds, 10.0) or reps < 2:
        pkg.host().vxh_serialize_dense_batch(blocks.ctypes.data, n, lods.ctypes.data, threads)
        reps += 1
    cpu_s = (time.time() - tc) / reps
    line = {
        "metric": "chunks/s (ESVO chunk serialization)", "value": n / (kernel_ms * 1e-3), "unit": "chunks/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"ESVO serialization of the {n} chunks of the generated-terrain ({terrain_name(args)}) r={args.radius} world with the LOD rule (SURVEY §8f n3)",
                   "in_bytes": int(blocks.nbytes), "out_bytes": int(out_bytes), "l2": "flushed before every launch (160 MiB fill)",
                   "parity": "byte-identical to the host serializer's RangeBuffer image" if ok else "MISMATCH vs host serializer", "world_gen_s": round(gen_s, 2)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "serialize_chunks_kernel", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": int(alg)},
        "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "chunks/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(blocks.nbytes + n),
                "d2h_bytes_per_step": int(n * 24), "path": "pinned host block arrays -> H2D -> serialize_chunks_kernel -> chunk infos to host (records stay in HBM)"},
        "cpu_baseline": {"value": n / cpu_s, "unit": "chunks/s", "cores": threads, "kind": "port",
                         "sample": f"the same {n} chunks through the host serializer on {threads} threads x {reps} passes"},
        "clocks": clocks, "gpu_launches": int(args.steps),
    }
    print(json.dumps(line), flush=True)
    svo.close()


def picker_tasks(pkg, world, radius, n, seed=0, max_dst=-1.0):
    """BASELINE configs[3]: uniform random origins in the world AABB above the terrain band, uniform random directions."""
    rng = np.random.default_rng(seed)
    size = 32.0 * (2 * radius + 1)
    tasks = np.zeros(n, dtype=pkg.TASK_DTYPE)
    tasks["max_dst"] = max_dst
    pos = rng.uniform(0, size, (n, 3)).astype(np.float32)
    origin = world.cnv_block_pos((0.0, 0.0, 0.0))                      # SVO-space position of world block (0,0,0)
    pos[:, 1] = origin[1] + rng.uniform(60.0, 260.0, n).astype(np.float32)   # terrain heights are 20..200 (gamelogic/world.rs:56-78)
    tasks["pos"] = pos
    d = rng.normal(size=(n, 3)).astype(np.float32)
    tasks["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return tasks


def run_picker(args):
    """--workload picker = BASELINE configs[3]: N incoherent picker rays (picker.glsl) against a large no-LOD world whose SVO
    exceeds L2 — the divergence / HBM stress case. Rays are split into contiguous ranges over the ranks (no collective)."""
    import torch
    import torch.distributed as dist
    world_size = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the ray-cast path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    graft.build()
    pkg = graft.load_pkg()
    radius = args.radius if args.radius != 20 else 40
    t0 = time.time()
    world = pkg.World(radius=radius, center=(-1, 2, 5), seed=1, no_lod=True, terrain=args.terrain,
                      fmt=pkg.FORMAT_CSVO if args.format == "csvo" else pkg.FORMAT_ESVO)
    world.generate(0, 8)
    world.serialize()
    gen_s = time.time() - t0
    reg = pkg.content_registry(pkg.load_atlas())
    n_total = args.rays
    n = n_total // world_size
    svo = pkg.Svo(reg, size_mb=int(world.size_bytes // 1_000_000 + 64), max_width=32, max_height=16, max_rays=n, device=local_rank, flags=world.svo_flags)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    svo.set_streams(stream.cuda_stream, stream.cuda_stream, stream.cuda_stream)
    if args.refill:
        svo.set_option(pkg.OPT_REFILL_PICKER, args.refill)
    if args.bin >= 0:
        svo.set_option(15, args.bin)
    svo.update(world)
    tasks = picker_tasks(pkg, world, radius, n_total, max_dst=args.max_dst)[rank * n:(rank + 1) * n]
    t_host = torch.from_numpy(tasks.view(np.uint8).reshape(-1)).pin_memory()
    r_host = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
    t_dev = t_host.to(dev)
    r_dev = torch.empty(n * 48, dtype=torch.uint8, device=dev)
    flush_buf = torch.empty(160 << 20, dtype=torch.uint8, device=dev)

    svo.set_option(pkg.OPT_COUNT, 1)
    svo.raycast_device(t_dev.data_ptr(), n, r_dev.data_ptr())
    st = svo.frame_stats(1)
    svo.set_option(pkg.OPT_COUNT, 0)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        l0 = svo.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world_size > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / args.steps, svo.launch_count() - l0, t0, time.time()

    def step_resident():
        flush_buf.fill_(1)
        svo.raycast_device(t_dev.data_ptr(), n, r_dev.data_ptr())

    def step_e2e():
        flush_buf.fill_(1)
        import ctypes as C
        svo._check(pkg.lib().vx_raycast(svo.ctx, C.c_void_p(t_host.data_ptr()), n, C.c_void_p(r_host.data_ptr())))

    kms = []
    for i in range(args.warmup + min(args.steps, 10)):
        flush_buf.fill_(1)
        svo.raycast_device(t_dev.data_ptr(), n, r_dev.data_ptr())
        if i >= args.warmup:
            kms.append(svo.frame_stats(1)["kernel_ms"])
    kernel_ms = float(np.mean(kms))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, t0, t1 = timed(step_resident)
    clocks = sampler.stop(t0, t1) if sampler else None
    e2e_ms = None
    if not args.skip_e2e:
        e2e_ms, _, _, _ = timed(step_e2e)
    if rank != 0:
        if world_size > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    alg = 4 * st["steps"] + 4 * st["pushes"] + 96 * n      # node words + child pointers + 48-B task in + 48-B result out (the picker reads no leaf word)
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("picker_kernel_dram_bytes_per_launch") if (world_size == 1 and n_total == 1 << 24 and args.terrain == "reference"
                                                                              and args.format == "esvo" and radius == 40) else None
    except Exception:
        pass
    hits = None
    line = {
        "metric": "Mrays/s (picker rays)", "value": n_total / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world_size, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{n_total} random-origin random-direction picker rays, generated-terrain ({terrain_name(args)}) r={radius} no-LOD world (BASELINE configs[3])",
                   "svo_bytes": int(world.size_bytes), "svo_depth": int(world.depth), "chunks": int(world.chunk_count), "max_dst": args.max_dst,
                   "parallelism": f"contiguous ray ranges over {world_size} GPU(s), SVO replicated, no collective", "refill_threshold": args.refill or 20,
                   "ray_binning": ("off (task order)" if args.bin <= 0 else f"Z-order of origin cells, {args.bin & 15} bits/axis" + (" + direction octant" if args.bin & 16 else "") + " (bin_count/scan/scatter kernels inside the timed region)"),
                   "l2": "flushed between steps (160 MiB fill in the timed region); the SVO itself is larger than L2", "world_gen_s": round(gen_s, 2)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": "trace_picker_kernel", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": int(alg),
                     "counts": {k: int(st[k]) for k in ("steps", "pushes", "leaf_tests")}},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if e2e_ms is not None:
        line["e2e"] = {"value": n_total / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(48 * n),
                       "d2h_bytes_per_step": int(48 * n), "path": "vx_raycast: pinned host tasks -> H2D -> trace_picker_kernel -> D2H results, in slices of 1 Mi rays so that upload, tracing and read-back overlap (full-duplex PCIe); returns after the last copy, like the reference's fence"}
    if not args.skip_cpu and world_size == 1:
        ora = graft.load_oracle()
        tex, mips = reg.textures()
        scene = ora.Scene(world.gpu_buffer(), reg.materials().tobytes(), tex, mips, fmt=world.fmt)
        threads = host_threads()
        m = min(n, 1 << 18)
        scene.raycast(tasks[:4096], threads=threads)
        tc = time.time()
        reps = 0
        while time.time() - tc < args.cpu_seconds and reps < 64:
            want, _ = scene.raycast(tasks[:m], threads=threads)
            reps += 1
        dt = time.time() - tc
        line["cpu_baseline"] = {"value": m * reps / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                "sample": f"first {m} rays of the same batch x {reps} passes ({dt:.1f} s)"}
        got = r_dev[:m * 48].cpu().numpy().view(pkg.RESULT_DTYPE)
        line["config"]["sample_parity"] = "byte-identical to the oracle on the CPU sample" if got.tobytes() == want.tobytes() else "MISMATCH vs oracle"
    print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
