"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): two ranks over NCCL render a tile-sharded frame with per-frame
dirty-range broadcast; rank 0's framebuffer must be byte-identical to the single-GPU frame of the same (edited) world, for
both gather modes — "p2p" (render kernels store into GPU 0's framebuffer over NVLink peer memory) and "nccl" (pack/send/recv)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 1000, 562


def _worker(rank, world_size, port, out_dir, fmt=0):
    sys.path.insert(0, ROOT)
    import ctypes
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_pkg()
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size, device_id=dev)
    try:
        reg = pkg.content_registry(pkg.load_atlas())
        world = pkg.World(radius=5, center=(-1, 2, 5), seed=1, fmt=fmt)
        world.generate(0, 8)
        world.serialize()
        HB = world.header_bytes
        svo = pkg.Svo(reg, size_mb=world.size_bytes // 1_000_000 + 16, max_width=W, max_height=H, max_rays=16, device=rank, flags=world.svo_flags)
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        svo.set_streams(stream.cuda_stream, stream.cuda_stream, stream.cuda_stream)
        world.mark_all_dirty()
        svo.update(world)          # every replica starts from the same SVO
        # a different view every frame: a root that reads its framebuffer too early (stale frame flags, a lost gate) gets pixels of
        # the previous view and the comparison below fails — with identical frames it could not
        views = []
        for fwd in ((1.0, -0.3, 0.0), (0.6, -0.35, 0.5), (0.2, -0.25, -1.0)):
            p = pkg.render_params(cam_pos=(-24.0, 80.0, 174.0), cam_fwd=fwd, fov_y_deg=72.0, aspect=W / H)
            q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
            q.cam_pos = (ctypes.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
            views.append(pkg.to_vx_render_params(q))
        vxp = views[-1]

        # rank 0 edits the world; the other replicas only ever see the broadcast
        meta = [None]
        packed_host = None
        if rank == 0:
            hgt = world.height_at(-10, 174)
            for dy in range(1, 9):
                world.edit_block(-10, hgt + dy, 174, 4)
            world.serialize()
            new = world.gpu_buffer()
            mirror = svo.host_mirror(len(new))
            old = mirror.copy()
            mirror[:] = new
            diff = np.nonzero(new[HB:len(old)] != old[HB:])[0]
            lo, hi = int(diff.min()), int(diff.max()) + 1
            if fmt == 0:
                lo, hi = lo // 4 * 4, (hi + 3) // 4 * 4
            elif lo % 4 == 0 and lo > 0:
                lo -= 1                                                  # CSVO: make sure the unaligned path of the scatter kernel runs
            ranges = [(lo, hi - lo)]
            if len(new) > len(old):
                ranges.append((len(old) - HB, len(new) - len(old)))
            n = svo.pack_dirty(ranges, None)
            packed_host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            svo.pack_dirty(ranges, packed_host.numpy())
            meta = [(len(ranges), n - 16 * len(ranges), world.size_bytes, world.depth)]
        dist.broadcast_object_list(meta, src=0)
        n_ranges, payload, used, depth = meta[0]

        frames = {}
        for gather in ("p2p", "nccl", "p2p8"):
            sf = pkg.sharded.ShardedFrame(svo, rank, world_size, dist=dist, torch=torch, device=dev, gather=gather)
            sf.configure(W, H, max_dirty_bytes=16 * n_ranges + payload)
            # three frames: the later ones exercise the frame-to-frame ordering (gate / signal / wait / release flags) and the
            # one-frame-ahead dirty broadcast on the side stream
            sf.prefetch_dirty(n_ranges, payload, used, depth, packed_host=packed_host)
            for i in range(3):
                sf.apply_dirty()
                if i < 2:
                    sf.prefetch_dirty(n_ranges, payload, used, depth, packed_host=packed_host)
                sf.render(views[i])
                sf.finish()
                if rank == 0 and i == 2:
                    svo.width, svo.height = W, H
                    frames[gather] = svo.read_rgba8() if gather == "p2p8" else svo.read_rgba32f()
                sf.release()
            torch.cuda.synchronize()
            assert svo.frame_sync_errors() == 0, "a frame-flag wait timed out"
            dist.barrier()
            sf.close()
        if rank == 0:
            # single-GPU frame of the edited world on the same context (its replica got the same dirty ranges)
            svo.render_raw(vxp, W, H)
            full = svo.read_rgba32f()
            np.save(os.path.join(out_dir, "full.npy"), full)
            np.save(os.path.join(out_dir, "full8.npy"), svo.read_rgba8())
            for k, f in frames.items():
                np.save(os.path.join(out_dir, f"{k}.npy"), f)
        svo.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fmt", [0, 1], ids=["esvo", "csvo"])
def test_two_gpus_sharded_frame(pkg, tmp_path, fmt):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), fmt), nprocs=2, join=True)
    full = np.load(tmp_path / "full.npy")
    assert np.isfinite(full).all()
    for k in ("p2p", "nccl"):
        f = np.load(tmp_path / f"{k}.npy")
        assert f.tobytes() == full.tobytes(), (k, int((f != full).any(axis=2).sum()))
    full8, f8 = np.load(tmp_path / "full8.npy"), np.load(tmp_path / "p2p8.npy")
    assert f8.tobytes() == full8.tobytes(), int((f8 != full8).any(axis=2).sum())   # RGBA8 gather == RGBA32F frame read through glReadPixels rounding
