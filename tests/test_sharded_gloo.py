"""N>1 host logic on CPU: two processes over gloo run the SAME ShardedFrame orchestration the GPU bench runs over NCCL —
dirty-range pack -> broadcast -> apply on the replica, tile-sharded render, tiles gathered to rank 0 — with a host stand-in
for the device engine (numpy buffers + the CPU oracle as the pixel producer; the oracle is test infrastructure).

Checks: (1) after the broadcast the replica's world buffer is byte-identical to rank 0's re-serialized buffer although
rank 1 never saw the edit; (2) rank 0's gathered frame is byte-identical to an unsharded oracle frame; (3) shard geometry:
every pixel has exactly one owner, ragged frame sizes included.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 100, 70   # ragged: not a multiple of the 32x16 macro block


class HostEngine:
    """Stand-in with the method surface ShardedFrame uses from voxelrs_b200.Svo, on host memory."""

    def __init__(self, pkg, ora, sharded, world_buffer, reg, width, height, fmt=0, head=24):
        self.pkg, self.ora, self.sharded, self.reg, self.fmt, self.head = pkg, ora, sharded, reg, fmt, head
        self.world_buffer = world_buffer          # this rank's replica of the GPU world buffer (uint8)
        self.frame = np.zeros((height, width, 4), np.float32)
        self.width, self.height = width, height

    @staticmethod
    def _view(ptr, n):
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr))

    def commit_packed_device(self, ptr, n_ranges, payload_bytes, used_bytes, depth):
        packed = self._view(ptr, 16 * n_ranges + payload_bytes)
        consumed, skipped = self.sharded.apply_packed_host(self.world_buffer, packed, n_ranges, head=self.head)
        assert consumed == 16 * n_ranges + payload_bytes and skipped == 0

    def render_raw(self, vx_params, width, height, shard=None):
        tex, mips = self.reg.textures()
        scene = self.ora.Scene(self.world_buffer, self.reg.materials().tobytes(), tex, mips, fmt=self.fmt)
        full, _ = scene.render(vx_params, width, height)
        tiles = self.sharded.TileShards(width, height, shard[1])
        mine = tiles.owner_map() == shard[0]
        self.frame[mine] = full[mine]

    def pack_shard(self, shard, ptr):
        tiles = self.sharded.TileShards(self.width, self.height, shard[1])
        out = self._view(ptr, tiles.shard_bytes(shard[0]))
        out[:] = tiles.pack(self.frame, shard[0]).view(np.uint8).reshape(-1)

    def unpack_shard(self, shard, ptr):
        tiles = self.sharded.TileShards(self.width, self.height, shard[1])
        tiles.unpack(self.frame, shard[0], self._view(ptr, tiles.shard_bytes(shard[0])).copy())


def _worker(rank, world_size, port, result_path, fmt=0):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    import helpers
    pkg, ora = g.load_pkg(), g.load_oracle()
    import importlib.util
    spec = importlib.util.spec_from_file_location("vx_sharded", os.path.join(ROOT, "voxel-rs_b200", "sharded.py"))
    sharded = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sharded)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world_size)
    try:
        reg = pkg.content_registry(pkg.load_atlas())
        world = pkg.World(radius=2, center=(-1, 2, 5), seed=1, fmt=fmt)
        world.generate(0, 8)
        world.serialize()
        HB = world.header_bytes                          # 24 ESVO, 8 CSVO
        capacity = world.size_bytes + HB + (1 << 20)
        replica = np.zeros(capacity, np.uint8)
        base = world.gpu_buffer()
        replica[:len(base)] = base                       # every rank starts from the same committed SVO
        world.dirty_ranges()

        # the edit happens on rank 0 ONLY; rank 1 learns about it through the broadcast
        n_ranges = payload = used = depth = 0
        packed_host = None
        if rank == 0:
            hgt = world.height_at(-20, 174)
            for dy in range(1, 7):
                world.edit_block(-20, hgt + dy, 174, 4)
            world.serialize()
            ranges = world.dirty_ranges() or []
            new = world.gpu_buffer()
            mirror = np.zeros(capacity, np.uint8)
            mirror[:len(new)] = new
            if not ranges:   # dirty list already drained by serialize(): fall back to a diff of the two images
                diff = np.nonzero(new[HB:len(base)] != base[HB:])[0]
                lo, hi = int(diff.min()) // 4 * 4, (int(diff.max()) // 4 + 1) * 4
                ranges = [(lo, hi - lo)]
                if len(new) > len(base):
                    ranges.append((len(base) - HB, len(new) - len(base)))
            if fmt == 1:     # CSVO ranges are byte-granular: at least one of them must be unaligned for this test to mean anything
                assert any(o % 4 or l % 4 for o, l in ranges), ranges
            packed = sharded.pack_dirty_host(mirror, ranges, head=HB)
            packed_host = torch.from_numpy(packed.copy())
            n_ranges, payload, used, depth = len(ranges), len(packed) - 16 * len(ranges), world.size_bytes, world.depth
        meta = [(n_ranges, payload, used, depth)]
        dist.broadcast_object_list(meta, src=0)
        n_ranges, payload, used, depth = meta[0]

        engine = HostEngine(pkg, ora, sharded, replica, reg, W, H, fmt=fmt, head=HB)
        sf = sharded.ShardedFrame(engine, rank, world_size, dist=dist, torch=torch, device=torch.device("cpu"), gather="nccl")
        sf.configure(W, H, max_dirty_bytes=16 * n_ranges + payload)
        sf.broadcast_dirty(n_ranges, payload, used, depth, packed_host=packed_host)

        # (1) replicas agree byte for byte with rank 0's re-serialized buffer
        want = [world.gpu_buffer() if rank == 0 else None]
        dist.broadcast_object_list(want, src=0)
        assert replica[:len(want[0])].tobytes() == want[0].tobytes(), f"rank {rank}: replica differs after the dirty broadcast"

        import ctypes
        p = pkg.render_params(cam_pos=(-24.0, 80.0, 174.0), cam_fwd=(1.0, -0.3, 0.0), fov_y_deg=72.0, aspect=W / H)
        q = pkg.VxhRenderParams.from_buffer_copy(bytes(p))
        q.cam_pos = (ctypes.c_float * 3)(*world.cnv_block_pos(tuple(p.cam_pos)))
        vxp = pkg.to_vx_render_params(q)
        sf.render(vxp)
        sf.finish()
        sf.release()
        if rank == 0:
            tex, mips = reg.textures()
            full, _ = ora.Scene(replica, reg.materials().tobytes(), tex, mips, fmt=fmt).render(vxp, W, H)
            # (2) gathered frame == unsharded frame
            assert engine.frame.tobytes() == full.tobytes()
            assert np.isfinite(full).all() and full[..., :3].std() > 0.01
            with open(result_path, "w") as f:
                f.write("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("fmt", [0, 1], ids=["esvo", "csvo"])
def test_two_ranks_over_gloo(pkg, ora, tmp_path, fmt):
    import torch.multiprocessing as mp
    result = tmp_path / "rank0.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(result), fmt), nprocs=2, join=True)
    assert result.read_text() == "ok"


def test_shard_geometry(pkg):
    import importlib.util
    spec = importlib.util.spec_from_file_location("vx_sharded", os.path.join(ROOT, "voxel-rs_b200", "sharded.py"))
    sharded = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sharded)
    L = pkg.lib()
    for (w, h) in ((3840, 2160), (333, 190), (32, 16), (1, 1), (7680, 4320)):
        for n in (1, 2, 3, 4, 8):
            t = sharded.TileShards(w, h, n)
            owners = np.concatenate([t.owned(r) for r in range(n)])
            assert sorted(owners.tolist()) == list(range(t.n_macros))               # disjoint, complete
            sizes = [len(t.owned(r)) for r in range(n)]
            assert max(sizes) - min(sizes) <= 1                                       # balanced to one macro block
            for r in range(n):
                sh = pkg.VxShard(r, n)
                assert L.vx_shard_bytes(w, h, C.byref(sh)) == t.shard_bytes(r)        # the C ABI agrees with the layout
    # pack / unpack round trip on a ragged frame
    rng = np.random.default_rng(0)
    frame = rng.random((70, 100, 4), dtype=np.float32)
    t = sharded.TileShards(100, 70, 3)
    out = np.zeros_like(frame)
    for r in range(3):
        t.unpack(out, r, t.pack(frame, r))
    assert out.tobytes() == frame.tobytes()
