"""Multi-GPU frame: image-space tile shards over the ranks of one box, SVO replicated.

One process per GPU (torch.distributed). Per frame:

  1. rank 0 packs the frame's dirty SVO ranges (vx_svo_pack_dirty) -> one broadcast (NCCL over NVLink) -> every rank applies
     them to its replica with a scatter kernel (vx_svo_commit_packed_device);
  2. every rank renders the macro blocks it owns (macro block m belongs to rank m % world_size);
  3. the tiles end up in GPU 0's framebuffer, either
       gather="p2p"   fused: non-root ranks map GPU 0's framebuffer (CUDA IPC) and their shade / shadow kernels store the
                      finished pixels straight into it over NVLink — no pack, no collective, only a barrier at the end;
       gather="nccl"  pack the shard, grouped NCCL send/recv to rank 0, unpack (baseline; also what the gloo CPU test runs).

The reference has no counterpart (single GPU). This module is host-side plumbing only: it never touches pixels
itself. `engine` is the object that does the device work — voxelrs_b200.Svo on a GPU; the CPU tests pass a stand-in with
the same methods so the collectives, the packed-range format and the shard layout run under gloo with world_size 2.

Layouts (shared with the CUDA kernels in csrc/kernels.cuh and restated in numpy below for tests):
  packed dirty set   n VxRange{u64 offset, u64 length} headers | first 24 bytes of the world buffer | range 0 bytes | range 1 bytes ...
  packed shard       for each owned macro block, in increasing macro index: 16 rows x 32 pixels RGBA32F (zero outside the frame)
"""
import numpy as np

MACRO_W, MACRO_H = 32, 16


class TileShards:
    """Shard geometry: which 32x16-pixel macro blocks a rank owns, and the packed-shard layout (numpy restatement)."""

    def __init__(self, width, height, world_size):
        self.width, self.height, self.world_size = width, height, world_size
        self.macro_x = (width + MACRO_W - 1) // MACRO_W
        self.macro_y = (height + MACRO_H - 1) // MACRO_H
        self.n_macros = self.macro_x * self.macro_y

    def owned(self, rank):
        return np.arange(rank, self.n_macros, self.world_size)

    def shard_bytes(self, rank):
        return len(self.owned(rank)) * MACRO_W * MACRO_H * 16

    def owner_map(self):
        """(height, width) array of the rank that renders each pixel."""
        ys, xs = np.mgrid[0:self.height, 0:self.width]
        return ((ys // MACRO_H) * self.macro_x + xs // MACRO_W) % self.world_size

    def pack(self, frame, rank):
        """frame: (height, width, 4) float32 -> packed shard (n_owned, 16, 32, 4)."""
        out = np.zeros((len(self.owned(rank)), MACRO_H, MACRO_W, 4), np.float32)
        for k, m in enumerate(self.owned(rank)):
            x0, y0 = (m % self.macro_x) * MACRO_W, (m // self.macro_x) * MACRO_H
            blk = frame[y0:y0 + MACRO_H, x0:x0 + MACRO_W]
            out[k, :blk.shape[0], :blk.shape[1]] = blk
        return out

    def unpack(self, frame, rank, packed):
        packed = np.asarray(packed).view(np.float32).reshape(-1, MACRO_H, MACRO_W, 4)
        for k, m in enumerate(self.owned(rank)):
            x0, y0 = (m % self.macro_x) * MACRO_W, (m // self.macro_x) * MACRO_H
            h, w = min(MACRO_H, self.height - y0), min(MACRO_W, self.width - x0)
            frame[y0:y0 + h, x0:x0 + w] = packed[k, :h, :w]


def pack_dirty_host(world_buffer, ranges):
    """numpy restatement of vx_svo_pack_dirty. world_buffer: uint8 GPU-buffer image (byte 0 = octree_scale)."""
    hdr = np.array([[o, l] for o, l in ranges], dtype=np.uint64).reshape(-1, 2)
    parts = [hdr.view(np.uint8).reshape(-1), world_buffer[:24]] + [world_buffer[24 + o:24 + o + l] for o, l in ranges]
    return np.concatenate(parts)


def apply_packed_host(world_buffer, packed, n_ranges):
    """numpy restatement of scatter_ranges_kernel: applies a packed dirty set to a replica's buffer in place."""
    packed = np.asarray(packed, dtype=np.uint8)
    hdr = packed[:16 * n_ranges].view(np.uint64).reshape(-1, 2)
    world_buffer[:24] = packed[16 * n_ranges:16 * n_ranges + 24]
    off = 16 * n_ranges + 24
    for o, l in hdr:
        o, l = int(o), int(l)
        world_buffer[24 + o:24 + o + l] = packed[off:off + l]
        off += l
    return off


class ShardedFrame:
    """One rank's side of the sharded frame. `dist` is torch.distributed (already initialised) or None for world_size 1."""

    def __init__(self, engine, rank, world_size, dist=None, torch=None, device=None, gather="p2p"):
        self.engine, self.rank, self.world_size, self.dist, self.torch, self.device = engine, rank, world_size, dist, torch, device
        self.gather = gather if world_size > 1 else "none"
        self.shard = (rank, world_size)
        self.width = self.height = 0
        self._peer_open = False
        self._flag = None
        self._frame_open = False   # a collective since the last finish() already ordered this frame after rank 0's readers

    # ---- setup -----------------------------------------------------------------------------------------------------------
    def configure(self, width, height, max_dirty_bytes):
        """Collective. Sizes the staging buffers; in p2p mode exchanges GPU 0's framebuffer handle and maps it."""
        t = self.torch
        self.width, self.height = width, height
        self.tiles = TileShards(width, height, self.world_size)
        if self.world_size == 1:
            return
        self.packed_dirty = t.empty(max_dirty_bytes, dtype=t.uint8, device=self.device)
        self._flag = t.zeros(1, dtype=t.int32, device=self.device)
        if self.gather == "p2p":
            box = [self.engine.frame_ipc_handle() if self.rank == 0 else None]
            self.dist.broadcast_object_list(box, src=0)
            if self.rank != 0:
                self.engine.open_peer_frame(box[0])
                self._peer_open = True
        else:
            self.my_pack = t.empty(self.tiles.shard_bytes(self.rank), dtype=t.uint8, device=self.device)
            self.recv = [t.empty(self.tiles.shard_bytes(r), dtype=t.uint8, device=self.device) for r in range(self.world_size)] \
                if self.rank == 0 else None

    def close(self):
        if self._peer_open:
            self.engine.close_peer_frame()
            self._peer_open = False

    # ---- per frame -------------------------------------------------------------------------------------------------------
    def broadcast_dirty(self, n_ranges, payload_bytes, used_bytes, depth, packed_host=None):
        """Collective. Rank 0's packed dirty set (already in self.packed_dirty, or given as a pinned host tensor) -> every
        replica, applied by the scatter kernel on the render stream's upload side."""
        if self.world_size == 1:
            return
        total = 16 * n_ranges + payload_bytes
        buf = self.packed_dirty[:total]
        if self.rank == 0 and packed_host is not None:
            buf.copy_(packed_host[:total], non_blocking=True)
        self.dist.broadcast(buf, src=0)
        self._frame_open = True
        self.engine.commit_packed_device(buf.data_ptr(), n_ranges, payload_bytes, used_bytes, depth)

    def render(self, vx_params):
        """This rank's tiles. In p2p mode the pixels land in GPU 0's framebuffer as they are finished, so the frame may only
        start once rank 0 is done reading the previous one: the dirty-range broadcast (root = rank 0) orders that; a frame
        without one takes a 4-byte all-reduce instead."""
        if self.gather == "p2p" and not self._frame_open:
            self.dist.all_reduce(self._flag)
        self._frame_open = False
        self.engine.render_raw(vx_params, self.width, self.height, shard=self.shard)

    def finish(self):
        """Collective. After it returns (stream-ordered), GPU 0's framebuffer holds the whole frame."""
        if self.world_size == 1:
            return
        if self.gather == "p2p":
            # the stores were issued by the render kernels themselves; one 4-byte all-reduce orders every rank's render
            # stream before rank 0 reads the frame
            self.dist.all_reduce(self._flag)
            return
        self.engine.pack_shard(self.shard, self.my_pack.data_ptr())
        if self.rank == 0:
            ops = [self.dist.P2POp(self.dist.irecv, self.recv[r], r) for r in range(1, self.world_size)]
        else:
            ops = [self.dist.P2POp(self.dist.isend, self.my_pack, 0)]
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()
        if self.rank == 0:
            for r in range(1, self.world_size):
                self.engine.unpack_shard((r, self.world_size), self.recv[r].data_ptr())
