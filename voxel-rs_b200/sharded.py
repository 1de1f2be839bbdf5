"""Multi-GPU frame: image-space tile shards over the ranks of one box, SVO replicated.

One process per GPU (torch.distributed). Per frame:

  1. rank 0 packs the frame's dirty SVO ranges (vx_svo_pack_dirty) -> one broadcast (NCCL over NVLink) -> every rank applies
     them to its replica with a scatter kernel (vx_svo_commit_packed_device);
  2. every rank renders the macro blocks it owns (macro block m belongs to rank m % world_size);
  3. the tiles end up in GPU 0's framebuffer, either
       gather="p2p"   fused: non-root ranks map GPU 0's framebuffer (CUDA IPC) and their shade / shadow kernels store the
                      finished pixels straight into it over NVLink — no pack, no collective, only a barrier at the end;
       gather="p2p8"  the same with RGBA8 pixels (vx_set_option 8): what Framebuffer::read_pixels hands out anyway, 4 instead of
                      16 bytes per pixel into GPU 0 — at 8 GPUs the RGBA32F gather is bound by GPU 0's NVLink ingress;
       gather="nccl"  pack the shard, grouped NCCL send/recv to rank 0, unpack (baseline; also what the gloo CPU test runs).

The reference has no counterpart (single GPU). This module is host-side plumbing only: it never touches pixels
itself. `engine` is the object that does the device work — voxelrs_b200.Svo on a GPU; the CPU tests pass a stand-in with
the same methods so the collectives, the packed-range format and the shard layout run under gloo with world_size 2.

Layouts (shared with the CUDA kernels in csrc/kernels.cuh and restated in numpy below for tests):
  packed dirty set   n VxRange{u64 offset, u64 length} headers | head bytes of the world buffer (24 ESVO / 8 CSVO) | range 0 bytes | range 1 bytes ...
                     (byte-packed: CSVO ranges sit at arbitrary byte offsets)
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


def pack_dirty_host(world_buffer, ranges, head=24):
    """numpy restatement of vx_svo_pack_dirty. world_buffer: uint8 GPU-buffer image (byte 0 = octree_scale); head = bytes in front
    of the RangeBuffer image: 24 for ESVO (scale + preamble), 8 for CSVO (scale + root pointer) — World.header_bytes."""
    hdr = np.array([[o, l] for o, l in ranges], dtype=np.uint64).reshape(-1, 2)
    parts = [hdr.view(np.uint8).reshape(-1), world_buffer[:head]] + [world_buffer[head + o:head + o + l] for o, l in ranges]
    return np.concatenate(parts)


def apply_packed_host(world_buffer, packed, n_ranges, head=24):
    """numpy restatement of scatter_ranges_kernel: applies a packed dirty set to a replica's buffer in place. Byte-granular (CSVO
    ranges are not word-aligned); a range that would leave the buffer is skipped. Returns (bytes consumed, ranges skipped)."""
    packed = np.asarray(packed, dtype=np.uint8)
    hdr = packed[:16 * n_ranges].view(np.uint64).reshape(-1, 2)
    world_buffer[:head] = packed[16 * n_ranges:16 * n_ranges + head]
    off = 16 * n_ranges + head
    skipped = 0
    for o, l in hdr:
        o, l = int(o), int(l)
        if head + o + l > len(world_buffer):
            skipped += 1
        else:
            world_buffer[head + o:head + o + l] = packed[off:off + l]
        off += l
    return off, skipped


class ShardedFrame:
    """One rank's side of the sharded frame. `dist` is torch.distributed (already initialised) or None for world_size 1.

    Frame protocol, gather="p2p" (flags are 32-bit frame counters in GPU 0's memory, mapped by every rank, slot r = rank r is
    done with the frame, slot 0 = GPU 0 released the frame to be overwritten):
        every rank   apply_dirty (scatter kernel) -> trace primary rays -> [rank != 0: wait flag 0 >= frame-1] -> shade / shadow
                     (pixels stored into GPU 0's framebuffer) -> [rank != 0: signal flag rank = frame]
        rank 0       wait flags 1..n-1 >= frame -> consumer reads the frame -> release(): signal flag 0 = frame
    The dirty ranges of the NEXT frame travel meanwhile: prefetch_dirty() runs the NCCL broadcast on a side stream into one
    of two staging buffers, so it overlaps the current frame instead of sitting in front of the next one.
    """

    def __init__(self, engine, rank, world_size, dist=None, torch=None, device=None, gather="p2p"):
        self.engine, self.rank, self.world_size, self.dist, self.torch, self.device = engine, rank, world_size, dist, torch, device
        self.gather = gather if world_size > 1 else "none"
        self.shard = (rank, world_size)
        self.width = self.height = 0
        self._peer_open = False
        self.frame_no = 0
        self._cuda = device is not None and getattr(device, "type", "cpu") == "cuda"
        self._staged = []      # FIFO of (buffer index, n_ranges, payload, used, depth) broadcast but not yet applied
        self._next_buf = 0

    # ---- setup -----------------------------------------------------------------------------------------------------------
    def configure(self, width, height, max_dirty_bytes):
        """Collective. Sizes the staging buffers; in p2p mode exchanges GPU 0's framebuffer + flag handles and maps them."""
        t = self.torch
        self.width, self.height = width, height
        self.tiles = TileShards(width, height, self.world_size)
        if self.world_size == 1:
            return
        self.dirty_bufs = [t.empty(max_dirty_bytes, dtype=t.uint8, device=self.device) for _ in range(2)]
        self.packed_dirty = self.dirty_bufs[0]
        if self._cuda:
            self.side = t.cuda.Stream(device=self.device)
            # the library's own streams, as torch streams: the scatter kernel runs on ITS upload stream and the shard pack / unpack
            # kernels on ITS render stream, whatever torch's current stream is — the events below are recorded / waited on exactly those
            self.up = t.cuda.ExternalStream(self.engine.stream(1), device=self.device)
            self.rs = t.cuda.ExternalStream(self.engine.stream(0), device=self.device)
            self.ev_staged = [t.cuda.Event() for _ in range(2)]     # broadcast into buffer k finished
            self.ev_applied = [t.cuda.Event() for _ in range(2)]    # scatter out of buffer k finished
            self.ev_gather = t.cuda.Event()
            for e in self.ev_applied:
                e.record(self.up)
        # a new sequence of frames counts from 1 again: the flags of an earlier ShardedFrame on the same engine must not satisfy it
        if self.rank == 0 and hasattr(self.engine, "frame_flags_reset"):
            self.engine.frame_flags_reset()
        self.dist.barrier()
        if self.gather in ("p2p", "p2p8"):
            eight = self.gather == "p2p8"
            if eight:
                self.engine.set_option(8, 1)       # every rank (rank 0 too) stores RGBA8 pixels: a quarter of the NVLink bytes
            box = [((self.engine.frame8_ipc_handle() if eight else self.engine.frame_ipc_handle()), self.engine.sync_ipc_handle())
                   if self.rank == 0 else None]
            self.dist.broadcast_object_list(box, src=0)
            if self.rank != 0:
                (self.engine.open_peer_frame8 if eight else self.engine.open_peer_frame)(box[0][0])
                self.engine.open_peer_sync(box[0][1])
                self._peer_open = True
        else:
            self.my_pack = t.empty(self.tiles.shard_bytes(self.rank), dtype=t.uint8, device=self.device)
            self.recv = [t.empty(self.tiles.shard_bytes(r), dtype=t.uint8, device=self.device) for r in range(self.world_size)] \
                if self.rank == 0 else None

    def check(self):
        """Raises if a frame-flag wait timed out (lost / slow peer: frames since then may be torn) or the scatter kernel refused a
        dirty range. Synchronises this rank's streams: call it at the end of a sequence of frames, not every frame."""
        if self.world_size == 1 or not self._cuda:
            return
        lost, refused = self.engine.frame_sync_errors(), self.engine.scatter_errors()
        if lost or refused:
            raise RuntimeError(f"ShardedFrame rank {self.rank}: {lost} frame-flag waits timed out, {refused} dirty ranges refused")

    def close(self):
        self.check()
        if self._peer_open:
            self.engine.close_peer_frame()
            self.engine.close_peer_sync()
            self._peer_open = False
        if self.gather == "p2p8":
            self.engine.set_option(8, 0)

    # ---- per frame -------------------------------------------------------------------------------------------------------
    def prefetch_dirty(self, n_ranges, payload_bytes, used_bytes, depth, packed_host=None):
        """Collective. Starts moving rank 0's packed dirty set (a pinned host tensor, or already in self.dirty_bufs[next]) to
        every replica: NCCL broadcast on the side stream into the next staging buffer. apply_dirty() consumes it."""
        if self.world_size == 1:
            return
        t = self.torch
        k = self._next_buf
        self._next_buf ^= 1
        total = 16 * n_ranges + payload_bytes
        buf = self.dirty_bufs[k][:total]
        if self._cuda:
            self.side.wait_event(self.ev_applied[k])          # the scatter that last read this buffer is done
            with t.cuda.stream(self.side):
                if self.rank == 0 and packed_host is not None:
                    buf.copy_(packed_host[:total], non_blocking=True)
                self.dist.broadcast(buf, src=0)
                self.ev_staged[k].record(self.side)
        else:
            if self.rank == 0 and packed_host is not None:
                buf.copy_(packed_host[:total])
            self.dist.broadcast(buf, src=0)
        self._staged.append((k, n_ranges, payload_bytes, used_bytes, depth))

    def apply_dirty(self):
        """Applies the oldest prefetched dirty set to this rank's replica (scatter kernel, stream-ordered before the frame)."""
        if self.world_size == 1 or not self._staged:
            return
        k, n_ranges, payload_bytes, used_bytes, depth = self._staged.pop(0)
        if self._cuda:
            self.up.wait_event(self.ev_staged[k])           # the scatter runs on the library's upload stream
        self.engine.commit_packed_device(self.dirty_bufs[k].data_ptr(), n_ranges, payload_bytes, used_bytes, depth)
        if self._cuda:
            self.ev_applied[k].record(self.up)

    def broadcast_dirty(self, n_ranges, payload_bytes, used_bytes, depth, packed_host=None):
        """Collective. prefetch_dirty + apply_dirty back to back (no overlap with the previous frame)."""
        self.prefetch_dirty(n_ranges, payload_bytes, used_bytes, depth, packed_host=packed_host)
        self.apply_dirty()

    def render(self, vx_params):
        """This rank's tiles. In p2p mode the pixels land in GPU 0's framebuffer as they are finished; shading is gated on
        GPU 0 having released the previous frame (its primary rays are traced meanwhile)."""
        self.frame_no += 1
        if self.gather in ("p2p", "p2p8") and self.rank != 0:
            self.engine.frame_gate(0, self.frame_no - 1)
        self.engine.render_raw(vx_params, self.width, self.height, shard=self.shard)

    def finish(self):
        """After it returns (stream-ordered), GPU 0's framebuffer holds the whole frame. Collective only in nccl mode."""
        if self.world_size == 1:
            return
        if self.gather in ("p2p", "p2p8"):
            if self.rank != 0:
                self.engine.frame_signal(self.rank, self.frame_no)
            else:
                self.engine.frame_wait(1, self.world_size - 1, self.frame_no)
            return
        self.engine.pack_shard(self.shard, self.my_pack.data_ptr())      # on the library's render stream
        cur = None
        if self._cuda:
            cur = self.torch.cuda.current_stream(self.device)
            self.ev_gather.record(self.rs)
            cur.wait_event(self.ev_gather)                               # NCCL (torch's current stream) after the pack
        if self.rank == 0:
            ops = [self.dist.P2POp(self.dist.irecv, self.recv[r], r) for r in range(1, self.world_size)]
        else:
            ops = [self.dist.P2POp(self.dist.isend, self.my_pack, 0)]
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()
        if self._cuda:
            self.ev_gather.record(cur)
            self.rs.wait_event(self.ev_gather)                           # unpack / the next frame's pack after the transfer
        if self.rank == 0:
            for r in range(1, self.world_size):
                self.engine.unpack_shard((r, self.world_size), self.recv[r].data_ptr())

    def release(self):
        """Rank 0, after everything that reads the gathered frame has been enqueued: lets the other ranks overwrite it."""
        if self.gather in ("p2p", "p2p8") and self.rank == 0:
            self.engine.frame_signal(0, self.frame_no)
