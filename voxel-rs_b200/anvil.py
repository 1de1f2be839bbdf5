"""Minecraft Anvil (.mca) world reader -> engine chunks: the producer behind the reference's `--mc-world` option.

Mirrors `systems::storage::MinecraftStorage` (src/systems/storage.rs:50-160, built on the crates fastanvil 0.31 / fastnbt 2.5,
which are NOT part of the reference checkout): every chunk of every `r.X.Z.mca` file of a region directory is pre-loaded;
engine chunk (cx, cy, cz) (32^3 voxels) is assembled from the 2 x 2 Minecraft chunk columns (16 x 16) at
(2 cx + {0,1}, 2 cz + {0,1}) and the world heights 32 cy .. 32 cy + 31; block names map to the engine's BlockIds with the rules
of storage.rs:126-151.

Restated from the published formats (Region file: 8 KiB header of 1024 big-endian {offset:3, sectors:1} entries, chunks =
u32 length, u8 compression (2 = zlib), payload; NBT; chunk format of Minecraft 1.18+: `sections[].Y`,
`sections[].block_states.palette[].Name`, `sections[].block_states.data` = packed palette indices, max(4, ceil(log2(n)))
bits each, not spanning 64-bit words). Host-side Python only — world content, not the ray-cast path; parity with fastanvil is
unpinned (the reference has no test for it).
"""
import os
import struct
import zlib

import numpy as np

# src/gamelogic/content.rs:6-18
AIR, GRASS, DIRT, STONE, STONE_BRICKS, GLASS, GRAVEL, SAND, WATER, OAK_LOG, OAK_LEAVES, OAK_PLANKS, COBBLESTONE = range(13)


def block_id(name):
    """storage.rs:126-151: Minecraft block name -> engine BlockId (0 = nothing)."""
    if "_ore" in name:
        return AIR
    if "_leaves" in name:
        return OAK_LEAVES
    if "_log" in name:
        return OAK_LOG
    if "_planks" in name:
        return OAK_PLANKS
    return {
        "minecraft:dirt": DIRT, "minecraft:grass_block": GRASS, "minecraft:gravel": GRAVEL, "minecraft:clay": GRAVEL,
        "minecraft:sand": SAND, "minecraft:sandstone": SAND, "minecraft:water": WATER,
        "minecraft:stone": STONE, "minecraft:andesite": STONE, "minecraft:diorite": STONE, "minecraft:deepslate": STONE,
        "minecraft:tuff": STONE, "minecraft:granite": STONE, "minecraft:cobblestone": COBBLESTONE,
    }.get(name, AIR)


# ------------------------------------------------------------------------------------------------------------- NBT --

class _Reader:
    def __init__(self, data):
        self.d, self.i = data, 0

    def take(self, fmt):
        v = struct.unpack_from(fmt, self.d, self.i)
        self.i += struct.calcsize(fmt)
        return v[0]

    def string(self):
        n = self.take(">H")
        s = self.d[self.i:self.i + n].decode("utf-8", errors="replace")
        self.i += n
        return s

    def payload(self, tag):
        if tag == 1: return self.take(">b")
        if tag == 2: return self.take(">h")
        if tag == 3: return self.take(">i")
        if tag == 4: return self.take(">q")
        if tag == 5: return self.take(">f")
        if tag == 6: return self.take(">d")
        if tag == 7:
            n = self.take(">i"); v = self.d[self.i:self.i + n]; self.i += n; return v
        if tag == 8: return self.string()
        if tag == 9:
            t = self.take(">b"); n = self.take(">i")
            return [self.payload(t) for _ in range(n)]
        if tag == 10:
            out = {}
            while True:
                t = self.take(">b")
                if t == 0:
                    return out
                name = self.string()
                out[name] = self.payload(t)
        if tag == 11:
            n = self.take(">i"); v = np.frombuffer(self.d, dtype=">i4", count=n, offset=self.i); self.i += 4 * n; return v
        if tag == 12:
            n = self.take(">i"); v = np.frombuffer(self.d, dtype=">u8", count=n, offset=self.i); self.i += 8 * n; return v
        raise ValueError(f"NBT tag {tag}")


def parse_nbt(data):
    r = _Reader(data)
    tag = r.take(">b")
    r.string()
    return r.payload(tag)


# ---------------------------------------------------------------------------------------------------------- region --

def read_region(path):
    """Yields (chunk_x_in_region, chunk_z_in_region, nbt dict) for every chunk stored in the file (Region::iter)."""
    d = open(path, "rb").read()
    if len(d) < 8192:
        return
    for i in range(1024):
        off = int.from_bytes(d[4 * i:4 * i + 3], "big")
        if off == 0 or off * 4096 + 5 > len(d):
            continue
        p = off * 4096
        length, comp = struct.unpack_from(">IB", d, p)
        raw = d[p + 5:p + 4 + length]
        try:
            if comp == 2:
                raw = zlib.decompress(raw)
            elif comp == 1:
                import gzip
                raw = gzip.decompress(raw)
            elif comp != 3:
                continue
            yield i % 32, i // 32, parse_nbt(raw)
        except Exception:
            continue   # storage.rs:78-80: chunks that fail to load are skipped


class JavaChunk:
    """One Minecraft chunk column: block(x, y, z) -> engine BlockId (fastanvil CurrentJavaChunk::block + the name mapping)."""

    def __init__(self, nbt):
        self.sections = {}
        for s in nbt.get("sections", []):
            bs = s.get("block_states")
            if not bs or "palette" not in bs:
                continue
            ids = np.array([block_id(p.get("Name", "")) for p in bs["palette"]], dtype=np.uint8)
            if len(ids) == 1 or "data" not in bs:
                if ids[0] == AIR:
                    continue
                blocks = np.full(4096, ids[0], dtype=np.uint8)
            else:
                bits = max(4, int(len(ids) - 1).bit_length())
                per = 64 // bits
                words = np.asarray(bs["data"], dtype=np.uint64)
                shifts = (np.arange(per, dtype=np.uint64) * np.uint64(bits))
                idx = ((words[:, None] >> shifts[None, :]) & np.uint64((1 << bits) - 1)).reshape(-1)[:4096].astype(np.int64)
                blocks = ids[np.minimum(idx, len(ids) - 1)]
            if blocks.any():
                self.sections[int(s["Y"])] = blocks.reshape(16, 16, 16)   # [y][z][x]

    def column(self, y0):
        """16 x 16 blocks for world heights y0 .. y0 + 31 as [y][z][x] (32, 16, 16), or None if nothing is there."""
        out = None
        for k in range(2):
            sec = self.sections.get((y0 >> 4) + k)
            if sec is not None:
                if out is None:
                    out = np.zeros((32, 16, 16), dtype=np.uint8)
                out[16 * k:16 * k + 16] = sec
        return out


class MinecraftWorld:
    """MinecraftStorage::new + load (storage.rs:58-160)."""

    def __init__(self, region_dir):
        self.chunks = {}
        for fn in sorted(os.listdir(region_dir)):
            parts = fn.split(".")
            if len(parts) != 4 or parts[0] != "r" or parts[3] != "mca":
                continue
            rx, rz = int(parts[1]), int(parts[2])
            for cx, cz, nbt in read_region(os.path.join(region_dir, fn)):
                jc = JavaChunk(nbt)
                if jc.sections:
                    self.chunks[(rx * 32 + cx, rz * 32 + cz)] = jc

    def engine_chunk(self, cx, cy, cz):
        """Dense 32^3 uint32 block array (index x + 32*(y + 32*z)) of engine chunk (cx, cy, cz), or None if empty."""
        out = None
        for dz in range(2):
            for dx in range(2):
                jc = self.chunks.get((2 * cx + dx, 2 * cz + dz))
                if jc is None:
                    continue
                col = jc.column(32 * cy)
                if col is None:
                    continue
                if out is None:
                    out = np.zeros((32, 32, 32), dtype=np.uint8)   # [z][y][x]
                out[16 * dz:16 * dz + 16, :, 16 * dx:16 * dx + 16] = col.transpose(1, 0, 2)
        if out is None or not out.any():
            return None
        return out.reshape(-1).astype(np.uint32)
