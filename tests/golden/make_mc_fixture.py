"""Makes tests/golden/mc_world.npz from the reference's bundled Minecraft world (assets/worlds/benchmark, the 16 region files
that are present in the checkout: SURVEY F12). Run in the build container only (it reads /root/reference):

    python tests/golden/make_mc_fixture.py

The fixture holds the engine chunks (32^3 BlockIds as uint8, index x + 32*(y + 32*z)) of the complete 9 x 8 block of chunk
columns of region r.-5.3 (engine chunks x -73..-65, z 48..55) for the y-chunks 0..2 (world heights 0..95; the terrain there
tops out at y = 95 and the reference only loads y-chunks >= 0, gamelogic/world.rs:85), decoded by voxel-rs_b200/anvil.py with
the block mapping of src/systems/storage.rs:126-151. Sea level (y = 63) runs through it: water, sand, gravel, trees — the
translucent materials the generated terrain never contains.
"""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location("anvil", os.path.join(ROOT, "voxel-rs_b200", "anvil.py"))
anvil = importlib.util.module_from_spec(spec)
spec.loader.exec_module(anvil)

CX, CZ, CY = range(-73, -64), range(48, 56), range(0, 3)

if __name__ == "__main__":
    w = anvil.MinecraftWorld("/root/reference/assets/worlds/benchmark")
    coords, blocks = [], []
    for cx in CX:
        for cz in CZ:
            for cy in CY:
                b = w.engine_chunk(cx, cy, cz)
                if b is None:
                    continue
                coords.append((cx, cy, cz))
                blocks.append(b.astype(np.uint8))
    coords = np.array(coords, dtype=np.int32)
    blocks = np.stack(blocks)
    out = os.path.join(ROOT, "tests", "golden", "mc_world.npz")
    np.savez_compressed(out, coords=coords, blocks=blocks)
    ids, counts = np.unique(blocks, return_counts=True)
    print(f"{len(coords)} chunks, {os.path.getsize(out) / 1e6:.2f} MB, block histogram {dict(zip(ids.tolist(), counts.tolist()))}")
