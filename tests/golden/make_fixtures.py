"""Generates the binary fixtures under tests/golden/ from the reference checkout (run in the build container,
where /root/reference exists; the GPU box only sees the committed outputs).

  atlas.npz                                the 25 block textures of assets/textures/*.png, decoded RGBA8, row 0 = top
                                           (inputs of the reference's VoxelRegistry, src/gamelogic/content.rs:20-46)
  graphics_svo_render_expected.png         expected image of svo_tests::render (src/graphics/svo.rs:342-399)
  gamelogic_world_end_to_end_expected.png  expected image of tests::end_to_end (src/gamelogic/world.rs:461-498)
"""
import glob
import os
import shutil

import numpy as np
from PIL import Image

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

atlas = {}
for f in sorted(glob.glob(os.path.join(REF, "assets/textures/*.png"))):
    atlas[os.path.splitext(os.path.basename(f))[0]] = np.asarray(Image.open(f).convert("RGBA"), dtype=np.uint8)
np.savez_compressed(os.path.join(OUT, "atlas.npz"), **atlas)
for name in ("graphics_svo_render_expected.png", "gamelogic_world_end_to_end_expected.png"):
    shutil.copyfile(os.path.join(REF, "assets/tests", name), os.path.join(OUT, name))
print("wrote", len(atlas), "textures,", os.path.getsize(os.path.join(OUT, "atlas.npz")), "bytes")
