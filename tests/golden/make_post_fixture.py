"""Writes tests/golden/post_sequence.npz: a seeded synthetic tile-framebuffer sequence and what the image-space
oracle (oracle/vrt_post_oracle.c) makes of it.  A REGRESSION fixture of our own canonical arithmetic — the reference's
GLSL cannot run in this image, so it is not an output of the reference (parity unpinned for this step)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pypostoracle as pp  # noqa: E402
from scenes import gbuffer_synth as pu  # noqa: E402

W, H, FRAMES, PASSES, SEED = 48, 32, 5, 5, 4242
seq = pu.synthetic_sequence(W, H, FRAMES, seed=SEED)
orc = pp.PostOracle(W, H)
orc.set_passes(PASSES)
rgba, tiles_all = [], []
for f, (proj, inv, pos, tiles) in enumerate(seq):
    orc.set_camera(proj, inv, pos, reset_history=(f == 3))
    rgba.append(orc.denoise_present(tiles))
    tiles_all.append(tiles.view(np.uint32).copy())
np.savez_compressed(
    Path(__file__).resolve().parent / "post_sequence.npz",
    w=W, h=H, frames=FRAMES, passes=PASSES, seed=SEED,
    tiles=np.stack(tiles_all), rgba=np.stack(rgba),
    hist=orc.read(orc.HIST), moments=orc.read(orc.MOMENTS), prev_irr=orc.read(orc.PREV_IRR),
)
print("wrote post_sequence.npz")
