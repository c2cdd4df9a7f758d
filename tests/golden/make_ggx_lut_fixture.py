"""Extracts the R,G channels of the reference's split-sum LUT into a small fixture.

Run in the build container only (reads /root/reference/ggx_lut.png, which does
not exist on the GPU box).  The reference `include_bytes!`s this PNG and uploads
it as R8G8B8A8_UNORM (src/main.rs:295-330); only .xy is ever sampled
(shader/src/lib.rs:126-133), B == 0 and A == 255 everywhere, so the fixture
keeps R and G and the loader re-expands to RGBA8.
"""
import hashlib
import sys

import numpy as np
from PIL import Image

SRC = "/root/reference/ggx_lut.png"
EXPECT_SHA256 = "c7e47e1c5df98c3450ba22391e2a0e14129ab3db8919027d8a4cec085bad7bc3"

raw = open(SRC, "rb").read()
sha = hashlib.sha256(raw).hexdigest()
assert sha == EXPECT_SHA256, sha
img = np.array(Image.open(SRC).convert("RGBA"))
assert img.shape == (1024, 1024, 4)
assert (img[..., 2] == 0).all() and (img[..., 3] == 255).all()
out = sys.argv[1] if len(sys.argv) > 1 else "tests/golden/ggx_lut_rg.npz"
np.savez_compressed(out, rg=img[..., :2], source_sha256=np.array(sha))
print("wrote", out, img[0, 0], img[0, 1023], img[1023, 0], img[1023, 1023])
