"""Freeze the decoded texels of the reference's assets/earthmap.jpg as a fixture.

Runs in the build container only (needs /root/reference); the GPU box uses the committed .npz.
The reference decodes with the `image` 0.24.5 / `jpeg-decoder` 0.3.0 crates (src/textures/image_texture.rs:20);
those are not available here, so PIL (libjpeg-turbo) is used — SURVEY.md §8c: PIL and OpenCV agree bit-exactly on
this 4:4:4 baseline JPEG; a different IDCT may differ by +-1 LSB ("parity unpinned").
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/earthmap.jpg"
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hyper-ray-tracer_b200", "assets", "earthmap_rgb8.npz")
rgb = np.asarray(Image.open(src).convert("RGB"), dtype=np.uint8)
assert rgb.shape == (512, 1024, 3), rgb.shape
np.savez_compressed(dst, rgb=rgb)
print("wrote", os.path.normpath(dst), rgb.shape, "sha1(decoded)=", hashlib.sha1(rgb.tobytes()).hexdigest())
