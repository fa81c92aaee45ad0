"""The tensor-core adaptive threshold against the C oracle on a few shapes, with the guard-band statistics.
Developer tool, run on a B200:   python tests/tools/tc_adaptive_probe.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from smart_image_processing_b200 import ops  # noqa: E402


def page_like(rng, h, w):
    im = np.full((h, w), 205.0, np.float32)
    for i in range(max(1, h // 14)):
        y, x = 4 + 14 * i, 4
        while x < w - 12:
            ww = int(rng.integers(4, 40))
            im[y:y + 7, x:x + ww] = rng.integers(15, 95)
            x += ww + int(rng.integers(3, 14))
    im = im * (0.55 + 0.45 * np.linspace(0, 1, w)[None, :]) + rng.normal(0, 3, (h, w))
    return np.clip(im, 0, 255).astype(np.uint8)


rng = np.random.default_rng(3)
os.environ["DOCSCAN_TC_DEBUG"] = tempfile.NamedTemporaryFile(suffix=".bin", delete=False).name
bad_total = 0
for (h, w) in [(300, 260), (128, 128), (97, 131), (1600, 1131), (257, 1031), (33, 70)]:
    for k, c in [(35, 10), (31, 3), (11, 10), (3, 0), (61, -2)]:
        for kind in ("page", "noise"):
            g = page_like(rng, max(h, 16), max(w, 16))[:h, :w] if kind == "page" else rng.integers(0, 256, (h, w), dtype=np.uint8)
            got = ops.adaptive_threshold(g, "gaussian", k, c)
            want = O.adaptive_threshold(g, "gaussian", k, c)
            bad = int(np.count_nonzero(got != want))
            bad_total += bad
            print(f"{h}x{w} k={k} C={c} {kind}: {bad} differ")
print("TOTAL BAD", bad_total)
