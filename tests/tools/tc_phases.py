"""Per-phase clock counts of the tensor-core blur kernel (CTA 0: epilogue warp 0 and the issuing warp) on a big image.
Developer tool, run on a B200:   python tests/tools/tc_phases.py [k] [mode]      mode: blur | sub | ink"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from smart_image_processing_b200 import DocScanner as DS  # noqa: E402
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


k = int(sys.argv[1]) if len(sys.argv) > 1 else 23
mode = sys.argv[2] if len(sys.argv) > 2 else "blur"
rng = np.random.default_rng(0)
kind = sys.argv[3] if len(sys.argv) > 3 else "page"
img = np.tile(page_like(rng, 2000, 2000), (4, 4)) if kind == "page" else np.clip(rng.normal(200, 3, (8000, 8000)), 0, 255).astype(np.uint8)
ops.gaussian_blur(img, k)                                # warm-up (tables, arena)
path = tempfile.NamedTemporaryFile(suffix=".bin", delete=False).name
os.environ["DOCSCAN_TC_DEBUG"] = path
if mode == "blur":
    ops.gaussian_blur(img, k)
elif mode == "sub":
    DS.illumination_correction(img, "subtract", (k - 0.4) / 8000.0)
else:
    DS._compute_ink_mask(img, mask_blur_ksize=k)
del os.environ["DOCSCAN_TC_DEBUG"]
d = np.fromfile(path, np.uint32)
print(f"k={k} mode={mode} image={kind}")
for who, names in ((0, ["top->D1 ready", "drain+arrive", "wait D2", "epilogue units", "hist flush"]),
                   (1, ["ensure matrices", "wait A2 (16 warps)", "pass 2 issue", "wait S + pass 1 issue", "next TMA"])):
    st = d[40960 + who * 2048:40960 + (who + 1) * 2048].reshape(64, 16, 2).astype(np.uint64)
    clk = (st[:, :, 0] | (st[:, :, 1] << np.uint64(32))).astype(np.int64)[:, :6]
    valid = [i for i in range(64) if 0 < clk[i, 0] < clk[i, 5] < (1 << 62)]
    rows = np.array([np.diff(clk[i]) for i in valid[2:-1]])
    print(f" {'epilogue warp 0' if who == 0 else 'issuer'}: {len(valid)} tiles")
    for j, nm in enumerate(names):
        print(f"  {nm:26s} mean {rows[:, j].mean():8.0f}  median {int(np.median(rows[:, j])):6d}  min {rows[:, j].min():6d}  max {rows[:, j].max():6d}")
    print(f"  tile-to-tile               mean {np.mean(np.diff([clk[i, 0] for i in valid[2:-1]])):8.0f}")
