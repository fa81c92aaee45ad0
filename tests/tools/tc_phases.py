"""Per-phase clock counts of the tensor-core blur kernel (CTA 0, thread 0) on a big image.  Developer tool, run on a B200:
    python tests/tools/tc_phases.py [k] [mode]      mode: blur | sub | ink"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from smart_image_processing_b200 import DocScanner as DS  # noqa: E402
from smart_image_processing_b200 import ops  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 23
mode = sys.argv[2] if len(sys.argv) > 2 else "blur"
rng = np.random.default_rng(0)
img = rng.integers(0, 256, (8000, 8000), dtype=np.uint8)
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
st = d[40960:40960 + 64 * 32].reshape(64, 16, 2).astype(np.uint64)
clk_all = (st[:, :, 0] | (st[:, :, 1] << np.uint64(32))).astype(np.int64)
clk = clk_all[:, :11]
valid = [i for i in range(64) if 0 < clk[i, 0] < clk[i, 10] < (1 << 62)]
names = ["top->consts", "consts+prefetch->S landed", "S->MMA1 issued", "MMA1 issued->D1 ready", "D1 drain", "sync", "MMA2 issue",
         "MMA2->D2 ready", "epilogue", "end sync"]
print(f"k={k} mode={mode} tiles seen by CTA 0: {len(valid)}")
rows = np.array([np.diff(clk[i]) for i in valid[1:]])
for j, nm in enumerate(names):
    print(f"  {nm:28s} mean {rows[:, j].mean():9.0f}  median {int(np.median(rows[:, j])):7d}  min {rows[:, j].min():7d}  max {rows[:, j].max():7d}")
print(f"  tile total                   mean {np.mean([clk[i, 10] - clk[i, 0] for i in valid[1:]]):9.0f}")
print(f"  loop top (stamp11) -> stamp0  mean {np.mean([clk_all[i, 0] - clk_all[i, 11] for i in valid[1:]]):9.0f}")
print(f"  end sync -> loop bottom (12)  mean {np.mean([clk_all[i, 12] - clk_all[i, 10] for i in valid[1:]]):9.0f}")
print(f"  bottom -> next top            mean {np.mean([clk_all[valid[n + 1], 11] - clk_all[valid[n], 12] for n in range(len(valid) - 1)]):9.0f}")
print(f"  tile-to-tile                 mean {np.mean(np.diff([clk[i, 0] for i in valid])):9.0f}")
