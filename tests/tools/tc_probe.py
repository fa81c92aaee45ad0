"""Stage-by-stage check of the tensor-core Gaussian (csrc/tcblur.cu) on a B200: the final image against the C oracle, and
the first tile's accumulators (DOCSCAN_TC_DEBUG dump: D1, D2lo, D2hi) against a numpy restatement of the two contractions.
Run on the GPU box:  python tests/tools/tc_probe.py [k ...]"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # tests/tools -> repo root
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from smart_image_processing_b200 import _capi, ops  # noqa: E402


def reflect101(p, n):
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * n - 2 - p
    return p


def kernel_q8(k):
    q = (C.c_int32 * k)()
    assert _capi.lib().docscan_gaussian_kernel_q8(k, q) == 0
    q = np.array(list(q))
    nz = np.nonzero(q)[0]
    return q[nz[0]:nz[-1] + 1]


def expected_tile0(img, k):
    q = kernel_q8(k)
    keff, R = len(q), len(q) // 2
    K1 = (128 + 2 * R + 31) // 32 * 32
    RL = (R + 15) // 16 * 16
    wide = R > 48                          # the 256-column tile (tcblur.cu: WIDE)
    NIN = 256 if wide else 128
    NOUT = min(64, (256 - 2 - RL - R) // 16 * 16) if wide else min(80, (128 - 2 - RL - R) // 16 * 16)
    h, w = img.shape
    S = np.zeros((K1, NIN), np.int64)
    for j in range(K1):
        y = -R + j
        if 0 <= y < h:
            for c in range(NIN):
                x = -RL + c
                if 0 <= x < w:
                    S[j, c] = img[y, x]
    T = np.zeros((128, K1), np.int64)
    for i in range(min(128, h)):
        for t in range(keff):
            T[i, reflect101(i + t - R, h) + R] += q[t]
    Th = np.zeros((NOUT, NIN), np.int64)
    for n in range(min(NOUT, w)):
        for t in range(keff):
            Th[n, reflect101(n + t - R, w) + RL] += q[t]
    D1 = T @ S
    lo, hi = D1 & 255, D1 >> 8
    Th[:, NIN - 2:] = 128                  # the rounding constant rides in the last two slots
    lo[:, NIN - 2:] = 128
    hi[:, NIN - 2:] = 0
    return D1, lo @ Th.T, hi @ Th.T, NOUT


def show(name, got, want):
    bad = np.count_nonzero(got != want)
    print(f"  {name}: {bad} of {want.size} differ")
    if bad:
        ys, xs = np.nonzero(got != want)
        print("    first mismatches (row, col, got, want):", [(int(y), int(x), int(got[y, x]), int(want[y, x])) for y, x in list(zip(ys, xs))[:6]])
        print("    got[0:4,0:8]  =", got[0:4, 0:8].tolist())
        print("    want[0:4,0:8] =", want[0:4, 0:8].tolist())
        # hypotheses
        if got.shape == want.shape and got.shape[0] == got.shape[1] and np.array_equal(got.T, want):
            print("    -> transposed")
    return bad


def main():
    ks = [int(a) for a in sys.argv[1:]] or [23, 51]
    rng = np.random.default_rng(1)
    total_bad = 0
    for k in ks:
        for (h, w) in [(300, 260), (128, 128), (97, 131), (1600, 1131)] + ([(2200, 3000)] if k > 100 else []):
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
            with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
                path = f.name
            os.environ["DOCSCAN_TC_DEBUG"] = path
            got = ops.gaussian_blur(img, k)
            del os.environ["DOCSCAN_TC_DEBUG"]
            want = O.gaussian_blur_u8(img, k)
            print(f"k={k} {h}x{w}:")
            total_bad += show("final", got, want)
            if os.path.getsize(path) == 0:
                print("  (no debug dump: the tensor-core path did not run)")
                continue
            dbg = np.fromfile(path, np.uint32).astype(np.int64)
            os.remove(path)
            D1, D2lo, D2hi, NOUT = expected_tile0(img, k)
            if D1.shape[1] == 128:
                show("D1  ", dbg[:16384].reshape(128, 128), D1)
            show("D2lo", dbg[16384:16384 + 12288].reshape(128, 96)[:, :NOUT], D2lo)
            show("D2hi", dbg[16384 + 12288:16384 + 2 * 12288].reshape(128, 96)[:, :NOUT], D2hi)
    # the fused epilogues through the stage functions
    from smart_image_processing_b200 import DocScanner as DS
    g = rng.integers(0, 256, (700, 900), dtype=np.uint8)
    for method in ("subtract", "divide"):
        bad = np.count_nonzero(DS.illumination_correction(g, method, 0.03) != O.illumination_correction(g, method, 0.03))
        print(f"illumination_correction {method}: {bad} differ"); total_bad += bad
    bad = np.count_nonzero(DS._compute_ink_mask(g, mask_blur_ksize=51) != O._compute_ink_mask(g, mask_blur_ksize=51))
    print(f"ink mask: {bad} differ"); total_bad += bad
    print("TOTAL BAD", total_bad)


if __name__ == "__main__":
    main()
