"""CPU: the oracle's restatement of deskew()'s skew estimate (DocScanner.py:218-231) — cv2.Canny, cv2.HoughLines, the
numpy float32 median — against the cv2 of this image, against the angles the reference itself produced for
public/sample.jpg (tests/golden/sample_golden.json), and libdocscan's host arithmetic for the median against both."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_npz
from oracle import oracle as O
from smart_image_processing_b200 import ops


def _binary_page(rng, h, w):
    im = np.full((h, w), 255, np.uint8)
    for i in range(max(1, h // 14)):
        y, x = 4 + 14 * i, 4
        while x < w - 12:
            ww = int(rng.integers(4, 40))
            im[y:y + 7, x:x + ww] = 0
            x += ww + int(rng.integers(3, 14))
    return im


def _numpy_reference_angle(thetas, max_rotate=10.0):
    """DocScanner.py:221-231, verbatim arithmetic (numpy >= 2: the float32 thetas keep everything in float32)."""
    angles = []
    for theta in thetas:
        ang = (theta * 180.0 / np.pi)
        ang = (ang + 90.0) % 180.0 - 90.0
        angles.append(ang)
    if not angles:
        return 0.0
    a = float(np.median(angles))
    return 0.0 if abs(a) > max_rotate else a


def test_median_angle_matches_numpy_semantics():
    rng = np.random.default_rng(3)
    theta_of = [np.float32(0) + np.float32(n) * np.float32(np.pi / 180) for n in range(180)]
    for t in range(400):
        per = np.zeros(180, np.int32)
        for n in rng.integers(0, 180, int(rng.integers(1, 7))):
            per[n] += int(rng.integers(1, 5))
        if t % 5 == 0:                                      # near-upright text: angles around 90 +- few degrees
            per[:] = 0
            for n in rng.integers(84, 97, int(rng.integers(1, 9))):
                per[n] += int(rng.integers(1, 4))
        thetas = [theta_of[n] for n in range(180) for _ in range(per[n])]
        ref = _numpy_reference_angle(thetas)
        assert O.median_angle(per) == ref
        assert ops.median_angle(per) == ref                 # libdocscan host arithmetic (no GPU needed)
    assert O.median_angle(np.zeros(180, np.int32)) == 0.0 and ops.median_angle(np.zeros(180, np.int32)) == 0.0


def test_sample_jpg_angles_from_the_reference():
    """The blended page of sample.jpg (committed golden) must give the angle the reference's deskew() used."""
    meta = json.load(open(os.path.join(GOLDEN, "sample_golden.json")))
    for preset in ("cli", "gui"):
        p = meta["presets"][preset]
        weighted = load_npz(f"sample_{preset}_bin.npz")["weighted"]
        a = O.estimate_skew_angle(weighted, p["params"]["canny_low"], p["params"]["canny_high"], p["params"]["max_rotate"])
        assert a == float.fromhex(p["angle_hex"]), (preset, a, p["angle"])


def test_canny_and_hough_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    for (h, w, ang) in [(400, 300, 0.0), (500, 380, 2.0), (300, 500, -3.5), (64, 64, 0.0), (5, 7, 0.0), (1, 40, 0.0), (40, 1, 0.0)]:
        g = _binary_page(rng, h, w) if min(h, w) > 10 else rng.integers(0, 256, (h, w), dtype=np.uint8)
        if ang:
            m = cv2.getRotationMatrix2D((w / 2, h / 2), ang, 1.0)
            g = cv2.warpAffine(g, m, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        noisy = np.clip(g.astype(np.int32) + rng.integers(-40, 41, g.shape), 0, 255).astype(np.uint8)
        for img in (g, noisy):
            for lo, hi in ((50, 150), (30, 100), (150, 50), (12.7, 99.2)):
                e = cv2.Canny(img, lo, hi)
                assert np.array_equal(O.canny(img, lo, hi), e), (h, w, ang, lo, hi)
            e = cv2.Canny(img, 50, 150)
            for thr in (150, 60, 25):
                ref = cv2.HoughLines(e, 1, np.pi / 180, thr)
                mine, per = O.hough_lines(e, thr)
                assert (ref is None) == (mine is None)
                if ref is not None:
                    assert np.array_equal(ref, mine), (h, w, ang, thr)
                    assert per.sum() == len(ref)
                    assert O.median_angle(per) == _numpy_reference_angle(ref[:, 0, 1])
