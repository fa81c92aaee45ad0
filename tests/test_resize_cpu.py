"""CPU: the oracle's resize_long_side (the whole-photo fallback, DocScanner.py:27-36) against fixtures made by running
the reference's own function (tests/golden/make_golden_resize.py) and against the cv2 of this image."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_npz
from oracle import oracle as O


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


META = json.load(open(os.path.join(GOLDEN, "resize_golden.json")))


@pytest.mark.parametrize("case", sorted(META["cases"]))
def test_resize_long_side_golden_crops(case):
    z = load_npz("resize.npz")
    tag, sl = case.rsplit("_", 1)
    out = O.resize_long_side(z[f"in_{tag}"], int(sl))
    ref = z[f"out_{case}"]
    assert out.shape == ref.shape and np.array_equal(out, ref), case


@pytest.mark.parametrize("name", ["sample3_1600", "sample3_1200"])
def test_resize_long_side_golden_full_photo(name):
    img = load_npz("resize.npz")["sample3_bgr"]
    m = META["full"][name]
    assert sha(img) == m["input_sha256"]
    out = O.resize_long_side(img, int(name.split("_")[1]))
    assert list(out.shape) == m["shape"] and sha(out) == m["sha256"]


def test_resize_returns_argument_for_nonpositive_scale():
    img = np.zeros((5, 7, 3), np.uint8)
    assert O.resize_long_side(img, 0) is img


def test_resize_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(17)
    for (H, W, nh, nw) in [(300, 400, 187, 250), (301, 403, 150, 201), (90, 120, 45, 60), (90, 120, 30, 40), (90, 120, 30, 60),
                           (97, 131, 48, 65), (35, 60, 7, 12), (50, 50, 50, 50), (64, 48, 1, 1)]:
        for cn in (1, 3):
            src = rng.integers(0, 256, (H, W, cn) if cn == 3 else (H, W), dtype=np.uint8)
            assert np.array_equal(O.resize_area(src, (nw, nh)), cv2.resize(src, (nw, nh), interpolation=cv2.INTER_AREA)), (H, W, nh, nw, cn)
    was = cv2.ipp.useIPP()
    try:
        for (H, W, nh, nw) in [(300, 400, 375, 500), (60, 80, 75, 100), (97, 131, 200, 333), (5, 3, 17, 9), (1, 50, 2, 80),
                               (100, 77, 100, 77), (40, 30, 41, 30)]:
            for cn in (1, 3):
                src = rng.integers(0, 256, (H, W, cn) if cn == 3 else (H, W), dtype=np.uint8)
                cv2.ipp.setUseIPP(False)
                own = cv2.resize(src, (nw, nh), interpolation=cv2.INTER_CUBIC)
                cv2.ipp.setUseIPP(True)
                ipp = cv2.resize(src, (nw, nh), interpolation=cv2.INTER_CUBIC)
                out = O.resize_cubic(src, (nw, nh))
                assert np.array_equal(out, own), (H, W, nh, nw, cn)                       # OpenCV's own code: bit-exact
                assert np.abs(out.astype(int) - ipp.astype(int)).max() <= 1               # the IPP flavour: +-1 LSB
    finally:
        cv2.ipp.setUseIPP(was)
