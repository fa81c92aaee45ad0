"""CPU: pins the oracle against the real OpenCV the reference calls (cv2 is in this image; the
reference's requirements.txt:1 leaves it unpinned — this image has 4.13.0).  Randomised; sized to run
in well under a minute.  Skipped when cv2 cannot be imported."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import oracle as O  # noqa: E402


def _page(rng, h, w):
    im = np.full((h, w), 200, np.float32)
    for i in range(max(1, h // 12)):
        y, x = 5 + 12 * i, 5
        while x < w - 20:
            ww = int(rng.integers(5, 40))
            im[y:y + 6, x:x + ww] = rng.integers(20, 90)
            x += ww + int(rng.integers(4, 12))
    im = im * (0.6 + 0.4 * np.linspace(0, 1, w)[None, :]) + rng.normal(0, 3, (h, w))
    return np.clip(im, 0, 255).astype(np.uint8)


def test_gray_exhaustive_slice():
    b, g, r = np.meshgrid(np.arange(0, 256, 3), np.arange(256), np.arange(0, 256, 5), indexing="ij")
    img = np.stack([b, g, r], -1).astype(np.uint8).reshape(-1, 64, 3)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), O.bgr2gray(img))
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), O.bgr2gray(img, True))


@pytest.mark.parametrize("k", [3, 5, 7, 9, 11, 23, 43, 51, 57, 101, 141, 217, 255])
def test_blur_u8(k):
    rng = np.random.default_rng(k)
    for shape in ((97, 131), (33, 260), (1, 50), (50, 1), (7, 5)):
        g = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(cv2.GaussianBlur(g, (k, k), 0), O.gaussian_blur_u8(g, k)), shape


def test_normalize_all_ranges():
    for mn in range(0, 256):
        for mx in range(mn, 256, 3):
            v = np.arange(mn, mx + 1, dtype=np.uint8)[None, :]
            assert np.array_equal(cv2.normalize(v, None, 0, 255, cv2.NORM_MINMAX), O.normalize_minmax(v))


def test_otsu_random():
    rng = np.random.default_rng(5)
    for _ in range(40):
        im = np.clip(rng.normal(rng.uniform(30, 220), rng.uniform(2, 70), (80, 90)), 0, 255).astype(np.uint8)
        if rng.random() < 0.5:
            im[rng.random(im.shape) < 0.2] = rng.integers(0, 60)
        t, _ = cv2.threshold(im, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert t == O.otsu_threshold(im)


def test_morphology_even_odd_iters():
    rng = np.random.default_rng(6)
    g = rng.integers(0, 256, (61, 47), dtype=np.uint8)
    for kw, kh in ((2, 2), (3, 3), (4, 4), (9, 19), (5, 2), (31, 31), (99, 3), (2, 99)):
        se = cv2.getStructuringElement(cv2.MORPH_RECT, (kw, kh))
        for it in (1, 2, 3):
            assert np.array_equal(cv2.erode(g, se, iterations=it), O.erode(g, kw, kh, it))
            assert np.array_equal(cv2.dilate(g, se, iterations=it), O.dilate(g, kw, kh, it))
            assert np.array_equal(cv2.morphologyEx(g, cv2.MORPH_CLOSE, se, iterations=it), O.morph_close(g, kw, kh, it))


def test_adaptive_threshold_shapes_and_tails():
    rng = np.random.default_rng(7)
    for (h, w) in ((120, 163), (90, 200), (77, 81), (60, 64), (33, 7), (64, 260), (40, 71), (1, 99), (99, 1)):
        g = _page(rng, h, w)
        for k in (3, 9, 11, 15, 31, 35):
            for c in (3, 10, -2):
                for meth, name in ((cv2.ADAPTIVE_THRESH_GAUSSIAN_C, "gaussian"), (cv2.ADAPTIVE_THRESH_MEAN_C, "mean")):
                    ref = cv2.adaptiveThreshold(g, 255, meth, cv2.THRESH_BINARY, k, c)
                    assert np.array_equal(ref, O.adaptive_threshold(g, name, k, c)), (h, w, k, c, name)


def test_gaussian_mean_is_float_exact_including_cv2_tail_columns():
    """The fp32 local mean of GAUSSIAN_C equals cv2's float GaussianBlur bit for bit, including the
    last w % 8 columns where cv2's AVX2 build switches between fma and mul+add (oracle header)."""
    rng = np.random.default_rng(8)
    for _ in range(60):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 300))
        k = int(rng.choice(np.arange(11, 64, 2)))
        g = rng.integers(0, 256, (h, w), dtype=np.uint8)
        m = cv2.GaussianBlur(g.astype(np.float32), (k, k), 0, borderType=cv2.BORDER_REPLICATE | cv2.BORDER_ISOLATED)
        _, _, mf = O.adaptive_threshold(g, "gaussian", k, 10, return_mean_f32=True)
        assert np.array_equal(m, mf), (h, w, k)


def test_matrices_bitwise():
    rng = np.random.default_rng(9)
    for _ in range(500):
        W, H = rng.integers(200, 4000, 2)
        quad = (np.array([[0.1 * W, 0.07 * H], [0.9 * W, 0.09 * H], [0.93 * W, 0.93 * H], [0.07 * W, 0.91 * H]])
                + rng.uniform(-60, 60, (4, 2))).astype(np.float32)
        tw, th = rng.integers(50, 3000, 2)
        dst = np.array([[0, 0], [tw - 1, 0], [tw - 1, th - 1], [0, th - 1]], np.float32)
        assert np.array_equal(cv2.getPerspectiveTransform(quad, dst), O.get_perspective_transform(quad, dst))
        ang = float(rng.integers(-20, 21)) * 0.5
        assert np.array_equal(cv2.getRotationMatrix2D((W / 2.0, H / 2.0), ang, 1.0),
                              O.rotation_matrix((W / 2.0, H / 2.0), ang))


def test_warps_random():
    rng = np.random.default_rng(10)
    for t in range(25):
        H, W = int(rng.integers(20, 300)), int(rng.integers(20, 400))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        quad = (np.array([[0.1 * W, 0.07 * H], [0.9 * W, 0.09 * H], [0.93 * W, 0.93 * H], [0.07 * W, 0.91 * H]])
                + rng.uniform(-0.15 * min(W, H), 0.15 * min(W, H), (4, 2))).astype(np.float32)
        tw, th = int(rng.integers(2, 333)), int(rng.integers(3, 300))
        dst = np.array([[0, 0], [tw - 1, 0], [tw - 1, th - 1], [0, th - 1]], np.float32)
        m = cv2.getPerspectiveTransform(quad, dst)
        assert np.array_equal(cv2.warpPerspective(img, m, (tw, th), flags=cv2.INTER_LINEAR),
                              O.warp_perspective(img, m, (tw, th)))
        g = img[:, :, 0].copy()
        ang = float(rng.integers(-20, 21)) * 0.5 if t % 2 else float(rng.uniform(-180, 180))
        m2 = cv2.getRotationMatrix2D((W / 2.0, H / 2.0), ang, 1.0)
        assert np.array_equal(cv2.warpAffine(g, m2, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE),
                              O.warp_affine(g, m2, (W, H)))


def test_mean_c_large_blocks_follow_cv2s_fp32_scale():
    """cv::boxFilter scales the box sum in fp32; from k = 165 on that differs from the exactly rounded quotient for some
    sums.  A checkerboard of n / n+1 puts every window mean within 1 / (2 k^2) of a tie."""
    yy, xx = np.mgrid[0:120, 0:173]
    for k in (35, 101, 163, 165, 201, 255):
        for n in (7, 100, 254):
            g = (n + ((yy + xx) & 1)).astype(np.uint8)
            for c in (0, 1):
                ref = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, k, c)
                assert np.array_equal(O.adaptive_threshold(g, "mean", k, c), ref), (k, n, c)
    rng = np.random.default_rng(12)
    g = rng.integers(0, 256, (150, 211), dtype=np.uint8)
    for k in (51, 151, 255):
        assert np.array_equal(O.adaptive_threshold(g, "mean", k, 3),
                              cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, k, 3))
