"""CPU: host-side control helpers that gate the pixel path (no GPU needed)."""
import numpy as np
import pytest

from smart_image_processing_b200 import control


def test_quad_area_is_cv2_contour_area():
    """process_document's min_quad_area_ratio gate uses cv2.contourArea on the float32 quad (DocScanner.py:291-293);
    control.quad_area restates it (double accumulation of float32 products, same order) and must agree to the bit."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for t in range(2000):
        scale = [1.0, 37.5, 4000.0, 1e-3][t % 4]
        q = (rng.uniform(-1, 1, (4, 2)) * scale).astype(np.float32)
        if t % 9 == 0:
            q[1] = q[0]                                   # degenerate quads are not rejected by the reference
        assert control.quad_area(q) == float(cv2.contourArea(q.reshape(-1, 1, 2))), q.tolist()
    q = np.array([[0, 12], [962, 12], [962, 1279], [-3.879069035927184e-14, 1279]], np.float32)
    assert control.quad_area(q) == float(cv2.contourArea(q.reshape(-1, 1, 2)))


def test_quad_overlay_matches_reference_drawing():
    cv2 = pytest.importorskip("cv2")
    img = np.full((120, 160, 3), 90, np.uint8)
    quad = np.array([[10.6, 8.2], [140.9, 12.1], [150.2, 100.7], [5.5, 110.4]], np.float32)
    want = img.copy()
    cv2.polylines(want, [quad.astype(np.int32).reshape((-1, 1, 2))], True, (0, 255, 0), 2)
    assert np.array_equal(control.quad_overlay(img, quad), want)
    want = img.copy()
    full = np.array([[0, 0], [159, 0], [159, 119], [0, 119]], dtype=np.int32).reshape((-1, 1, 2))
    cv2.polylines(want, [full], True, (0, 165, 255), 2)
    assert np.array_equal(control.quad_overlay(img, None), want)
    assert (img == 90).all()                               # input untouched
