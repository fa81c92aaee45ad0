"""numpy-in / numpy-out wrappers of the single-op C-ABI entry points: one per cv2 call the reference's
DocScanner.py makes on its per-pixel path.  Every function runs a CUDA kernel of libdocscan.so."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import image_of


def _ctx(ctx):
    return ctx if ctx is not None else _capi.default_context()


def _gray(a, what="image"):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise TypeError(f"{what}: expected a 2-D uint8 array, got {a.dtype} {a.shape}")
    if a.size == 0:
        raise ValueError(f"{what}: empty image")
    return a


def _dptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def bgr2gray(img, swap_rb=False, ctx=None):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) (DocScanner.py:316); swap_rb=True gives COLOR_RGB2GRAY."""
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise TypeError("bgr2gray: expected HxWx3 uint8")
    out = np.empty(img.shape[:2], np.uint8)
    s, d = image_of(img), image_of(out)
    _ctx(ctx).call("docscan_bgr2gray", C.byref(s), C.byref(d), int(swap_rb))
    return out


def gaussian_blur(gray, ksize, ctx=None):
    """cv2.GaussianBlur(gray, (ksize, ksize), 0) (DocScanner.py:153,184)."""
    gray = _gray(gray)
    out = np.empty_like(gray)
    s, d = image_of(gray), image_of(out)
    _ctx(ctx).call("docscan_gaussian_blur", C.byref(s), int(ksize), C.byref(d))
    return out


def _binary(op, a, b, ctx):
    a, b = _gray(a), _gray(b)
    if a.shape != b.shape:
        raise ValueError("operands differ in shape")
    out = np.empty_like(a)
    ia, ib, d = image_of(a), image_of(b), image_of(out)
    _ctx(ctx).call("docscan_binary_op", op, C.byref(ia), C.byref(ib), C.byref(d))
    return out


def subtract(a, b, ctx=None):
    """cv2.subtract(a, b) (DocScanner.py:158,185)."""
    return _binary(_capi.OP_SUB, a, b, ctx)


def divide255(a, b, ctx=None):
    """cv2.divide(a, b, scale=255) (DocScanner.py:155)."""
    return _binary(_capi.OP_DIV255, a, b, ctx)


def maximum(a, b, ctx=None):
    """cv2.max(a, b) (DocScanner.py:207)."""
    return _binary(_capi.OP_MAX, a, b, ctx)


def mask_select(base, mask, ctx=None):
    """out = base.copy(); out[mask == 0] = 255 (DocScanner.py:338-339)."""
    return _binary(_capi.OP_MASK_SELECT, base, mask, ctx)


def minmax(gray, ctx=None):
    gray = _gray(gray)
    mn, mx = C.c_int32(), C.c_int32()
    s = image_of(gray)
    _ctx(ctx).call("docscan_minmax", C.byref(s), C.byref(mn), C.byref(mx))
    return mn.value, mx.value


def hist256(gray, ctx=None):
    gray = _gray(gray)
    hist = np.zeros(256, np.int32)
    s = image_of(gray)
    _ctx(ctx).call("docscan_hist256", C.byref(s), _dptr(hist, C.c_int32))
    return hist


def normalize_minmax(gray, ctx=None):
    """cv2.normalize(gray, None, 0, 255, NORM_MINMAX) (DocScanner.py:156,159,172,186,201)."""
    gray = _gray(gray)
    out = np.empty_like(gray)
    s, d = image_of(gray), image_of(out)
    _ctx(ctx).call("docscan_normalize_minmax", C.byref(s), C.byref(d))
    return out


def otsu_threshold(gray, ctx=None, return_image=False):
    """cv2.threshold(gray, 0, 255, THRESH_BINARY + THRESH_OTSU) -> (t, binary) (DocScanner.py:187,202)."""
    gray = _gray(gray)
    t = C.c_double()
    s = image_of(gray)
    if return_image:
        out = np.empty_like(gray)
        d = image_of(out)
        _ctx(ctx).call("docscan_otsu_threshold", C.byref(s), C.byref(t), C.byref(d))
        return t.value, out
    _ctx(ctx).call("docscan_otsu_threshold", C.byref(s), C.byref(t), None)
    return t.value


def threshold_binary(gray, t, ctx=None):
    """cv2.threshold(gray, t, 255, THRESH_BINARY)[1] (DocScanner.py:189,204)."""
    gray = _gray(gray)
    out = np.empty_like(gray)
    s, d = image_of(gray), image_of(out)
    _ctx(ctx).call("docscan_threshold_binary", C.byref(s), int(math.floor(t)), C.byref(d))
    return out


def morph_rect(gray, op, kw, kh=None, iterations=1, ctx=None):
    gray = _gray(gray)
    out = np.empty_like(gray)
    s, d = image_of(gray), image_of(out)
    _ctx(ctx).call("docscan_morph_rect", int(op), C.byref(s), int(kw), int(kh if kh is not None else kw), int(iterations), C.byref(d))
    return out


def erode(gray, kw, kh=None, iterations=1, ctx=None):
    """cv2.erode(gray, getStructuringElement(MORPH_RECT, (kw, kh)), iterations=...)."""
    return morph_rect(gray, _capi.MORPH_ERODE, kw, kh, iterations, ctx)


def dilate(gray, kw, kh=None, iterations=1, ctx=None):
    """cv2.dilate(...) (DocScanner.py:211-212)."""
    return morph_rect(gray, _capi.MORPH_DILATE, kw, kh, iterations, ctx)


def morph_close(gray, kw, kh=None, iterations=1, ctx=None):
    """cv2.morphologyEx(gray, MORPH_CLOSE, rect) (DocScanner.py:254)."""
    return morph_rect(gray, _capi.MORPH_CLOSE, kw, kh, iterations, ctx)


def morph_open(gray, kw, kh=None, iterations=1, ctx=None):
    return morph_rect(gray, _capi.MORPH_OPEN, kw, kh, iterations, ctx)


def blackhat(gray, kw, kh=None, ctx=None):
    """cv2.morphologyEx(gray, MORPH_BLACKHAT, rect) (DocScanner.py:199-200)."""
    return morph_rect(gray, _capi.MORPH_BLACKHAT, kw, kh, 1, ctx)


def adaptive_threshold(gray, method, block_size, c, cv_tail_compat=True, ctx=None):
    """cv2.adaptiveThreshold(gray, 255, method, THRESH_BINARY, block_size, c) (DocScanner.py:167)."""
    gray = _gray(gray)
    out = np.empty_like(gray)
    s, d = image_of(gray), image_of(out)
    m = _capi.ADAPTIVE_GAUSSIAN if str(method).lower() in ("gaussian", "1") else _capi.ADAPTIVE_MEAN
    _ctx(ctx).call("docscan_adaptive_threshold", C.byref(s), m, int(block_size), int(math.ceil(c)), int(bool(cv_tail_compat)), C.byref(d))
    return out


def get_perspective_transform(quad, dst):
    """cv2.getPerspectiveTransform (DocScanner.py:142)."""
    q = np.ascontiguousarray(quad, np.float32).reshape(8)
    d = np.ascontiguousarray(dst, np.float32).reshape(8)
    m = np.zeros(9, np.float64)
    rc = _capi.lib().docscan_get_perspective_transform(_dptr(q, C.c_float), _dptr(d, C.c_float), _dptr(m, C.c_double))
    if rc:
        raise _capi.DocscanError("docscan_get_perspective_transform failed")
    return m.reshape(3, 3)


def get_rotation_matrix(center, angle_deg):
    """cv2.getRotationMatrix2D(center, angle, 1.0) (DocScanner.py:234)."""
    m = np.zeros(6, np.float64)
    rc = _capi.lib().docscan_get_rotation_matrix(float(center[0]), float(center[1]), float(angle_deg), _dptr(m, C.c_double))
    if rc:
        raise _capi.DocscanError("docscan_get_rotation_matrix failed")
    return m.reshape(2, 3)


def warp_perspective(img, m, dsize, return_gray=False, ctx=None):
    """cv2.warpPerspective(img, m, dsize, flags=INTER_LINEAR) (DocScanner.py:143)."""
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8 or img.ndim not in (2, 3):
        raise TypeError("warp_perspective: expected uint8 HxW or HxWx3")
    dw, dh = int(dsize[0]), int(dsize[1])
    out = np.empty((dh, dw) + img.shape[2:], np.uint8)
    m = np.ascontiguousarray(m, np.float64).reshape(9)
    s, d = image_of(img), image_of(out)
    if return_gray:
        g = np.empty((dh, dw), np.uint8)
        gi = image_of(g)
        _ctx(ctx).call("docscan_warp_perspective", C.byref(s), _dptr(m, C.c_double), C.byref(d), C.byref(gi))
        return out, g
    _ctx(ctx).call("docscan_warp_perspective", C.byref(s), _dptr(m, C.c_double), C.byref(d), None)
    return out


def warp_affine(gray, m, dsize, ctx=None):
    """cv2.warpAffine(gray, m, dsize, flags=INTER_LINEAR, borderMode=BORDER_REPLICATE) (DocScanner.py:235)."""
    gray = _gray(gray)
    dw, dh = int(dsize[0]), int(dsize[1])
    out = np.empty((dh, dw), np.uint8)
    m = np.ascontiguousarray(m, np.float64).reshape(6)
    s, d = image_of(gray), image_of(out)
    _ctx(ctx).call("docscan_warp_affine", C.byref(s), _dptr(m, C.c_double), C.byref(d))
    return out


def resize(img, dsize, interpolation, cv_tail_compat=True, ctx=None):
    """cv2.resize(img, dsize, interpolation=...) for INTER_AREA (shrink) and INTER_CUBIC (DocScanner.py:35-36)."""
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8 or img.ndim not in (2, 3):
        raise TypeError("resize: expected uint8 HxW or HxWx3")
    dw, dh = int(dsize[0]), int(dsize[1])
    if dw <= 0 or dh <= 0:
        raise ValueError("resize: empty destination")
    out = np.empty((dh, dw) + img.shape[2:], np.uint8)
    s, d = image_of(img), image_of(out)
    _ctx(ctx).call("docscan_resize", C.byref(s), C.byref(d), int(interpolation), int(bool(cv_tail_compat)))
    return out


def canny(gray, low, high, ctx=None):
    """cv2.Canny(gray, low, high) (DocScanner.py:218)."""
    gray = _gray(gray)
    out = np.empty_like(gray)
    s, d = image_of(gray), image_of(out)
    _ctx(ctx).call("docscan_canny", C.byref(s), float(low), float(high), C.byref(d))
    return out


def hough_lines(edges, threshold=150, max_lines=None, return_per_angle=False, ctx=None):
    """cv2.HoughLines(edges, 1, np.pi / 180, threshold) (DocScanner.py:219): (N, 1, 2) float32 (rho, theta) in OpenCV's
    order, or None when no line clears the threshold."""
    edges = _gray(edges)
    e = image_of(edges)
    n = C.c_int32()
    per = np.zeros(180, np.int32)
    cap = 4096 if max_lines is None else int(max_lines)
    while True:
        buf = np.zeros((max(cap, 1), 2), np.float32)
        _ctx(ctx).call("docscan_hough_lines", C.byref(e), int(threshold), _dptr(buf, C.c_float), cap, C.byref(n), _dptr(per, C.c_int32))
        if max_lines is not None or n.value <= cap:
            break
        cap = n.value
    k = min(n.value, cap)
    lines = buf[:k].reshape(-1, 1, 2).copy() if k else None
    return (lines, per) if return_per_angle else lines


def median_angle(per_angle, max_rotate=10.0):
    """DocScanner.py:221-231 from the number of Hough lines per angle index (host arithmetic of libdocscan)."""
    per = np.ascontiguousarray(per_angle, np.int32)
    a = C.c_double()
    rc = _capi.lib().docscan_median_angle(_dptr(per, C.c_int32), float(max_rotate), C.byref(a))
    if rc:
        raise _capi.DocscanError("docscan_median_angle failed")
    return float(a.value)


def skew_angle(gray, canny_low=50, canny_high=150, max_rotate=10.0, ctx=None):
    """The angle deskew() rotates by: Canny -> HoughLines(1, pi/180, 150) -> median of the folded line angles
    (DocScanner.py:218-231), all on the device."""
    gray = _gray(gray)
    s = image_of(gray)
    a = C.c_double()
    _ctx(ctx).call("docscan_skew_angle", C.byref(s), float(canny_low), float(canny_high), float(max_rotate), C.byref(a))
    return float(a.value)
