"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/libdocscan_oracle.so (the plain-C CPU restatement of the OpenCV
arithmetic the reference's DocScanner.py relies on) plus the reference's stage functions
restated on top of it, each citing the reference line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under smart_image_processing_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdocscan_oracle.so")
_lib = None

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "docscan_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdocscan_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_otsu_from_hist.restype = C.c_double
    return _lib


def _img(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8:
        raise TypeError("oracle images are uint8")
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


# --------------------------------------------------------------------------- primitive ops

def bgr2gray(img: np.ndarray, swap_rb: bool = False) -> np.ndarray:
    img = _img(img)
    h, w = img.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().orc_bgr2gray(_p(img), h, w, w * 3, _p(out), w, int(swap_rb))
    return out


def gaussian_kernel_q8(k: int) -> np.ndarray:
    q = np.zeros(k, np.int32)
    lib().orc_gaussian_kernel_q8(k, q.ctypes.data_as(_i32p))
    return q


def gaussian_kernel_f32(k: int) -> np.ndarray:
    g = np.zeros(k, np.float32)
    lib().orc_gaussian_kernel_f32(k, g.ctypes.data_as(_f32p))
    return g


def gaussian_blur_u8(gray: np.ndarray, k: int) -> np.ndarray:
    gray = _img(gray)
    h, w = gray.shape
    out = np.empty_like(gray)
    lib().orc_gaussian_blur_u8(_p(gray), h, w, w, k, _p(out), w)
    return out


def _binop(a, b, mode):
    a = _img(a); b = _img(b)
    h, w = a.shape
    out = np.empty_like(a)
    lib().orc_binary_op(_p(a), w, _p(b), w, h, w, mode, _p(out), w)
    return out


def subtract(a, b): return _binop(a, b, 0)
def divide255(a, b): return _binop(a, b, 1)
def maximum(a, b): return _binop(a, b, 2)
def mask_select(base, mask): return _binop(base, mask, 3)


def minmax(gray):
    gray = _img(gray)
    mn, mx = C.c_int(), C.c_int()
    lib().orc_minmax(_p(gray), gray.shape[0], gray.shape[1], gray.shape[1], C.byref(mn), C.byref(mx))
    return mn.value, mx.value


def normalize_lut(smin: int, smax: int) -> np.ndarray:
    lut = np.zeros(256, np.uint8)
    lib().orc_normalize_lut(smin, smax, _p(lut))
    return lut


def normalize_minmax(gray):
    gray = _img(gray)
    out = np.empty_like(gray)
    lib().orc_normalize_minmax(_p(gray), gray.shape[0], gray.shape[1], gray.shape[1], _p(out), gray.shape[1])
    return out


def hist256(gray):
    gray = _img(gray)
    hist = np.zeros(256, np.int32)
    lib().orc_hist256(_p(gray), gray.shape[0], gray.shape[1], gray.shape[1], hist.ctypes.data_as(_i32p))
    return hist


def otsu_from_hist(hist, total):
    hist = np.ascontiguousarray(hist, np.int32)
    return float(lib().orc_otsu_from_hist(hist.ctypes.data_as(_i32p), C.c_int64(int(total))))


def otsu_threshold(gray) -> float:
    return otsu_from_hist(hist256(gray), gray.size)


def threshold_binary(gray, t: int):
    gray = _img(gray)
    out = np.empty_like(gray)
    lib().orc_threshold_binary(_p(gray), gray.shape[0], gray.shape[1], gray.shape[1], int(t), _p(out), gray.shape[1])
    return out


def morph_rect(gray, kw: int, kh: int, op: int, iterations: int = 1):
    """op 0 = erode, 1 = dilate; MORPH_RECT kw (wide) x kh (high)."""
    gray = _img(gray)
    out = np.empty_like(gray)
    lib().orc_morph_rect(_p(gray), gray.shape[0], gray.shape[1], gray.shape[1], kw, kh, op, iterations,
                         _p(out), gray.shape[1])
    return out


def erode(gray, kw, kh=None, iterations=1): return morph_rect(gray, kw, kh or kw, 0, iterations)
def dilate(gray, kw, kh=None, iterations=1): return morph_rect(gray, kw, kh or kw, 1, iterations)


def morph_close(gray, kw, kh=None, iterations=1):
    kh = kh or kw
    return morph_rect(morph_rect(gray, kw, kh, 1, iterations), kw, kh, 0, iterations)


def blackhat(gray, kw, kh):
    return subtract(morph_close(gray, kw, kh), gray)


def adaptive_threshold(gray, method: str, k: int, c: int, unfused_tail=None, return_mean=False,
                       return_mean_f32=False):
    gray = _img(gray)
    h, w = gray.shape
    out = np.empty_like(gray)
    mean = np.empty_like(gray)
    mean_f = np.zeros((h, w), np.float32) if return_mean_f32 else None
    if unfused_tail is None:
        unfused_tail = w % 8   # what cv2 does on every AVX2-capable x86 host
    lib().orc_adaptive_threshold(_p(gray), h, w, w, 1 if method == "gaussian" else 0, k, int(math.ceil(c)),
                                 int(unfused_tail), _p(out), w, _p(mean),
                                 mean_f.ctypes.data_as(_f32p) if mean_f is not None else None)
    if return_mean_f32:
        return out, mean, mean_f
    return (out, mean) if return_mean else out


def get_perspective_transform(quad: np.ndarray, dst: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(quad, np.float32).reshape(8)
    d = np.ascontiguousarray(dst, np.float32).reshape(8)
    m = np.zeros(9, np.float64)
    lib().orc_get_perspective_transform(q.ctypes.data_as(_f32p), d.ctypes.data_as(_f32p), m.ctypes.data_as(_f64p))
    return m.reshape(3, 3)


def warp_perspective(img, m, dsize):
    img = _img(img)
    cn = 1 if img.ndim == 2 else img.shape[2]
    sh, sw = img.shape[:2]
    dw, dh = dsize
    out = np.empty((dh, dw) if img.ndim == 2 else (dh, dw, cn), np.uint8)
    m = np.ascontiguousarray(m, np.float64).reshape(9)
    lib().orc_warp_perspective_u8(_p(img), sh, sw, sw * cn, cn, m.ctypes.data_as(_f64p), _p(out), dh, dw, dw * cn)
    return out


def rotation_matrix(center, angle_deg: float) -> np.ndarray:
    m = np.zeros(6, np.float64)
    lib().orc_rotation_matrix(C.c_double(center[0]), C.c_double(center[1]), C.c_double(angle_deg),
                              m.ctypes.data_as(_f64p))
    return m.reshape(2, 3)


def warp_affine(gray, m, dsize):
    gray = _img(gray)
    sh, sw = gray.shape
    dw, dh = dsize
    out = np.empty((dh, dw), np.uint8)
    m = np.ascontiguousarray(m, np.float64).reshape(6)
    lib().orc_warp_affine_u8(_p(gray), sh, sw, sw, m.ctypes.data_as(_f64p), _p(out), dh, dw, dw)
    return out


# --------------------------------------------------------------------------- reference stages

def target_size(quad: np.ndarray, page: str = "A4", scale_long: int = 1600):
    """DocScanner.py:120-139 — size of the warped page."""
    tl, tr, br, bl = np.asarray(quad)
    w_top = np.linalg.norm(tr - tl)
    w_bottom = np.linalg.norm(br - bl)
    h_left = np.linalg.norm(bl - tl)
    h_right = np.linalg.norm(br - tr)
    width = max(int(w_top), int(w_bottom))
    height = max(int(h_left), int(h_right))
    portrait = height >= width
    if page.upper() in ("A4", "A3", "A5", "LETTER"):
        ratio = math.sqrt(2.0) if page.upper() != "LETTER" else (11.0 / 8.5)
    else:
        ratio = height / max(width, 1)
    if portrait:
        target_h = scale_long
        target_w = int(round(target_h / ratio))
    else:
        target_w = scale_long
        target_h = int(round(target_w * ratio))
    return target_w, target_h


def perspective_warp(img, quad, page="A4", scale_long=1600):
    """DocScanner.py:117-144."""
    tw, th = target_size(quad, page, scale_long)
    dst = np.array([[0, 0], [tw - 1, 0], [tw - 1, th - 1], [0, th - 1]], dtype=np.float32)
    m = get_perspective_transform(np.asarray(quad).astype(np.float32), dst)
    return warp_perspective(img, m, (tw, th))


def illum_ksize(h: int, w: int, blur_frac: float) -> int:
    base = max(15, int(round(min(h, w) * blur_frac)))   # DocScanner.py:150-152
    if base % 2 == 0:
        base += 1
    return base


def illumination_correction(gray, method="subtract", blur_frac=0.02):
    """DocScanner.py:147-160."""
    h, w = gray.shape[:2]
    bg = gaussian_blur_u8(gray, illum_ksize(h, w, blur_frac))
    tmp = divide255(gray, bg) if method.lower() == "divide" else subtract(gray, bg)
    return normalize_minmax(tmp)


def adaptive_binarize(gray, block_size=35, C=10, method="gaussian"):
    """DocScanner.py:163-168."""
    if block_size % 2 == 0:
        block_size += 1
    return adaptive_threshold(gray, "gaussian" if method.lower() == "gaussian" else "mean", block_size, C)


def contrast_stretch(gray):
    """DocScanner.py:171-172."""
    return normalize_minmax(gray)


def _compute_ink_mask(gray, mask_blur_ksize=61, blackhat_ksize=9, blackhat_vertical_ratio=2.0,
                      dilate_iters=1, threshold_offset=8):
    """DocScanner.py:175-214."""
    if mask_blur_ksize % 2 == 0:
        mask_blur_ksize += 1
    bg = gaussian_blur_u8(gray, mask_blur_ksize)
    ink_sub = normalize_minmax(subtract(bg, gray))
    t_sub = otsu_threshold(ink_sub)
    t_sub = max(0, int(round(t_sub - threshold_offset)))
    mask_sub = threshold_binary(ink_sub, t_sub)

    if blackhat_ksize < 3:
        blackhat_ksize = 3
    if blackhat_ksize % 2 == 0:
        blackhat_ksize += 1
    bh_h = max(3, int(round(blackhat_ksize * blackhat_vertical_ratio)))
    if bh_h % 2 == 0:
        bh_h += 1
    bh = normalize_minmax(blackhat(gray, blackhat_ksize, bh_h))
    t_bh = otsu_threshold(bh)
    t_bh = max(0, int(round(t_bh - threshold_offset)))
    mask_bh = threshold_binary(bh, t_bh)

    combined = maximum(mask_sub, mask_bh)
    if dilate_iters > 0:
        combined = dilate(combined, 2, 2, dilate_iters)
    return combined


def rotate(gray, angle_deg: float):
    """Rotation part of deskew — DocScanner.py:233-236 (angle supplied by the control path)."""
    h, w = gray.shape[:2]
    m = rotation_matrix((w / 2.0, h / 2.0), angle_deg)
    return warp_affine(gray, m, (w, h))


def morph_cleanup(bin_img, ksize=3, iterations=1):
    """DocScanner.py:247-259."""
    if ksize <= 1:
        return bin_img
    return morph_close(bin_img, ksize, ksize, iterations)


def hot_path(color, quad, angle_deg, *, page="A4", scale_long=1600, illum_method="subtract",
             illum_blur_frac=0.02, block_size=35, C=10, thresh_method="gaussian", mask_blur_ksize=51,
             blackhat_ksize=9, blackhat_vertical_ratio=2.0, ink_dilate_iters=1, mask_thresh_offset=8,
             morph_ksize=3, morph_iters=1):
    """Per-pixel part of process_document (DocScanner.py:310-346) with the control-path outputs
    (quad, deskew angle) supplied.  Returns every stage image keyed like the reference's dumps."""
    out = {}
    # DocScanner.py:310-313: no usable quad -> the whole photo through resize_long_side
    out["warped"] = (perspective_warp(color, quad, page=page, scale_long=scale_long) if quad is not None
                     else resize_long_side(color, scale_long))
    out["gray"] = bgr2gray(out["warped"])
    out["illum"] = illumination_correction(out["gray"], method=illum_method, blur_frac=illum_blur_frac)
    out["stretch"] = contrast_stretch(out["illum"])
    out["inkmask"] = _compute_ink_mask(out["stretch"], mask_blur_ksize=mask_blur_ksize,
                                       blackhat_ksize=blackhat_ksize,
                                       blackhat_vertical_ratio=blackhat_vertical_ratio,
                                       dilate_iters=ink_dilate_iters, threshold_offset=mask_thresh_offset)
    out["adapt"] = adaptive_binarize(out["stretch"], block_size=block_size, C=C, method=thresh_method)
    out["weighted"] = mask_select(out["adapt"], out["inkmask"])          # DocScanner.py:338-339
    out["deskew"] = rotate(out["weighted"], angle_deg)
    out["clean"] = morph_cleanup(out["deskew"], ksize=morph_ksize, iterations=morph_iters)
    return out


# --------------------------------------------------------------------------- morph_seq (pyc only)

def resize_area(img, dsize):
    """cv2.resize(img, dsize, interpolation=cv2.INTER_AREA) for a shrink (DocScanner.py:35-36)."""
    img = _img(img)
    dw, dh = dsize
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((dh, dw) if img.ndim == 2 else (dh, dw, cn), np.uint8)
    rc = lib().orc_resize_area_u8(_p(img), img.shape[0], img.shape[1], img.strides[0], cn, _p(out), dh, dw, out.strides[0])
    if rc != 0:
        raise ValueError("resize_area: destination must not be larger than the source")
    return out


def resize_cubic(img, dsize, simd_tail=True):
    """cv2.resize(img, dsize, interpolation=cv2.INTER_CUBIC) as OpenCV's own code (IPP off) computes it."""
    img = _img(img)
    dw, dh = dsize
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((dh, dw) if img.ndim == 2 else (dh, dw, cn), np.uint8)
    rc = lib().orc_resize_cubic_u8(_p(img), img.shape[0], img.shape[1], img.strides[0], cn, _p(out), dh, dw, out.strides[0],
                                   int(bool(simd_tail)))
    if rc != 0:
        raise ValueError("resize_cubic: empty image")
    return out


def resize_long_side(img, scale_long: int):
    """DocScanner.py:27-36."""
    h, w = img.shape[:2]
    if scale_long <= 0:
        return img
    long = max(h, w)
    sf = scale_long / float(long)
    new_w = int(round(w * sf))
    new_h = int(round(h * sf))
    return resize_area(img, (new_w, new_h)) if sf < 1.0 else resize_cubic(img, (new_w, new_h))


def canny(gray, low, high):
    """cv2.Canny(gray, low, high) (DocScanner.py:218)."""
    gray = _img(gray)
    out = np.empty_like(gray)
    lib().orc_canny(_p(gray), gray.shape[0], gray.shape[1], gray.strides[0], C.c_double(low), C.c_double(high), _p(out), out.strides[0])
    return out


def hough_lines(edges, threshold=150, max_lines=1 << 20):
    """cv2.HoughLines(edges, 1, np.pi / 180, threshold) (DocScanner.py:219): (lines (N, 1, 2) float32 or None, lines per angle)."""
    edges = _img(edges)
    per_angle = np.zeros(180, np.int32)
    buf = np.zeros((max_lines, 2), np.float32)
    n = lib().orc_hough_lines(_p(edges), edges.shape[0], edges.shape[1], edges.strides[0], int(threshold),
                              buf.ctypes.data_as(_f32p), max_lines, per_angle.ctypes.data_as(_i32p))
    return (buf[:min(n, max_lines)].reshape(-1, 1, 2).copy() if n else None), per_angle


def median_angle(per_angle, max_rotate=10.0) -> float:
    """DocScanner.py:221-231 from the number of Hough lines per angle index."""
    per_angle = np.ascontiguousarray(per_angle, np.int32)
    fn = lib().orc_median_angle
    fn.restype = C.c_double
    return float(fn(per_angle.ctypes.data_as(_i32p), C.c_double(max_rotate)))


def estimate_skew_angle(gray, canny_low=50, canny_high=150, max_rotate=10.0) -> float:
    """The control half of deskew (DocScanner.py:218-231)."""
    _, per_angle = hough_lines(canny(gray, canny_low, canny_high), 150, max_lines=1)   # only the per-angle counts matter
    return median_angle(per_angle, max_rotate)


def deskew(gray, canny_low=50, canny_high=150, max_rotate=10.0):
    """DocScanner.py:217-236."""
    return rotate(gray, estimate_skew_angle(gray, canny_low, canny_high, max_rotate))


def to_grayscale(rgb):
    """morph_seq.to_grayscale (pyc src l.46-47): cv2.cvtColor(RGB2GRAY)."""
    return bgr2gray(rgb, swap_rb=True)


def grayscale_erosion(gray, ksize=2, iterations=1):
    """morph_seq.grayscale_erosion (pyc src l.50-52): erode, rect ksize x ksize."""
    return erode(gray, ksize, ksize, iterations)


def otsu_binarize(gray):
    """morph_seq.otsu_binarize (pyc src l.55-59) — the value the reference computes (it forgets to return it)."""
    return threshold_binary(gray, int(otsu_threshold(gray)))


def binary_closing(binary, ksize=2, iterations=1):
    """morph_seq.binary_closing (pyc src l.62-68): threshold 127 then MORPH_CLOSE rect ksize."""
    return morph_close(threshold_binary(binary, 127), ksize, ksize, iterations)
