"""Host-side control path around the accelerated pixel path: file decode, page-quad localisation, skew-angle
estimate, optional stage dumps.  These parts of the reference (DocScanner.py:15-24, 76-109, 218-231,
282-346) produce a handful of scalars per page from irregular, sequential algorithms (contours, Hough
peaks) and are out of scope for the CUDA path (SURVEY.md §8f lists them as the next rows).  They run on
the host with OpenCV when it is installed; every function raises if it is not, so callers can instead
supply `quad=` / `angle=` themselves.  Since the skew estimate moved to the device (deskew.cu), what is left here is
localize_document's sequential half, decode and the dumps; its gray conversion and Canny already use the device kernels.
"""
from __future__ import annotations

import os

import numpy as np


def _cv2():
    try:
        import cv2
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("the control path (image decode, quad localisation, skew estimate) needs OpenCV on the "
                           "host; install opencv-python or pass quad= / angle= explicitly") from e
    return cv2


def load_image(path: str) -> np.ndarray:
    img = _cv2().imread(path, _cv2().IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(f"Cannot load image: {path}")
    return img


# ---- optional JPEG decode on the device -------------------------------------------------------------
# cv2.imread (DocScanner.py:15-19) decodes on the host and the 36 MB of raw pixels then cross PCIe; nvJPEG (a CUDA library
# that ships with the toolkit, bound here with ctypes) takes the ~2-4 MB file instead and leaves the BGR photo in device
# memory, where docscan_process_pages reads it directly.  OPT-IN: nvJPEG's inverse DCT and chroma upsampling are not
# bit-identical with libjpeg-turbo's, so a page decoded this way differs from the reference's by a grey level here and there
# and the chain after it is no longer bit-exact with DocScanner.py — use it when throughput matters more than that.
_nvjpeg_state = {}


def _nvjpeg():
    import ctypes as C
    if "lib" not in _nvjpeg_state:
        lib = None
        for name in ("libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so"):
            try:
                lib = C.CDLL(name)
                break
            except OSError:
                continue
        if lib is None:
            raise RuntimeError("decode='device' needs nvJPEG (libnvjpeg.so from the CUDA toolkit)")
        handle, jstate = C.c_void_p(), C.c_void_p()
        if lib.nvjpegCreateSimple(C.byref(handle)) != 0 or lib.nvjpegJpegStateCreate(handle, C.byref(jstate)) != 0:
            raise RuntimeError("nvJPEG initialisation failed")
        _nvjpeg_state.update(lib=lib, handle=handle, jstate=jstate)
    return _nvjpeg_state


def load_image_device(path: str, ctx=None):
    """JPEG file -> HxWx3 BGR photo in device memory (a _capi.DeviceBuffer), decoded by nvJPEG on the context's stream.
    Raises ValueError for files nvJPEG cannot decode (PNG, progressive modes it does not support ...): use load_image."""
    import ctypes as C
    from . import _capi
    ctx = ctx if ctx is not None else _capi.default_context()
    try:
        data = open(path, "rb").read()
    except OSError:
        raise FileNotFoundError(f"Cannot load image: {path}")
    st = _nvjpeg()
    lib = st["lib"]
    ncomp, subs = C.c_int(0), C.c_int(0)
    widths, heights = (C.c_int * 4)(), (C.c_int * 4)()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    if lib.nvjpegGetImageInfo(st["handle"], buf, C.c_size_t(len(data)), C.byref(ncomp), C.byref(subs), widths, heights) != 0:
        raise ValueError(f"nvJPEG cannot read {path}")
    w, h = int(widths[0]), int(heights[0])
    dev = _capi.DeviceBuffer(ctx, h, w, 3)

    class _NvjpegImage(C.Structure):
        _fields_ = [("channel", C.c_void_p * 4), ("pitch", C.c_size_t * 4)]

    out = _NvjpegImage()
    out.channel[0], out.pitch[0] = dev.ptr, dev.pitch
    NVJPEG_OUTPUT_BGRI = 6
    rc = lib.nvjpegDecode(st["handle"], st["jstate"], buf, C.c_size_t(len(data)), NVJPEG_OUTPUT_BGRI, C.byref(out), C.c_void_p(ctx.stream))
    if rc != 0:
        raise ValueError(f"nvJPEG failed to decode {path} (status {rc})")
    return dev


def save_image(path: str, img: np.ndarray) -> None:
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    _cv2().imwrite(path, img)


def preprocess(img, bilateral_d=9, bilateral_sigmaColor=75, bilateral_sigmaSpace=75, gaussian_ksize=0) -> np.ndarray:
    """DocScanner.py:39-45: the bilateral-denoised gray image the reference dumps as scan_01_pre.png and never reads
    again.  The gray conversion and the optional Gaussian run on the device; cv2.bilateralFilter has no device kernel
    here (dead output) and runs on the host."""
    from . import ops
    gray = ops.bgr2gray(img) if img.ndim == 3 else img
    den = _cv2().bilateralFilter(gray, bilateral_d, bilateral_sigmaColor, bilateral_sigmaSpace)
    if gaussian_ksize and gaussian_ksize > 1:
        den = ops.gaussian_blur(den, gaussian_ksize) if gaussian_ksize % 2 else _cv2().GaussianBlur(den, (gaussian_ksize, gaussian_ksize), 0)
    return den


def quad_overlay(color: np.ndarray, quad) -> np.ndarray:
    """scan_02_quad.png (DocScanner.py:300-308): the detected quad in green, or the frame in orange for whole-photo pages."""
    cv2 = _cv2()
    overlay = color.copy()
    if quad is not None:
        pts = np.asarray(quad).astype(np.int32).reshape((-1, 1, 2))
        cv2.polylines(overlay, [pts], True, (0, 255, 0), 2)
    else:
        h, w = color.shape[:2]
        full = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.int32).reshape((-1, 1, 2))
        cv2.polylines(overlay, [full], True, (0, 165, 255), 2)
    return overlay


def quad_area(quad) -> float:
    """cv2.contourArea of the float32 quad (DocScanner.py:291-292).  OpenCV accumulates the shoelace sum in double from
    float32 coordinates: `a00 += (double)xi_1 * yi - (double)xi * yi_1`, starting from the last point, then
    `fabs(a00 * 0.5)` -- the products of two float32 values are exact in double, so only the order of the sum matters."""
    q = np.asarray(quad, np.float32).reshape(4, 2)
    a00 = 0.0
    px, py = float(q[3, 0]), float(q[3, 1])
    for i in range(4):
        x, y = float(q[i, 0]), float(q[i, 1])
        a00 += px * y - x * py
        px, py = x, y
    return abs(a00 * 0.5)


def _ordered(pts: np.ndarray) -> np.ndarray:
    total, delta = pts.sum(axis=1), np.diff(pts, axis=1).ravel()
    return np.stack([pts[np.argmin(total)], pts[np.argmin(delta)], pts[np.argmax(total)], pts[np.argmax(delta)]]).astype(np.float32)


def localize_document(img, canny_low=50, canny_high=150, min_area_ratio=0.2, max_area_ratio=0.98):
    """Largest 4-gon among the external contours of (Canny edges OR their probabilistic-Hough segments);
    same recipe as DocScanner.py:76-109.  Returns TL, TR, BR, BL as float32 (4, 2), or None."""
    cv2 = _cv2()
    # BGR2GRAY + Canny (DocScanner.py:78-79) run on the device (bit-exact with cv2: tests/test_gpu_parity.py); the
    # probabilistic Hough transform, contour tracing and polygon fitting that follow are sequential and stay on the host
    from . import ops
    edges = ops.canny(ops.bgr2gray(img), canny_low, canny_high)
    segments = cv2.HoughLinesP(edges, 1, np.pi / 180, threshold=80, minLineLength=80, maxLineGap=10)
    strokes = np.zeros_like(edges)
    for seg in ([] if segments is None else segments):
        x1, y1, x2, y2 = seg[0]
        cv2.line(strokes, (x1, y1), (x2, y2), 255, 2)
    contours, _ = cv2.findContours(cv2.bitwise_or(edges, strokes), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not contours:
        return None
    frame = max(img.shape[0] * img.shape[1], 1)
    sized = [c for c in contours if min_area_ratio <= abs(cv2.contourArea(c)) / frame <= max_area_ratio]
    best, best_area = None, 0.0
    for c in (sized or contours):
        poly = cv2.approxPolyDP(c, 0.02 * cv2.arcLength(c, True), True)
        if len(poly) == 4 and abs(cv2.contourArea(poly)) > best_area:
            best, best_area = poly.reshape(-1, 2).astype(np.float32), abs(cv2.contourArea(poly))
    if best is None:
        best = cv2.boxPoints(cv2.minAreaRect(max(contours, key=cv2.contourArea))).astype(np.float32)
    return _ordered(best)


def estimate_skew_angle(gray, canny_low=50, canny_high=150, max_rotate=10.0) -> float:
    """Median Hough-line angle folded into (-90, 90], zero when it exceeds max_rotate (DocScanner.py:218-231).
    The arithmetic is kept in the dtype numpy gives it (float32 thetas), because the reference hands exactly
    that Python float to getRotationMatrix2D."""
    cv2 = _cv2()
    lines = cv2.HoughLines(cv2.Canny(gray, canny_low, canny_high), 1, np.pi / 180, 150)
    if lines is None or len(lines) == 0:
        return 0.0
    folded = [((theta * 180.0 / np.pi) + 90.0) % 180.0 - 90.0 for _, theta in lines[:, 0, :]]
    angle = float(np.median(folded))
    return 0.0 if abs(angle) > max_rotate else angle


_DUMP_NAMES = {"warped": "scan_03_warped.png", "illum": "scan_04_illum.png", "stretch": "scan_05_stretch.png",
               "inkmask": "scan_05a_inkmask.png", "adapt": "scan_06_adapt.png", "weighted": "scan_06b_weighted.png",
               "deskew": "scan_07_deskew.png", "clean": "scan_08_clean.png"}


def save_stage_dumps(out_dir: str, stages: dict) -> None:
    cv2 = _cv2()
    os.makedirs(out_dir, exist_ok=True)
    for key, name in _DUMP_NAMES.items():
        if key in stages:
            cv2.imwrite(os.path.join(out_dir, name), stages[key])
